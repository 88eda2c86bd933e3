#!/usr/bin/env python
"""Machine-extracts the field lists of the wire messages on the path from the reference's .proto files
(message/flow_log.proto, message/metric.proto) into tests/golden/proto_fields.json:
  {file: {message: [[field name, number, type, "repeated" | ""]]}}
Run in the build container (needs /root/reference); the JSON is committed, the tests read only the JSON."""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "proto_fields.json")
FIELD = re.compile(r"^\s*(repeated\s+|optional\s+)?([A-Za-z_][\w.]*)\s+([A-Za-z_]\w*)\s*=\s*(\d+)\s*(\[[^\]]*\])?\s*;")


def parse(path):
    msgs, stack = {}, []
    for raw in open(path, encoding="utf-8"):
        line = raw.split("//")[0]
        m = re.match(r"^\s*(message|enum)\s+(\w+)\s*\{", line)
        if m:
            stack.append((m.group(1), m.group(2)))
            if m.group(1) == "message":
                msgs.setdefault(m.group(2), [])
            continue
        if re.match(r"^\s*\}", line):
            if stack:
                stack.pop()
            continue
        if stack and stack[-1][0] == "message":
            f = FIELD.match(line)
            if f:
                rep = "repeated" if (f.group(1) or "").strip() == "repeated" else ""
                msgs[stack[-1][1]].append([f.group(3), int(f.group(4)), f.group(2), rep])
    return msgs


def parse_enums(path, wanted):
    """{enum: {NAME: value}} of the named enums (SmartEncoding type codes come from trident.proto)."""
    src = open(path, encoding="utf-8").read()
    out = {}
    for m in re.finditer(r"enum\s+(\w+)\s*\{(.*?)\}", src, re.S):
        if m.group(1) in wanted:
            body = re.sub(r"//[^\n]*", "", m.group(2))
            out[m.group(1)] = {k: int(v) for k, v in re.findall(r"(\w+)\s*=\s*(-?\d+)\s*;", body)}
    return out


def main():
    out = {}
    for name in ("flow_log.proto", "metric.proto"):
        out[name] = parse(os.path.join(REF, "message", name))
    out["trident.proto:enums"] = parse_enums(os.path.join(REF, "message", "trident.proto"),
                                             {"AutoServiceType", "DeviceType", "ServiceType", "ServiceProtocol",
                                              "CidrType"})
    with open(OUT, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print(OUT, {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
