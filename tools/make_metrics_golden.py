#!/usr/bin/env python
"""Machine-extract the flow_metrics output column lists from the reference's own
ToInput functions -> tests/golden/metrics_columns.json.

Run in the build container (needs /root/reference); the JSON travels with the repo.
Sources (server/libs/):
  ckdb/column.go                         COLUMN_* -> column name
  flow-metrics/tag_column_block.go       TagBlock struct (container types) + ToInput (order, Code condition)
  flow-metrics/basic_meter_column_block.go, app_meter_column_block.go, usage_meter_column_block.go
  flow-metrics/tag.go                    Code bit values, table codes
"""
import json
import os
import re
import sys

REF = "/root/reference/server/libs"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "metrics_columns.json")


def read(p):
    return open(os.path.join(REF, p), encoding="utf-8").read()


def column_names():
    names = {}
    for m in re.finditer(r"^\s*(COLUMN_\w+)\s*=\s*\"([^\"]*)\"", read("ckdb/column.go"), re.M):
        names[m.group(1)] = m.group(2)
    return names


def struct_types(src):
    types = {}
    for m in re.finditer(r"^\s*(Col\w+)\s+\*?proto\.(Col[\w\[\]]+)", src, re.M):
        types[m.group(1)] = m.group(2)
    return types


def code_bits():
    """Evaluate the three iota const blocks of tag.go."""
    src = read("flow-metrics/tag.go")
    bits = {}
    for block_start, first_shift in ((r"IP Code = 0x1 << iota", 0), (r"IPPath Code = 0x100000 << iota", 20),
                                     (r"Direction Code = 0x10000000000 << iota", 40)):
        m = re.search(re.escape(block_start) + r"(.*?)\n\)", src, re.S)
        lines = [block_start.split()[0]] + [l.strip() for l in m.group(1).split("\n")]
        i = 0
        for l in lines:
            l = l.split("//")[0].strip()
            if not l:
                continue
            bits[l] = 1 << (first_shift + i)
            i += 1
    bits.pop("_", None)
    bits["TunnelIPID"] = 1 << 62
    return bits


def tag_columns(names):
    src = read("flow-metrics/tag_column_block.go")
    types = struct_types(src)
    body = re.search(r"func \(b \*TagBlock\) ToInput.*?\n}\n", src, re.S).group(0)
    cols, cond = [], None
    for line in body.split("\n"):
        m = re.search(r"if code&(\w+) != 0", line)
        if m:
            cond = m.group(1)
        elif re.match(r"^\t}", line):
            cond = None
        for c in re.finditer(r"Name: ckdb\.(COLUMN_\w+), Data: &?b\.(Col\w+)", line):
            cols.append({"cond": cond, "name": names[c.group(1)], "type": types[c.group(2)]})
    # LowCardinality members are built with new(proto.ColStr).LowCardinality()
    for m in re.finditer(r"(Col\w+):\s+new\(proto\.ColStr\)\.LowCardinality\(\)", src):
        for c in cols:
            pass
    lc = set(re.findall(r"(Col\w+):\s+new\(proto\.ColStr\)\.LowCardinality\(\)", src))
    lc_names = set()
    for c in re.finditer(r"Name: ckdb\.(COLUMN_\w+), Data: b\.(Col\w+)", body):
        if c.group(2) in lc:
            lc_names.add(names[c.group(1)])
    for c in cols:
        if c["name"] in lc_names:
            c["type"] = "ColLowCardinality[string]"
    return cols


def meter_columns(path, names, funcs):
    src = read(path)
    types = struct_types(src)
    cols = []
    for fn in funcs:
        body = re.search(r"func \(b \*" + fn + r"\) ToInput.*?\n}\n", src, re.S).group(0)
        for c in re.finditer(r"Name: ckdb\.(COLUMN_\w+), Data: &?b\.(Col\w+)", body):
            cols.append({"name": names[c.group(1)], "type": types[c.group(2)]})
    return cols


def table_codes(bits):
    src = read("flow-metrics/tag.go")
    env = dict(bits)
    for name in ("BaseCode", "BasePathCode", "BasePortCode", "NETWORK", "NETWORK_MAP", "APPLICATION",
                 "APPLICATION_MAP", "TRAFFIC_POLICY"):
        m = re.search(r"^\s*" + name + r"\s*=\s*([^\n]+)", src, re.M)
        val = 0
        for term in m.group(1).split("|"):
            val |= env[term.strip()]
        env[name] = val
    return {k: env[k] for k in ("NETWORK", "NETWORK_MAP", "APPLICATION", "APPLICATION_MAP", "TRAFFIC_POLICY")}


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference")
    names = column_names()
    bits = code_bits()
    out = {
        "source": "deepflowio/deepflow server/libs/flow-metrics/*_column_block.go, tag.go, ckdb/column.go",
        "code_bits": bits,
        "table_codes": table_codes(bits),
        "tag_columns": tag_columns(names),
        "flow_meter_columns": meter_columns("flow-metrics/basic_meter_column_block.go", names,
                                            ["TrafficBlock", "LatencyBlock", "PerformanceBlock", "AnomalyBlock", "FlowLoadBlock"]),
        "app_meter_columns": meter_columns("flow-metrics/app_meter_column_block.go", names,
                                           ["AppTrafficBlock", "AppLatencyBlock", "AppAnomalyBlock"]),
        "usage_meter_columns": meter_columns("flow-metrics/usage_meter_column_block.go", names, ["UsageMeterBlock"]),
    }
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)
    print(OUT, len(out["tag_columns"]), len(out["flow_meter_columns"]), len(out["app_meter_columns"]),
          len(out["usage_meter_columns"]))


if __name__ == "__main__":
    main()
