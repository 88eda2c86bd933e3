#!/usr/bin/env python
"""Machine-extracts L7Protocol.String(isTLS) (server/libs/datatype/flow.go: constants :150-183, String
:623-803) -> tests/golden/l7_protocols.json: {"names": {"<value>": [plain, tls]}, "default": "N/A"}.
Run in the build container (needs /root/reference); the JSON is committed."""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "l7_protocols.json")
src = open(os.path.join(REF, "server/libs/datatype/flow.go"), encoding="utf-8").read()
consts = {m.group(1): int(m.group(2)) for m in re.finditer(r"^\s*(L7_PROTOCOL_\w+)\s+L7Protocol\s*=\s*(\d+)", src, re.M)}
body = re.search(r"func \(p L7Protocol\) String\(isTLS bool\) string \{(.*?)\n\}\n", src, re.S).group(1)
names = {}
for m in re.finditer(r"case (L7_PROTOCOL_\w+):\s*(.*?)(?=\n\tcase |\n\tdefault:)", body, re.S):
    rets = re.findall(r'return "([^"]*)"', m.group(2))
    if "if isTLS" in m.group(2):
        tls, plain = rets[0], rets[1]
    else:
        tls = plain = rets[0]
    names[str(consts[m.group(1)])] = [plain, tls]
default = re.search(r'default:\s*return "([^"]*)"', body).group(1)
json.dump({"names": names, "default": default}, open(OUT, "w"), indent=0, sort_keys=True)
print(OUT, len(names), default)
