#!/usr/bin/env python3
"""Build tests/golden/geo_slice.json from the reference's geo data
(server/libs/geo/ip_info_mini.json) the way ip_info.py builds ip_info.go: region
names in order of first appearance with index 0 = "未知", '黑龙'/'内蒙' fixed up.
Only a slice of the table is kept (the /8 blocks the probes live in) so the fixture
stays small.  Run in the build container; /root/reference is absent on the GPU box."""
import json
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
names, idx, rows = ["未知"], {}, []


def fix(s):
    return {"黑龙": "黑龙江", "内蒙": "内蒙古"}.get(s, s)


for line in open(f"{ref}/server/libs/geo/ip_info_mini.json", encoding="utf-8"):
    d = json.loads(line)
    r = 0
    if "region" in d:
        if d["region"] not in idx:
            idx[d["region"]] = len(names)
            names.append(fix(d["region"]))
        r = idx[d["region"]]
    rows.append((d["ip_start"], d["ip_end"], r))
keep = [x for x in rows if (x[0] >> 24) in (223, 1, 36, 61)]
keep.sort()
probes = [[3748071168, names[[r for s, e, r in keep if s <= 3748071168 <= e][0]]]]
for s, e, r in keep[::97]:
    probes.append([s, names[r]])
    probes.append([e, names[r]])
json.dump({"source": "server/libs/geo/ip_info_mini.json (slice: /8 blocks 1,36,61,223)", "region_names": names,
           "entries": keep, "probes": probes}, open("tests/golden/geo_slice.json", "w", encoding="utf-8"),
          ensure_ascii=False)
print(len(keep), "entries", len(names), "regions", len(probes), "probes")
