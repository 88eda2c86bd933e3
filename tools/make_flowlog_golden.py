#!/usr/bin/env python
"""Machine-extracts the output column lists of the flow_log tables from the reference's own ToInput
functions -> tests/golden/flowlog_columns.json:
  {"l4_flow_log" | "l7_flow_log" | "flow_tag.custom_field" | ...: [[column name, ch-go container type]]}
in the order ToInput appends them (INSERTs bind by name; this is the order a drop-in block presents).
Sources (server/): libs/ckdb/column.go (COLUMN_* -> name), ingester/flow_log/log_data/l4_column_block.go,
l7_column_block.go, ingester/flow_tag/flow_tag_column_block.go, ingester/flow_tag/app_service_tag*.go.
Run in the build container (needs /root/reference); the JSON is committed."""
import json
import os
import re
import sys

REF = (sys.argv[1] if len(sys.argv) > 1 else "/root/reference") + "/server"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                   "flowlog_columns.json")


def read(p):
    return open(os.path.join(REF, p), encoding="utf-8").read()


def column_names():
    return {m.group(1): m.group(2)
            for m in re.finditer(r"^\s*(COLUMN_\w+)\s*=\s*\"([^\"]*)\"", read("libs/ckdb/column.go"), re.M)}


def blocks(src):
    """{struct: {"types": {field: container}, "embedded": [struct...]}} and {struct: ToInput body}."""
    structs, bodies = {}, {}
    for m in re.finditer(r"^type (\w+) struct \{(.*?)^\}", src, re.M | re.S):
        types, emb = {}, []
        for line in m.group(2).split("\n"):
            line = line.split("//")[0].strip()
            f = re.match(r"^(Col\w+)\s+\*?proto\.(Col[\w\[\]]+)", line)
            if f:
                types[f.group(1)] = f.group(2)
            elif re.match(r"^\*?\w+$", line):
                emb.append(line.lstrip("*"))
        structs[m.group(1)] = {"types": types, "embedded": emb}
    for m in re.finditer(r"^func \(b \*(\w+)\) ToInput\(input proto\.Input\) proto\.Input \{(.*?)^\}", src,
                         re.M | re.S):
        bodies[m.group(1)] = m.group(2)
    return structs, bodies


def expand(name, structs, bodies, names):
    """Columns of struct `name` in ToInput order (nested b.X.ToInput(input) calls expanded in place)."""
    out = []
    for tok in re.finditer(r"b\.(\w+)\.ToInput\(input\)|Name:\s*ckdb\.(COLUMN_\w+),\s*Data:\s*&?b\.(Col\w+)",
                           bodies[name]):
        if tok.group(1):
            if tok.group(1) == "NativeTagsBlock":  # per-org native tags: nil unless an operator configures them
                continue
            out += expand(tok.group(1), structs, bodies, names)
        else:
            out.append([names[tok.group(2)], structs[name]["types"][tok.group(3)]])
    return out


def main():
    names = column_names()
    out = {}
    for path, tops in (("ingester/flow_log/log_data/l4_column_block.go", {"l4_flow_log": "L4FlowLogBlock"}),
                       ("ingester/flow_log/log_data/l7_column_block.go", {"l7_flow_log": "L7FlowLogBlock"})):
        structs, bodies = blocks(read(path))
        if "KnowledgeGraphBlock" not in structs:  # l7 embeds the block declared next to the l4 ones
            s2, b2 = blocks(read("ingester/flow_log/log_data/l4_column_block.go"))
            structs.update({k: v for k, v in s2.items() if k not in structs})
            bodies.update({k: v for k, v in b2.items() if k not in bodies})
        for table, top in tops.items():
            out[table] = expand(top, structs, bodies, names)
    # flow_tag tables (flow_tag_column_block.go:50-67: custom_field = the first append, custom_field_value adds
    # field_value + count; app_service_tag_column_block.go)
    structs, bodies = blocks(read("ingester/flow_tag/flow_tag_column_block.go"))
    body = bodies["FlowTagBlock"]
    cut = body.index("if b.TagType == TagFieldValue")
    bodies["FlowTagBlock"] = body[:cut]
    out["flow_tag.custom_field"] = expand("FlowTagBlock", structs, bodies, names)
    bodies["FlowTagBlock"] = body
    out["flow_tag.custom_field_value"] = expand("FlowTagBlock", structs, bodies, names)
    structs, bodies = blocks(read("ingester/flow_tag/app_service_tag_column_block.go"))
    (top,) = list(bodies)
    out["flow_tag.app_service"] = expand(top, structs, bodies, names)
    with open(OUT, "w") as f:
        json.dump(out, f, indent=0)
    print(OUT, {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
