"""Extracts the reference's own golden vectors and fixture meshes into small data files.

Run in the build container (where /root/reference is mounted, read-only):
    python tests/golden/make_golden.py
It parses the `// Output:` blocks of /root/reference/example_test.go (the four ExampleGeneralAssembler*
functions, example_test.go:14-149) and the TSV meshes under /root/reference/examples/{ruc,bimetallic}
into tests/golden/*.json, which travel to the GPU box (the reference does not).
Only data (expected numbers, mesh coordinates, connectivity) is extracted -- no reference code.
"""
import json
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def parse_examples():
    src = open(os.path.join(REF, "example_test.go"), encoding="utf-8").read()
    out = {}
    for m in re.finditer(r"func (Example\w+)\(\)(.*?)\n}\n", src, re.S):
        name, body = m.group(1), m.group(2)
        rows = []
        for line in body.splitlines():
            mm = re.match(r"\s*//\s*[⎡⎢⎣\[]\s*(.*?)\s*[⎤⎥⎦\]]\s*$", line)
            if mm:
                rows.append([float(x) for x in mm.group(1).split()])
        fmt = re.search(r'Printf\("K=\\n(%[.\dgf]+)', body).group(1)
        out[name] = {"format": fmt, "K": rows}
    return out


def parse_ruc():
    nodes = []
    for line in open(os.path.join(REF, "examples/ruc/crettonRUC_70_nodes.tsv")):
        f = line.split()
        if len(f) >= 4:
            nodes.append([float(f[1]), float(f[2]), float(f[3])])
    elems = []
    for line in open(os.path.join(REF, "examples/ruc/crettonRUC_70_elements.tsv")):
        f = line.split()
        if len(f) >= 9:
            elems.append([int(x) - 1 for x in f[1:9]])      # 1-indexed (ruc/main.go:79)
    return {"nodes": nodes, "hexa8": elems}


def parse_bimetallic():
    nodemap, nodes = {}, []
    for line in open(os.path.join(REF, "examples/bimetallic/bimetallic_nod.tsv")):
        f = line.split()
        if len(f) >= 4:
            nodemap[int(f[0])] = len(nodes)                  # bimetallic/main.go:36
            nodes.append([float(f[1]), float(f[2]), float(f[3])])
    elems = []
    for line in open(os.path.join(REF, "examples/bimetallic/bimetallic_elem.tsv")):
        f = line.split()
        if len(f) >= 9:
            elems.append([nodemap[int(x)] for x in f[1:9]])  # bimetallic/main.go:46
    return {"nodes": nodes, "quad8": elems}


if __name__ == "__main__":
    json.dump(parse_examples(), open(os.path.join(HERE, "example_test_outputs.json"), "w"), indent=1)
    json.dump(parse_ruc(), open(os.path.join(HERE, "ruc_hexa8.json"), "w"))
    json.dump(parse_bimetallic(), open(os.path.join(HERE, "bimetallic_quad8.json"), "w"))
    print("golden fixtures written to", HERE)
