#!/usr/bin/env python
"""Writes tests/golden/res10_symbol_ops.json from the reference's committed MXNet symbol
(/root/reference/policy_value_loss.json, written by mxnet 1.5.1 for the 10-block residual net of
train_mxnet.py:79-91): the op nodes in graph order with their attrs and the NAMES of their inputs, plus the
parameter / input (null) nodes with their attrs.  The fixture is the structural pin of oracle/net.py (the
arithmetic itself lives in MXNet and stays unpinned, see oracle/net.py).
    python tests/golden/make_graph_golden.py [/root/reference]"""
import json
import os
import sys


def extract(path):
    g = json.load(open(path))
    nodes = g["nodes"]
    ops, nulls = [], []
    for n in nodes:
        if n["op"] == "null":
            nulls.append({"name": n["name"], "attrs": n.get("attrs") or {}})
        else:
            ops.append({"op": n["op"], "name": n["name"], "attrs": n.get("attrs") or {},
                        "inputs": [nodes[i[0]]["name"] for i in n["inputs"]]})
    heads = [nodes[h[0]]["name"] for h in g["heads"]]
    return {"source": "policy_value_loss.json", "mxnet_version": g["attrs"]["mxnet_version"][1], "heads": heads,
            "ops": ops, "nulls": nulls}


if __name__ == "__main__":
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "res10_symbol_ops.json")
    json.dump(extract(os.path.join(ref, "policy_value_loss.json")), open(out, "w"), indent=0, sort_keys=True)
    print("wrote", out)
