"""Generates tests/golden/param_names.json from the reference's orbax `_METADATA` files
(/root/reference/pretrained/*/<step>/default/_METADATA): the flax parameter-tree key paths of real
FlaxDiff UNet checkpoints (names only - the weights are git-LFS pointers and shapes are not
recorded).  Run in the build container (the reference tree does not exist on the GPU box):
    python tests/golden/make_param_name_fixture.py
"""
import glob
import json
import os

REF = "/root/reference/pretrained"
out = {}
for path in sorted(glob.glob(os.path.join(REF, "*", "*", "*", "default", "_METADATA"))):
    meta = json.load(open(path))["tree_metadata"]
    names = set()
    for k in meta:
        parts = eval(k)  # keys are stringified tuples
        if parts[:3] == ("state", "params", "params"):
            names.add("/".join(parts[3:]))
    rel = os.path.relpath(path, REF)
    out[rel] = sorted(names)
json.dump(out, open(os.path.join(os.path.dirname(__file__), "param_names.json"), "w"), indent=0)
for k, v in out.items():
    print(k, len(v))
