"""Generates tests/golden/orbax_metadata_keys.json from one of the reference's real orbax checkpoints
(/root/reference/pretrained/EDM Unconditional/.../1948/default/_METADATA): for the `state` subtree, every
leaf's stringified key path with its per-component key_type list and value_metadata, plus the file's
top-level keys.  The payloads are git-LFS pointers; this fixture pins the STRUCTURE our writer
(flaxdiff_b200/checkpoint.py) must reproduce.  Run in the build container:
    python tests/golden/make_orbax_metadata_fixture.py
"""
import glob
import json
import os

REF = "/root/reference/pretrained/EDM Unconditional"
path = sorted(glob.glob(os.path.join(REF, "*", "1948", "default", "_METADATA")))[0]
doc = json.load(open(path))
tm = doc["tree_metadata"]
state = {}
for k, v in tm.items():
    parts = eval(k)
    if parts[0] != "state":
        continue
    state[k] = {"key_types": [m["key_type"] for m in v["key_metadata"]], "value_metadata": v["value_metadata"]}
out = {
    "source": os.path.relpath(path, "/root/reference"),
    "doc_keys": sorted(doc.keys()),
    "use_zarr3": doc["use_zarr3"],
    "top_level": sorted({eval(k)[0] for k in tm}),
    "sibling_files": sorted(os.listdir(os.path.dirname(path))) + sorted(
        f for f in os.listdir(os.path.dirname(os.path.dirname(path))) if not os.path.isdir(
            os.path.join(os.path.dirname(os.path.dirname(path)), f))),
    "state": state,
}
json.dump(out, open(os.path.join(os.path.dirname(__file__), "orbax_metadata_keys.json"), "w"), indent=0)
print(len(state), out["top_level"], out["sibling_files"])
