"""Markdown results tables from the committed bench lines (profiles/bench_rNN_*.json).
usage: python profiles/make_results_table.py [r02] [--readme]   (--readme: rewrite the block between the
<!-- results:begin --> / <!-- results:end --> markers of README.md)
One row per block of every line: the headline (C2 training), `sample`, `train_256`, `sample_256_heun`,
`sample_256_text_cfg` of the default run, plus single-workload and multi-GPU lines when present."""
import glob
import json
import os
import sys

here = os.path.dirname(os.path.abspath(__file__))
argv = [a for a in sys.argv[1:] if not a.startswith("--")]
rnd = argv[0] if argv else "r02"
out = []


def emit(line=""):
    out.append(line)


def last_json(path):
    txt = [l for l in open(path).read().strip().split("\n") if l.startswith("{")]
    return json.loads(txt[-1]) if txt else None


def train_row(tag, d):
    e2e = d.get("e2e") or {}
    roof = d.get("roofline") or {}
    allk = roof.get("all_tensor_kernels") or {}
    return (tag, d.get("config", {}).get("workload", "")[:58], f"{d['value']:.1f} img/s", f"{d['ms_per_step']:.2f}",
            f"{e2e.get('value', 0):.1f}", f"{d.get('train_frac_of_sustained_bf16', 0):.3f}",
            f"{roof.get('kernel', '')}: {roof.get('frac', 0):.3f} ({100 * roof.get('share_of_step', 0):.0f} % of step); "
            f"all tensor kernels {allk.get('frac', 0):.3f}", str(d.get("launches_per_step", "")))


train, sample = [], []
for f in sorted(glob.glob(os.path.join(here, f"bench_{rnd}_*.json"))):
    d = last_json(f)
    if d is None:
        continue
    tag = os.path.basename(f)[len(f"bench_{rnd}_"):-5]
    if d.get("impl") == "reference":
        cb = d["cpu_baseline"]
        train.append((tag, d["config"]["workload"][:58], f"{d['value']:.1f} img/s (CPU oracle port, {cb['cores']} cores"
                      f"{', ' + cb['cpu'] if cb.get('cpu') else ''})", f"{d['ms_per_step']:.0f}", "", "", "", ""))
        continue
    if d.get("metric") == "train_images_per_sec":
        train.append(train_row(tag, d))
        if d.get("train_256"):
            train.append(train_row(tag + ":train_256", d["train_256"]))
        s = d.get("sample")
        if s:
            sample.append((tag + ":sample", f"{s['sampler']} {s['diffusion_steps']} steps, 64x64, B={s['batch_per_gpu']}",
                           f"{s['denoise_steps_per_sec']:.1f}", f"{s['image_steps_per_sec']:.0f}",
                           f"{s['tensor_frac_of_sustained']:.3f}"))
        for key in ("sample_256_heun", "sample_256_text_cfg"):
            blk = d.get(key)
            for r in (blk or {}).get("runs", []):
                sample.append((f"{tag}:{key}", f"{blk['sampler']} {r['diffusion_steps']} steps, 256x256, "
                               f"B={blk['batch_per_gpu']} ({r['unet_evals_per_image']} UNet evals / image)",
                               f"{r['denoise_steps_per_sec']:.1f}", f"{r['image_steps_per_sec']:.0f}",
                               f"{r['tensor_frac_of_sustained']:.3f}"))
    elif "runs" in d:
        for r in d["runs"]:
            sample.append((tag, f"{d['config']['sampler']} {r['diffusion_steps']} steps, B={d['config']['batch_per_gpu']}",
                           f"{r['denoise_steps_per_sec']:.1f}", f"{r['image_steps_per_sec']:.0f}",
                           f"{r['tensor_frac_of_sustained']:.3f}"))
emit("| line | workload | value (device-resident) | ms/step | e2e img/s | whole-step frac of sustained bf16 | "
      "dominant kernel: frac of sustained bf16 | launches/step |")
emit("|---|---|---|---|---|---|---|---|")
for r in train:
    emit("| " + " | ".join(r) + " |")
emit()
emit("| line | sampler | denoise steps/s | image-steps/s (all GPUs) | UNet FLOPs frac of sustained bf16 |")
emit("|---|---|---|---|---|")
for r in sample:
    emit("| " + " | ".join(r) + " |")
text = "\n".join(out)
if "--readme" in sys.argv:
    readme = os.path.join(os.path.dirname(here), "README.md")
    doc = open(readme).read()
    a, b = doc.index("<!-- results:begin -->"), doc.index("<!-- results:end -->")
    open(readme, "w").write(doc[:a] + "<!-- results:begin -->\n" + text + "\n" + doc[b:])
else:
    print(text)
