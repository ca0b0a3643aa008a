"""Markdown results table from the committed bench lines (profiles/bench_r01_*.json).
usage: python profiles/make_results_table.py"""
import glob
import json
import os

here = os.path.dirname(os.path.abspath(__file__))
rows = []
for f in sorted(glob.glob(os.path.join(here, "bench_r01_*.json"))):
    txt = [l for l in open(f).read().strip().split("\n") if l.startswith("{")]
    if not txt:
        continue
    d = json.loads(txt[-1])
    tag = os.path.basename(f)[len("bench_r01_"):-5]
    if d.get("impl") == "reference":
        rows.append((tag, f"{d['value']:.1f} {d['unit']} (CPU oracle port, {d['cpu_baseline']['cores']} cores)", "", "", "", ""))
        continue
    e2e = d.get("e2e") or {}
    roof = d.get("roofline") or {}
    smp = d.get("sample") or {}
    extra = ""
    if "runs" in d:
        extra = "; ".join(f"{r['diffusion_steps']} steps: {r['denoise_steps_per_sec']:.1f}/s ({r['unet_image_evals_per_sec']:.0f} evals/s)"
                          for r in d["runs"])
    elif smp:
        extra = f"Euler {smp['diffusion_steps']}: {smp['denoise_steps_per_sec']:.1f} steps/s"
    tf = d.get("train_frac_of_sustained_bf16")
    rows.append((tag, f"{d['value']:.1f} {d['unit']}", f"{d['ms_per_step']:.2f}",
                 f"{e2e.get('value', float('nan')):.1f}" if e2e.get("value") else "",
                 f"{roof.get('frac', 0):.3f}" + (f" / {tf:.3f}" if tf else ""), extra))
print("| run (workload_nGPUs) | value | ms/step | e2e | tc-kernel frac / whole-step frac of sustained bf16 | sampling |")
print("|---|---|---|---|---|---|")
for r in rows:
    print("| " + " | ".join(r) + " |")
