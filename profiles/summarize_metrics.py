"""Per-kernel table from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --csv` pass over ONE eager training step.
usage: python profiles/summarize_metrics.py profiles/metrics_rNN_x.csv [--json key batch]
Prints time share, DRAM bytes and GB/s, time-weighted tensor-pipe activity per kernel; with --json it also
MERGES into profiles/roofline_traffic.json, under `key` (the bench workload), the per-launch DRAM bytes
(dram__bytes_read.sum + dram__bytes_write.sum) and duration of every kernel - bench.py's `roofline.traffic`."""
import collections
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
per = collections.defaultdict(dict)          # launch id -> metrics
name = {}
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    per[r[ix['ID']]][r[ix['Metric Name']]] = float(r[ix['Metric Value']].replace(',', ''))
    name[r[ix['ID']]] = re.sub(r'\(anonymous namespace\)::|<unnamed>::|void ', '', r[ix['Kernel Name']]).split('(')[0]
# exactly one step: from one loss_kernel launch (end of a forward) to the next
order = sorted(per, key=int)
marks = [i for i, k in enumerate(order) if name[k].startswith('loss_kernel')]
if len(marks) >= 2:
    keep = set(order[marks[0]:marks[1]])
    per = {k: v for k, v in per.items() if k in keep}
    print(f"# window cut to one step: launches {marks[0]}..{marks[1]} of {len(order)} (loss_kernel to loss_kernel)")
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])   # n, ns, rd, wr, tensor*ns
for k, m in per.items():
    a = agg[name[k]]
    t = m.get('gpu__time_duration.sum', 0.0)
    a[0] += 1
    a[1] += t
    a[2] += m.get('dram__bytes_read.sum', 0.0)
    a[3] += m.get('dram__bytes_write.sum', 0.0)
    a[4] += t * m.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0.0)
tot = sum(a[1] for a in agg.values())
print(f"launches {sum(a[0] for a in agg.values())}  total {tot / 1e6:.2f} ms (ncu: serialised, cold cache - compare SHARES)")
print(f"{'kernel':44s} {'n':>4s} {'ms':>8s} {'share':>6s} {'dram GB':>8s} {'GB/s':>7s} {'tensor%':>7s}")
tc_bytes = tc_ns = 0.0
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    gb = (a[2] + a[3]) / 1e9
    print(f"{k[:44]:44s} {a[0]:4d} {a[1] / 1e6:8.3f} {100 * a[1] / tot:5.1f}% {gb:8.2f} {gb / (a[1] / 1e9):7.0f} {a[4] / a[1]:7.1f}")
    if k.startswith(('fdx_tc_kernel', 'fdx_tct_kernel', 'fdx_wgrad9', 'fdx_attn')):
        tc_bytes += a[2] + a[3]
        tc_ns += a[1]
print(f"tensor-core engine (fdx_tc_kernel + fdx_tct_kernel + fdx_wgrad9): {tc_ns / 1e6:.2f} ms, {tc_bytes / 1e9:.2f} GB DRAM traffic per step")
if len(sys.argv) > 3 and sys.argv[2] == '--json':
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "roofline_traffic.json")
    doc = json.load(open(path)) if os.path.exists(path) else {}
    kern = {k: {"launches": a[0], "dram_bytes_per_launch": (a[2] + a[3]) / a[0], "time_us_per_launch": a[1] / a[0] / 1e3,
                "tensor_pipe_pct": a[4] / a[1] if a[1] else 0.0, "source": sys.argv[1]} for k, a in agg.items()}
    doc[sys.argv[3]] = {"batch_per_gpu": int(sys.argv[4]) if len(sys.argv) > 4 else None,
                        "dram_bytes_per_step": sum(a[2] + a[3] for a in agg.values()),
                        "tensor_engine_dram_bytes_per_step": tc_bytes, "source": sys.argv[1], "kernels": kern}
    json.dump(doc, open(path, "w"), indent=1)
    print(f"merged {len(kern)} kernels into {path} under '{sys.argv[3]}'")
