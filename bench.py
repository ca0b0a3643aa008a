#!/usr/bin/env python
"""bench.py - throughput of the UNet training / sampling hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo (libfdx kernels)
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle on host cores)

A "step" is one data-parallel training step of the EDM UNet (noise-add -> UNet fwd -> weighted
L2 -> UNet bwd -> [NCCL grad all-reduce] -> AdamW + EMA) on a synthetic batch.  Default workload
is BASELINE.json configs[1]: unconditional EDM UNet 64x64x3, bf16, batch 256 per GPU.  The same
run also measures the Euler sampler (denoise-steps/sec) on that workload.  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FWD_GFLOP = {  # algorithmic forward GFLOP per image (BASELINE.md $2)
    ("c2", 64): 18.17, ("c3", 256): 300.83, ("c5", 256): 300.83, ("c4", 256): 298.03,
}
WORKLOADS = {
    # name: (resolution, per-GPU batch, attention configs, description)
    "c2": (64, 256, (None, None, None, None), "EDM UNet 64x64x3 bf16 B=256/GPU (BASELINE configs[1])"),
    "c3": (256, 64, (None, None, None, {"heads": 8}), "EDM UNet 256x256x3 self-attn bf16 B=64/GPU (configs[2])"),
    # sampling-only workloads (python bench.py --workload c4|c5): value = denoise steps / s
    "c4": (256, 64, (None, {"heads": 8}, {"heads": 8}, {"heads": 8}),
           "text-cond UNet 256x256x3 (frozen random 77x768 text emb), CFG g=3, Euler-ancestral 30 steps, B=64 (configs[3])"),
    "c5": (256, 32, (None, None, None, {"heads": 8}),
           "EDM UNet 256x256x3 self-attn, Heun 18/50/100 steps, B=32/GPU (configs[4])"),
}


def usable_cores() -> int:
    """Host cores this process may actually use: affinity mask capped by the cgroup CPU quota
    (os.cpu_count() reports the whole host inside a container and oversubscribes the thread pool)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        n = os.cpu_count() or 1
    try:
        q = open("/sys/fs/cgroup/cpu.max").read().split()
        if q[0] != "max":
            n = min(n, max(1, int(int(q[0]) / int(q[1]))))
    except Exception:  # noqa: BLE001
        try:
            quota = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if quota > 0:
                n = min(n, max(1, quota // period))
        except Exception:  # noqa: BLE001
            pass
    return max(1, min(n, 64))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


_JSON_OUT = None


def claim_stdout():
    """Keep stdout for the ONE JSON line: everything else any library prints there (NCCL's version banner,
    for one) is sent to stderr for the rest of the run."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(obj) -> None:
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def roofline_traffic(workload: str, batch: int) -> dict:
    """DRAM bytes the tensor-core engine moved in one step, from the committed ncu pass (profiles/) of the
    same workload; null when no capture exists for this workload / batch."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "roofline_traffic.json")) as f:
            rec = json.load(f).get(workload)
        if rec and rec.get("batch_per_gpu") == batch:
            return {"traffic": rec["dram_bytes_per_step"], "traffic_unit": "bytes per step (all tc launches)",
                    "traffic_source": rec["source"]}
    except Exception:  # noqa: BLE001
        pass
    return {"traffic": None}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", os.environ.get("FDX_BENCH_SMI_MS", "200"), "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:  # noqa: BLE001
                pass
        sm.sort()
        # median of the upper half ~ clocks under load (idle samples at the edges excluded)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": (load[len(load) // 2] if load else None), "sm_max_mhz": (max(mx) if mx else None),
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """CPU reference arm: the oracle restatement of the reference's train step on the host cores."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from flaxdiff_b200.models.simple_unet import Unet
    from oracle import train_ref
    res, _, acfg, desc = WORKLOADS[args.workload]
    cores = usable_cores()
    torch.set_num_threads(cores)
    B = args.ref_batch
    model = Unet(attention_configs=acfg)
    fp = model.init(4, device=torch.device("cpu"))
    P = {k: v.clone().requires_grad_(True) for k, v in fp.named.items()}
    ema = {k: v.detach().clone() for k, v in P.items()}
    freqs = model._fourier_freqs("cpu")
    g = torch.Generator().manual_seed(0)
    img = torch.randint(0, 256, (B, res, res, 3), dtype=torch.uint8, generator=g)
    opt = {}
    times = []
    steps = max(1, min(args.steps, args.ref_max_steps))
    warm = max(1, min(args.warmup, 1))
    for i in range(warm + steps):
        noise = torch.randn(B, res, res, 3, generator=g)
        t = torch.randn(B, generator=g)
        t0 = time.perf_counter()
        train_ref.edm_train_step(P, opt, img, noise, t, freqs, attention_configs=acfg, ema=ema, step=i + 1)
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    val = B / (ms / 1e3)
    # sampler evaluations on the CPU
    x = torch.randn(B, res, res, 3, generator=g) * 80
    t0 = time.perf_counter()
    train_ref.karras_denoise_eval(P, x, torch.full((B,), 0.5), freqs, attention_configs=acfg)
    ev = time.perf_counter() - t0
    out = {
        "impl": "reference", "metric": "train_images_per_sec", "value": val, "unit": "images/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "sample": f"CPU oracle train step on B={B} images per step ({steps} timed)",
                   "resolution": res},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"oracle/train_ref.edm_train_step, B={B}, {steps} steps, torch fp32, {cores} threads",
                         "sample_unet_evals_per_sec": B / ev},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


# ------------------------------------------------------------------------------------------ this repo
def log(msg):
    if os.environ.get("FDX_BENCH_VERBOSE"):
        sys.stderr.write(f"[bench {time.strftime('%H:%M:%S')}] {msg}\n")
        sys.stderr.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="fdx", choices=["fdx", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override per-GPU batch")
    ap.add_argument("--sample-steps", type=int, default=50)
    ap.add_argument("--no-sample", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--ref-batch", type=int, default=16)
    ap.add_argument("--ref-max-steps", type=int, default=3)
    args = ap.parse_args()
    import faulthandler
    faulthandler.enable()
    if os.environ.get("FDX_BENCH_WATCHDOG"):
        faulthandler.dump_traceback_later(int(os.environ["FDX_BENCH_WATCHDOG"]), repeat=True, file=sys.stderr)
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    if args.workload in ("c4", "c5"):
        return run_sampling(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from flaxdiff_b200 import _lib, ops, utils
    from flaxdiff_b200.inputs import DiffusionInputConfig
    from flaxdiff_b200.models.simple_unet import Unet
    from flaxdiff_b200.predictors import KarrasPredictionTransform
    from flaxdiff_b200.samplers import EulerSampler
    from flaxdiff_b200.schedulers import EDMNoiseScheduler, KarrasVENoiseScheduler
    from flaxdiff_b200.trainer import GeneralDiffusionTrainer, adamw
    lib = _lib.load()
    lib.fdx_launch_count.restype = __import__("ctypes").c_ulonglong

    res, B, acfg, desc = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
    model = Unet(attention_configs=acfg, dtype=torch.bfloat16)
    trainer = GeneralDiffusionTrainer(
        model, adamw(2.7e-4), EDMNoiseScheduler(1, sigma_max=80, rho=7, sigma_data=0.5),
        DiffusionInputConfig("image", (res, res, 3), []), rngs=4, name="bench",
        model_output_transform=KarrasPredictionTransform(sigma_data=0.5), ema_decay=0.999, device=dev,
        use_cuda_graph=not args.no_graph)
    step_fn = trainer._define_train_step(B)
    gen = torch.Generator().manual_seed(1234 + rank)
    host_batches = [torch.randint(0, 256, (B, res, res, 3), dtype=torch.uint8, generator=gen).pin_memory()
                    for _ in range(2)]
    dev_batch = host_batches[0].to(dev)
    E = res * res * 3
    fwd_gflop = FWD_GFLOP[(args.workload, res)]

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    log('trainer built')
    # ---- device-resident timing (value) -------------------------------------------------------
    l0 = lib.fdx_launch_count()
    for i in range(args.warmup):
        trainer.state, loss, trainer.rngstate = step_fn(trainer.state, trainer.rngstate, {"image": dev_batch}, rank)
        torch.cuda.synchronize()
        log(f'warmup step {i} done')
    torch.cuda.synchronize()
    launches_eager_plus_capture = lib.fdx_launch_count() - l0
    # launches per step: measure one more eager (non-captured) replay-equivalent by counting a capture
    clocks = ClockSampler(local_rank)
    sync_all()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        trainer.state, loss, trainer.rngstate = step_fn(trainer.state, trainer.rngstate, {"image": dev_batch}, rank)
    e1.record()
    sync_all()
    ms_dev = e0.elapsed_time(e1) / args.steps
    log(f'device-resident: {ms_dev:.2f} ms/step')
    # ---- end-to-end timing: pinned host batch -> device each step, loss read back each step -----
    for i in range(2):                                  # untimed: first touch of each pinned batch / .item() path
        trainer.state, loss, trainer.rngstate = step_fn(trainer.state, trainer.rngstate,
                                                        {"image": host_batches[i % 2]}, rank)
        float(loss.item())
    sync_all()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    last = 0.0
    for i in range(args.steps):
        trainer.state, loss, trainer.rngstate = step_fn(trainer.state, trainer.rngstate,
                                                        {"image": host_batches[i % 2]}, rank)
        last = float(loss.item())                       # device -> host read of the step's result
    e3.record()
    sync_all()
    clk = clocks.stop()
    ms_e2e = e2.elapsed_time(e3) / args.steps
    log(f'e2e: {ms_e2e:.2f} ms/step')
    t = torch.tensor([ms_dev, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(t[0]), float(t[1])

    # launches per step: run one eager (graph-less) fwd/bwd to count kernels
    l1 = lib.fdx_launch_count()
    noise = torch.randn(B, res, res, 3, device=dev)
    tt = torch.randn(B, device=dev)
    trainer._fwd_bwd(dev_batch, noise, tt)
    torch.cuda.synchronize()
    launches_per_step = int(lib.fdx_launch_count() - l1) + 1   # + fused AdamW/EMA kernel

    # ---- tensor-core engine time inside one step (roofline of the dominant kernel) -------------
    log('counted launches')
    tc_ms, tc_calls = profile_tc(trainer, dev_batch, noise, tt)
    log(f'tc profile: {tc_ms:.2f} ms in {tc_calls} launches')
    hbm, tf_burst, tf_sust, src = load_peaks()
    step_tflop = 3 * fwd_gflop * B / 1e3
    tc_tflops = step_tflop / (tc_ms / 1e3) if tc_ms > 0 else 0.0

    # ---- sampling (Euler, BASELINE configs[1]) --------------------------------------------------
    sample = None
    if not args.no_sample:
        sampler = EulerSampler(model, KarrasVENoiseScheduler(1, sigma_max=80, rho=7, sigma_data=0.5).to(dev),
                               KarrasPredictionTransform(0.5), DiffusionInputConfig("image", (res, res, 3), []))
        params = trainer.state.ema_params
        n_s = args.sample_steps
        sampler.generate_samples(params, B, res, diffusion_steps=min(n_s, 4), start_step=1000, device=dev)  # warm + capture
        sync_all()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        imgs = sampler.generate_samples(params, B, res, diffusion_steps=n_s, start_step=1000, device=dev)
        s1.record()
        sync_all()
        ms_s = s0.elapsed_time(s1)
        ts = torch.tensor([ms_s], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        ms_s = float(ts[0])
        log(f'sampling: {ms_s:.1f} ms')
        sample = {"sampler": "EulerSampler", "diffusion_steps": n_s, "batch_per_gpu": B,
                  "denoise_steps_per_sec": n_s / (ms_s / 1e3),
                  "image_steps_per_sec": world * B * n_s / (ms_s / 1e3),
                  "ms_per_denoise_step": ms_s / n_s,
                  "tensor_frac_of_sustained": (fwd_gflop * B * n_s / 1e3) / (ms_s / 1e3) / tf_sust,
                  "finite": bool(torch.isfinite(imgs).all().item())}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * B / (ms_dev / 1e3)
    e2e_val = world * B / (ms_e2e / 1e3)
    out = {
        "metric": "train_images_per_sec", "value": value, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": desc, "resolution": res, "batch_per_gpu": B, "global_batch": B * world,
                   "parallelism": f"dp{world}", "params": model.layout().num_params,
                   "cuda_graph": not args.no_graph,
                   "l2": "per-step activation traffic (>= 9 GB) exceeds the 126 MB L2; no explicit flush needed"},
        "e2e": {"value": e2e_val, "unit": "images/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": B * E,
                "d2h_bytes_per_step": 4, "last_loss": last},
        "gpu_launches": launches_per_step * args.steps,
        "launches_per_step": launches_per_step,
        "train_tflops": step_tflop / (ms_dev / 1e3),
        "train_frac_of_sustained_bf16": step_tflop / (ms_dev / 1e3) / tf_sust,
        "roofline": {"bound": "tensor", "kernel": "tcgen05 engines fdx_tc / fdx_tct / fdx_wgrad9 (all conv and GEMM launches of one step)",
                     "achieved": tc_tflops, "peak": tf_sust, "unit": "TFLOP/s", "frac": tc_tflops / tf_sust,
                     "peak_source": f"{src} bf16_tflops_sustained", "launches": tc_calls, "ms_in_step": tc_ms,
                     "share_of_step": tc_ms / ms_dev, **roofline_traffic(args.workload, B)},
        "clocks": clk,
        "sample": sample,
    }
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, model, res, acfg)
    emit(out)
    if world > 1:
        dist.destroy_process_group()


def run_sampling(args):
    """Sampling-only workloads: c4 (CFG Euler-ancestral, text cross-attention) and c5 (Heun sweep)."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from flaxdiff_b200.inputs import ConditionalInputConfig, DiffusionInputConfig, RandomEmbeddingEncoder
    from flaxdiff_b200.models.simple_unet import Unet
    from flaxdiff_b200.predictors import KarrasPredictionTransform
    from flaxdiff_b200.samplers import EulerAncestralSampler, HeunSampler
    from flaxdiff_b200.schedulers import KarrasVENoiseScheduler
    res, B, acfg, desc = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
    hbm, tf_burst, tf_sust, src = load_peaks()
    sched = KarrasVENoiseScheduler(1, sigma_max=80, rho=7, sigma_data=0.5).to(dev)
    tr = KarrasPredictionTransform(0.5)
    fwd = FWD_GFLOP[(args.workload, res)]
    runs = []
    if args.workload == "c4":
        enc = RandomEmbeddingEncoder(77, 768, device=dev)
        cfg = DiffusionInputConfig("image", (res, res, 3), [ConditionalInputConfig(enc)])
        model = Unet(attention_configs=acfg, dtype=torch.bfloat16, context_dim=768)
        params = model.init(4 + rank, device=dev)
        smp = EulerAncestralSampler(model, sched, tr, cfg, guidance_scale=3.0)
        cond = (enc([f"prompt {i}" for i in range(B)]).to(dev),)
        plan = [(smp, 30, 2 * 30)]            # CFG doubles the model batch: 2 UNet evals per image-step
    else:
        cfg = DiffusionInputConfig("image", (res, res, 3), [])
        model = Unet(attention_configs=acfg, dtype=torch.bfloat16)
        params = model.init(4 + rank, device=dev)
        smp = HeunSampler(model, sched, tr, cfg)
        cond = ()
        plan = [(smp, n, 2 * n - 1) for n in (18, 50, 100)]
    clocks = ClockSampler(local_rank)
    first = True
    for smp_, n, nfe in plan:
        smp_.generate_samples(params, B, res, diffusion_steps=min(n, 3), start_step=1000, device=dev,
                              model_conditioning_inputs=cond)          # warm-up + graph capture
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if first:
            clocks.start()
            first = False
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        img = smp_.generate_samples(params, B, res, diffusion_steps=n, start_step=1000, device=dev,
                                    model_conditioning_inputs=cond)
        s1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([s0.elapsed_time(s1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms[0])
        runs.append({"diffusion_steps": n, "unet_evals_per_image": nfe, "ms": ms,
                     "denoise_steps_per_sec": n / (ms / 1e3), "image_steps_per_sec": world * B * n / (ms / 1e3),
                     "unet_image_evals_per_sec": world * B * nfe / (ms / 1e3),
                     "tensor_frac_of_sustained": (fwd * B * nfe / 1e3) / (ms / 1e3) / tf_sust,
                     "finite": bool(torch.isfinite(img).all().item())})
    clk = clocks.stop()
    if rank == 0:
        head = runs[0] if args.workload == "c4" else runs[1]
        out = {"metric": "denoise_steps_per_sec", "value": head["denoise_steps_per_sec"], "unit": "steps/s",
               "n_gpus": world, "steps": head["diffusion_steps"], "warmup": 3, "ms_per_step": head["ms"] / head["diffusion_steps"],
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
               "config": {"workload": desc, "resolution": res, "batch_per_gpu": B, "parallelism": f"replicas{world}",
                          "sampler": type(smp).__name__, "cuda_graph": True},
               "runs": runs, "clocks": clk,
               "roofline": {"bound": "tensor", "kernel": "UNet denoise evaluation (all libfdx launches of one step)",
                            "achieved": head["tensor_frac_of_sustained"] * tf_sust, "peak": tf_sust, "unit": "TFLOP/s",
                            "frac": head["tensor_frac_of_sustained"], "peak_source": f"{src} bf16_tflops_sustained",
                            "traffic": None}}
        emit(out)
    if world > 1:
        dist.destroy_process_group()


def profile_tc(trainer, images, noise, t):
    """CUDA-event time of every tcgen05 engine launch (conv3x3 fwd/dgrad/wgrad, gemm) in one eager step."""
    import torch
    from flaxdiff_b200 import ops
    names = ["conv3x3_fwd", "conv3x3_dgrad", "conv3x3_wgrad", "gemm", "upconv3x3_fwd", "upconv3x3_dgrad",
             "upconv3x3_wgrad"]
    orig = {n: getattr(ops, n) for n in names}
    events = []

    def wrap(fn):
        def inner(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn(*a, **k)
            e.record()
            events.append((s, e))
            return r
        return inner
    # one stream, no micro-batch / side-stream overlap: a kernel's duration is only meaningful when it runs alone
    saved_env = {k: os.environ.get(k) for k in ("FDX_MICROBATCH", "FDX_NO_SIDE")}
    os.environ["FDX_MICROBATCH"] = "1"
    os.environ["FDX_NO_SIDE"] = "1"
    try:
        for n in names:
            setattr(ops, n, wrap(orig[n]))
        trainer._fwd_bwd(images, noise, t)
        torch.cuda.synchronize()
    finally:
        for n in names:
            setattr(ops, n, orig[n])
        for k, v in saved_env.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return sum(s.elapsed_time(e) for s, e in events), len(events)


def cpu_baseline(args, model, res, acfg):
    """The oracle train step timed on the host cores on a bounded sample (rank 0, N=1 only)."""
    import torch
    from oracle import train_ref
    cores = usable_cores()
    torch.set_num_threads(cores)
    B = args.ref_batch
    fp = model.init(4, device=torch.device("cpu"))
    P = {k: v.clone().requires_grad_(True) for k, v in fp.named.items()}
    freqs = model._fourier_freqs("cpu")
    g = torch.Generator().manual_seed(0)
    img = torch.randint(0, 256, (B, res, res, 3), dtype=torch.uint8, generator=g)
    opt = {}
    times = []
    n = 2 if res <= 64 else 1
    for i in range(1 + n):
        noise = torch.randn(B, res, res, 3, generator=g)
        t = torch.randn(B, generator=g)
        t0 = time.perf_counter()
        train_ref.edm_train_step(P, opt, img, noise, t, freqs, attention_configs=acfg, step=i + 1)
        if i >= 1:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {"value": B / sec, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"oracle/train_ref.edm_train_step on B={B} images at {res}x{res}, {n} timed step(s), torch fp32"}


if __name__ == "__main__":
    main()
