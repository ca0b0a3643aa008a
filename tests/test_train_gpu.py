"""Optimiser variants, TrainState API, DynamicScale, EMA / shadow freshness, orbax-layout checkpoints,
`fit` with validation sampling and the uint8 prefetcher - on the GPU, through the C-ABI, against the oracle
(oracle/diffusion_ref.py, oracle/train_ref.py).  Tolerances: f32 streaming kernels <= 1e-5 relative."""
import os

import numpy as np
import pytest
import torch

from flaxdiff_b200 import ops, utils
from flaxdiff_b200._lib import OPT_ADAM, OPT_LAMB, FdxError
from flaxdiff_b200.inputs import DiffusionInputConfig
from flaxdiff_b200.models.params import FlatParams
from flaxdiff_b200.models.simple_unet import Unet
from flaxdiff_b200.predictors import KarrasPredictionTransform
from flaxdiff_b200.samplers import EulerSampler
from flaxdiff_b200.schedulers import EDMNoiseScheduler, KarrasVENoiseScheduler
from flaxdiff_b200.trainer import (DevicePrefetcher, DynamicScale, GeneralDiffusionTrainer, TrainState, adamw, chain,
                                   clip_by_global_norm, lamb)
from oracle import diffusion_ref as R
from oracle import train_ref

pytestmark = pytest.mark.gpu
dev = torch.device("cuda")


def rel(a, b):
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _bufs(n, seed=3):
    rng = np.random.default_rng(seed)
    p, g, m, v = (rng.standard_normal(n).astype(np.float32) for _ in range(4))
    return p, g, m, np.abs(v)


@pytest.mark.parametrize("clip,expect_clipped", [(0.5, True), (1e6, False)])
def test_clip_by_global_norm_matches_oracle(clip, expect_clipped):
    """optax.chain(clip_by_global_norm, adamw) (training.py:604-608): fdx_grad_stats + the fused step."""
    n = 8192
    p, g, m, v = _bufs(n)
    tp, tg, tm, tv = (torch.from_numpy(a.copy()).to(dev) for a in (p, g, m, v))
    te = tp.clone()
    gs = ops.grad_stats(tg)
    assert abs(gs[0].item() - float((g.astype(np.float64) ** 2).sum())) / gs[0].item() < 1e-5 and gs[1].item() == 0
    ops.optimizer_step(OPT_ADAM, tp, tg, tm, tv, te, None, 1e-3, 0.9, 0.999, 1e-8, 1e-4, 3, 0.999, grad_scale=0.5,
                       gstats=gs, clip_norm=clip)
    (gc,), nrm = R.clip_by_global_norm([0.5 * g], clip)
    assert (nrm > clip) == expect_clipped
    wp, wm, wv, we = R.adamw_ema(p.astype(np.float64), gc.astype(np.float64), m.astype(np.float64),
                                 v.astype(np.float64), p.astype(np.float64), 3, 1e-3, wd=1e-4)
    np.testing.assert_allclose(tp.cpu().numpy(), wp, rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(tm.cpu().numpy(), wm, rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(te.cpu().numpy(), we, rtol=2e-5, atol=1e-7)


def test_lamb_matches_oracle_per_tensor_trust_ratio():
    """optax.lamb over a three-tensor layout (offsets multiples of 64, zero padding between tensors)."""
    shapes = [(100,), (3, 3, 8, 16), (64,)]
    offs, off = [], 0
    for s in shapes:
        offs.append(off)
        off += (int(np.prod(s)) + 63) // 64 * 64
    total = off
    rng = np.random.default_rng(5)
    P, G, M, V = (np.zeros(total, np.float32) for _ in range(4))
    leaves = []
    for o, s in zip(offs, shapes):
        k = int(np.prod(s))
        P[o:o + k] = rng.standard_normal(k) * (0.0 if s == (64,) else 1.0)        # third tensor: ||p|| = 0
        G[o:o + k] = rng.standard_normal(k)
        M[o:o + k] = 0.1 * rng.standard_normal(k)
        V[o:o + k] = np.abs(rng.standard_normal(k))
        leaves.append(slice(o, o + k))
    tp, tg, tm, tv = (torch.from_numpy(a.copy()).to(dev) for a in (P, G, M, V))
    te = tp.clone()
    sh = torch.empty(total, dtype=torch.bfloat16, device=dev)
    seg = torch.tensor(offs, dtype=torch.int64, device=dev)
    ops.optimizer_step(OPT_LAMB, tp, tg, tm, tv, te, sh, 2e-3, 0.9, 0.999, 1e-6, 0.01, 5, 0.99,
                       seg_offsets=seg, seg_norms=torch.empty(2 * len(offs), device=dev),
                       u_ws=torch.empty(total, device=dev))
    wp, wm, wv = R.lamb_step([P[s] for s in leaves], [G[s] for s in leaves], [M[s] for s in leaves],
                             [V[s] for s in leaves], 5, 2e-3, eps=1e-6, wd=0.01)
    got = tp.cpu().numpy()
    for s, w, wmm in zip(leaves, wp, wm):
        np.testing.assert_allclose(got[s], w, rtol=3e-5, atol=1e-6)
        np.testing.assert_allclose(tm.cpu().numpy()[s], wmm, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(te.cpu().numpy(), 0.99 * P + 0.01 * got, rtol=1e-5, atol=1e-6)
    assert rel(sh.float(), tp) < 4e-3


def test_apply_gradients_then_apply_ema_equals_fused_step():
    """flax TrainState API (general_diffusion_trainer.py:327-330): state.apply_gradients(grads=g).apply_ema(d)."""
    model = Unet(attention_configs=(None,) * 4)
    torch.manual_seed(0)
    fp = model.init(4, device=dev)
    grads = FlatParams(fp.layout, torch.randn_like(fp.flat) * 1e-2)
    a = TrainState.create(model.apply, fp.clone(), fp.clone(), adamw(1e-3))
    b = TrainState.create(model.apply, fp.clone(), fp.clone(), adamw(1e-3))
    for _ in range(2):
        a = a.apply_gradients(grads=grads).apply_ema(0.9)
        b.apply_gradients_and_ema(grads, 0.9)
    assert a.step == b.step == 2 and a.opt_state["count"] == 2
    assert torch.equal(a.params.flat, b.params.flat)
    assert rel(a.ema_params.flat, b.ema_params.flat) < 1e-6
    assert not torch.equal(a.ema_params.flat, fp.flat)
    # the bf16 shadow follows the parameters, and a flax-named tree of gradients is accepted
    assert rel(a.params.shadow_flat().float(), a.params.flat) < 4e-3
    tree = {"params": {k: v for k, v in grads["params"].items()}}
    c = TrainState.create(model.apply, fp.clone(), fp.clone(), adamw(1e-3)).apply_gradients(grads=tree)
    d = TrainState.create(model.apply, fp.clone(), fp.clone(), adamw(1e-3)).apply_gradients(grads=grads)
    # (compare the parameter TENSORS: `grads.flat` above also carries random values in the 64-element alignment
    #  gaps between tensors, which a tree of leaves cannot)
    assert all(torch.equal(c.params.named[k], d.params.named[k]) for k in c.params.named)
    assert torch.equal(c.ema_params.flat, fp.flat)


def test_dynamic_scale_step_semantics():
    """general_diffusion_trainer.py:305-318 + flax DynamicScale: gradients arrive multiplied by the scale;
    a finite step equals the unscaled one; a non-finite step leaves params / moments / count untouched, still
    applies the EMA, and halves the scale; `growth_interval` finite steps double it."""
    n = 4096
    p, g, m, v = _bufs(n, 9)
    model_lay = Unet(attention_configs=(None,) * 4).layout()

    def mk(ds):
        lay = model_lay
        flat = torch.zeros(lay.total, device=dev)
        flat[:n] = torch.from_numpy(p)
        st = TrainState.create(None, FlatParams(lay, flat), FlatParams(lay, flat.clone()), adamw(1e-3), dynamic_scale=ds)
        st.opt_state["mu"][:n] = torch.from_numpy(m).to(dev)
        st.opt_state["nu"][:n] = torch.from_numpy(v).to(dev)
        return st

    def grads_of(st, scale):
        gf = torch.zeros_like(st.params.flat)
        gf[:n] = torch.from_numpy(g).to(dev) * scale
        return FlatParams(st.params.layout, gf)

    ds = DynamicScale(growth_interval=2)
    a, b = mk(ds), mk(None)
    assert ds.scale == 65536.0
    a.apply_gradients_and_ema(grads_of(a, 65536.0), 0.9)
    b.apply_gradients_and_ema(grads_of(b, 1.0), 0.9)
    assert rel(a.params.flat, b.params.flat) < 1e-6 and rel(a.ema_params.flat, b.ema_params.flat) < 1e-6
    assert (ds.scale, ds.fin_steps) == R.dynamic_scale_update(65536.0, 0, True, growth_interval=2)
    # non-finite gradients
    before_p, before_m = a.params.flat.clone(), a.opt_state["mu"].clone()
    before_e = a.ema_params.flat.clone()
    bad = grads_of(a, 65536.0)
    bad.flat[17] = float("inf")
    a.apply_gradients_and_ema(bad, 0.9)
    assert torch.equal(a.params.flat, before_p) and torch.equal(a.opt_state["mu"], before_m)
    assert rel(a.ema_params.flat, 0.9 * before_e + 0.1 * before_p) < 1e-6          # apply_ema still ran
    assert (ds.scale, ds.fin_steps) == R.dynamic_scale_update(65536.0, 1, False, growth_interval=2)
    assert a.step == 2
    # the skipped step must not advance the optimiser count: the next finite step is step 2 on both sides
    a.apply_gradients_and_ema(grads_of(a, ds.scale), 0.9)
    b.apply_gradients_and_ema(grads_of(b, 1.0), 0.9)
    assert a.opt_state["count"] == 2 == b.opt_state["count"]
    assert rel(a.params.flat, b.params.flat) < 1e-6
    # growth after `growth_interval` finite steps in a row
    s, f = ds.scale, ds.fin_steps
    for _ in range(3):
        a.apply_gradients_and_ema(grads_of(a, ds.scale), 0.9)
        s, f = R.dynamic_scale_update(s, f, True, growth_interval=2)
        assert (ds.scale, ds.fin_steps) == (s, f)
    assert ds.scale == 65536.0        # halved once, doubled once


def _trainer(res=16, graph=True, opt=None, **kw):
    model = Unet(attention_configs=(None,) * 4, dtype=torch.bfloat16)
    tr = GeneralDiffusionTrainer(model, opt or adamw(1e-3), EDMNoiseScheduler(1, sigma_max=80, rho=7, sigma_data=0.5),
                                 DiffusionInputConfig("image", (res, res, 3), []), rngs=4,
                                 model_output_transform=KarrasPredictionTransform(0.5), device=dev,
                                 use_cuda_graph=graph, **kw)
    return model, tr


@pytest.mark.parametrize("graph", [False, True])
def test_dynamic_scale_training_matches_unscaled(graph, monkeypatch):
    """use_dynamic_scale=True: loss and parameter update equal the unscaled run (power-of-two scale)."""
    torch.manual_seed(0)
    res, B = 16, 4
    img = torch.randint(0, 256, (B, res, res, 3), dtype=torch.uint8)
    outs = []
    for dyn in (False, True):
        _, tr = _trainer(res, graph, use_dynamic_scale=dyn)
        step = tr._define_train_step(B)
        g = torch.Generator().manual_seed(5)
        for it in range(2):
            noise, t = torch.randn(B, res, res, 3, generator=g), torch.randn(B, generator=g)
            monkeypatch.setattr(utils, "device_normal", lambda key, shape, device, dtype=torch.float32, _n=noise, _t=t:
                                (_t if len(shape) == 1 else _n).to(device))
            tr.state, loss, tr.rngstate = step(tr.state, tr.rngstate, {"image": img.clone()}, 0)
        outs.append((loss.item(), tr.state.params.flat.clone(), tr.state.ema_params.flat.clone()))
        if dyn:
            assert tr.state.dynamic_scale.scale == 65536.0 and tr.state.dynamic_scale.fin_steps == 2
    assert abs(outs[0][0] - outs[1][0]) / outs[0][0] < 1e-3
    p0 = Unet(attention_configs=(None,) * 4).init(utils.split(utils.PRNGKey(4))[1], device=dev).flat
    da, db = (outs[0][1] - p0).double(), (outs[1][1] - p0).double()
    # Adam's first steps are sign-like: elements whose gradient is ~0 flip with the f32-atomic summation order,
    # so two RUNS of the same step differ element-wise; direction and magnitude of the update must agree
    assert (da * db).sum() / (da.norm() * db.norm()) > 0.9
    assert abs(da.norm() / db.norm() - 1) < 0.05
    assert rel(outs[0][2], outs[1][2]) < 1e-3


@pytest.mark.parametrize("kind", ["clip", "lamb"])
def test_train_step_with_clipping_and_lamb_vs_oracle(kind, monkeypatch):
    """One EDM training step through the trainer with optax.chain(clip_by_global_norm(c), adamw) / optax.lamb
    against the oracle's gradients pushed through the oracle's optimiser rule."""
    torch.manual_seed(0)
    res, B = 16, 4
    opt = chain(clip_by_global_norm(0.05), adamw(1e-3)) if kind == "clip" else lamb(1e-3, weight_decay=0.01)
    model, tr = _trainer(res, True, opt)
    P = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in tr.state.params.named.items()}
    p_before = {k: v.detach().clone() for k, v in P.items()}
    freqs = model._fourier_freqs(dev).cpu()
    img = torch.randint(0, 256, (B, res, res, 3), dtype=torch.uint8)
    noise, t = torch.randn(B, res, res, 3), torch.randn(B)
    monkeypatch.setattr(utils, "device_normal", lambda key, shape, device, dtype=torch.float32:
                        (t if len(shape) == 1 else noise).to(device))
    step = tr._define_train_step(B)
    tr.state, loss, tr.rngstate = step(tr.state, tr.rngstate, {"image": img.clone()}, 0)
    # oracle gradients
    data = (img.float() - 127.5) / 127.5
    sigma = torch.exp(t * 1.2 - 1.2).view(-1, 1, 1, 1)
    c_in, c_out, c_skip = (torch.from_numpy(a).view(-1, 1, 1, 1) for a in R.karras_coeffs(sigma.numpy().reshape(-1)))
    x_t = data + sigma * noise
    from oracle import unet_ref
    F = unet_ref.unet_forward(P, x_t * c_in, torch.log(sigma.view(-1) + 1e-12) / 4, freqs)
    w = torch.from_numpy(R.karras_weight(sigma.numpy().reshape(-1))).view(-1, 1, 1, 1)
    want_loss = (0.5 * (c_out * F + c_skip * x_t - data) ** 2 * w).mean()
    assert abs(loss.item() - want_loss.item()) / want_loss.item() < 2e-2
    grads = torch.autograd.grad(want_loss, list(P.values()))
    names = list(P)
    gl = [g.numpy() for g in grads]
    zeros = [np.zeros_like(g) for g in gl]
    if kind == "clip":
        gl, nrm = R.clip_by_global_norm(gl, 0.05)
        assert nrm > 0.05                                   # the clip is active in this test
        want = [R.adamw_ema(p_before[k].numpy(), g, z, z.copy(), p_before[k].numpy(), 1, 1e-3, wd=1e-4)[0]
                for k, g, z in zip(names, gl, zeros)]
    else:
        want, _, _ = R.lamb_step([p_before[k].numpy() for k in names], gl, zeros, [z.copy() for z in zeros], 1, 1e-3,
                                 eps=1e-6, wd=0.01)
    num = den = 0.0
    for k, wv in zip(names, want):
        d_got = tr.state.params.named[k].cpu().numpy() - p_before[k].numpy()
        d_ref = wv - p_before[k].numpy()
        num += float(((d_got - d_ref) ** 2).sum())
        den += float((d_ref ** 2).sum())
    # first Adam-type step is sign-like: tiny gradients flip sign under bf16 noise -> compare in aggregate
    assert (num / den) ** 0.5 < 0.35, (num / den) ** 0.5
    if kind == "clip":
        # the global norm the kernel clipped with (f32 sumsq of the bf16-path gradients) vs the oracle's
        gn = ops.grad_stats(tr._grads.flat)[0].sqrt().item()
        assert abs(gn - nrm) / nrm < 3e-2


def test_ema_weights_stay_fresh_across_train_sample_train_sample():
    """the sampler must see the CURRENT EMA weights after further training steps - the fused
    optimiser writes the buffers behind torch's version counter."""
    torch.manual_seed(0)
    res, B = 16, 2
    model, tr = _trainer(res, True, adamw(5e-2), ema_decay=0.5)
    smp = EulerSampler(model, KarrasVENoiseScheduler(1, sigma_max=80, rho=7, sigma_data=0.5).to(dev),
                       KarrasPredictionTransform(0.5), DiffusionInputConfig("image", (res, res, 3), []))
    step = tr._define_train_step(B)
    img = torch.randint(0, 256, (B, res, res, 3), dtype=torch.uint8)
    prior = torch.randn(B, res, res, 3, device=dev) * 80
    outs = []
    for rnd in range(2):
        for _ in range(2):
            tr.state, _, tr.rngstate = step(tr.state, tr.rngstate, {"image": img}, 0)
        ema = tr.state.ema_params
        out = smp.generate_samples(ema, B, res, diffusion_steps=3, start_step=1000, priors=prior.clone(), device=dev)
        assert rel(ema.shadow_flat_noupdate().float(), ema.flat) < 4e-3, rnd        # shadow == current EMA
        eager = EulerSampler(model, smp.noise_schedule, smp.model_output_transform, smp.input_config,
                             use_cuda_graph=False)
        ref = eager.generate_samples(ema.clone(), B, res, diffusion_steps=3, start_step=1000, priors=prior.clone(),
                                     device=dev)
        assert rel(out, ref) < 2e-2, rnd
        outs.append(out)
    assert len(smp._graphs) == 2                      # (whole-step graph, last-step evaluation graph), both re-used
    assert rel(outs[0], outs[1]) > 1e-3               # the weights did change between the two samplings


def test_sampler_graph_cache_is_bounded_and_tree_params_are_packed_once():
    """a flax tree of params and fresh conditioning tensors must not grow the caches."""
    from flaxdiff_b200.inputs import ConditionalInputConfig, RandomEmbeddingEncoder
    torch.manual_seed(0)
    res, B = 16, 2
    model = Unet(attention_configs=(None, None, None, {"heads": 8}), dtype=torch.bfloat16, context_dim=768)
    fp = model.init(4, device=dev)
    enc = RandomEmbeddingEncoder(77, 768, device=dev)
    cfg = DiffusionInputConfig("image", (res, res, 3), [ConditionalInputConfig(enc)])
    smp = EulerSampler(model, KarrasVENoiseScheduler(1, sigma_max=80, rho=7, sigma_data=0.5).to(dev),
                       KarrasPredictionTransform(0.5), cfg, guidance_scale=2.0)
    tree = {"params": {k: v for k, v in fp["params"].items()}}
    outs = []
    for i in range(3):
        cond = enc([f"p{i}", f"q{i}"]).to(dev)                 # a NEW tensor every call
        outs.append(smp.generate_samples(tree, B, res, diffusion_steps=2, start_step=1000,
                                         priors=torch.ones(B, res, res, 3, device=dev) * 40,
                                         model_conditioning_inputs=(cond,), device=dev))
    assert len(smp._graphs) == 2 and len(smp._trees) == 1      # whole-step graph + last-step evaluation graph
    assert rel(outs[0], outs[1]) > 1e-4                        # conditioning really reaches the captured graph
    for i in range(8):                                         # different batch sizes: the LRU stays bounded
        smp.generate_samples(tree, 1 + i % 6, res, diffusion_steps=1, start_step=1000,
                             model_conditioning_inputs=(enc(["a"] * (1 + i % 6)).to(dev),), device=dev)
    assert len(smp._graphs) <= smp.MAX_GRAPHS


def test_checkpoint_roundtrip_orbax_layout(tmp_path):
    """save -> train further -> load restores params / EMA / moments / count / rngs / best_loss, refreshes the
    bf16 shadow so the captured graph computes with the restored weights; a different architecture raises."""
    torch.manual_seed(0)
    res, B = 16, 2
    model, tr = _trainer(res, True, checkpoint_base_path=str(tmp_path), name="ck test")
    step = tr._define_train_step(B)
    img = torch.randint(0, 256, (B, res, res, 3), dtype=torch.uint8)
    for _ in range(2):
        tr.state, loss, tr.rngstate = step(tr.state, tr.rngstate, {"image": img}, 0)
    tr.best_loss = 0.25
    d = tr.save(epoch=1, step=2)
    assert sorted(os.listdir(os.path.join(d, "default"))) == ["_METADATA", "checkpoint"]
    saved = {k: v.clone() for k, v in (("p", tr.state.params.flat), ("e", tr.state.ema_params.flat),
                                       ("m", tr.state.opt_state["mu"]), ("v", tr.state.opt_state["nu"]))}
    rng_saved = tr.rngstate.rng
    tr.state, l_next, tr.rngstate = step(tr.state, tr.rngstate, {"image": img}, 0)      # the step after the save
    p_next = tr.state.params.flat.clone()
    for _ in range(2):
        tr.state, _, tr.rngstate = step(tr.state, tr.rngstate, {"image": img}, 0)
    assert not torch.equal(tr.state.params.flat, saved["p"])
    got_step, st, best, rs = tr.load()
    assert got_step == 2 and st.step == 2 and st.opt_state["count"] == 2 and rs.rng == rng_saved
    assert tr.best_loss == 0.25
    assert torch.equal(st.params.flat, saved["p"]) and torch.equal(st.ema_params.flat, saved["e"])
    assert torch.equal(st.opt_state["mu"], saved["m"]) and torch.equal(st.opt_state["nu"], saved["v"])
    assert rel(st.params.shadow_flat_noupdate().float(), st.params.flat) < 4e-3
    # replaying the captured graph after load reproduces the step that followed the save
    tr.state, l_again, tr.rngstate = step(tr.state, tr.rngstate, {"image": img}, 0)
    assert abs(l_again.item() - l_next.item()) / l_next.item() < 5e-3
    assert rel(tr.state.params.flat - saved["p"], p_next - saved["p"]) < 0.1
    # a checkpoint of another architecture is rejected, not loaded with permuted offsets
    other = Unet(attention_configs=(None, None, None, {"heads": 8}), dtype=torch.bfloat16)
    tr2 = GeneralDiffusionTrainer(other, adamw(1e-3), EDMNoiseScheduler(1), DiffusionInputConfig("image", (res, res, 3), []),
                                  rngs=4, model_output_transform=KarrasPredictionTransform(0.5), device=dev,
                                  checkpoint_base_path=str(tmp_path), name="ck test")
    with pytest.raises(FdxError, match="do not match"):
        tr2.load()


def test_best_state_is_an_independent_copy():
    _, tr = _trainer(16, False)
    assert tr.best_state is not tr.state
    assert tr.best_state.params.flat.data_ptr() != tr.state.params.flat.data_ptr()
    assert torch.equal(tr.best_state.params.flat, tr.state.params.flat)


def test_fit_runs_validation_sampling_from_ema_and_prefetches_uint8():
    """fit (simple_trainer.py:601-677): sanity validation, train with the uint8 prefetch thread, validation
    sampling from the EMA weights with the requested sampler (general_diffusion_trainer.py:351-402)."""
    torch.manual_seed(0)
    res, B = 16, 4
    model, tr = _trainer(res, True)
    seen = []

    def gen():
        g = torch.Generator().manual_seed(0)
        while True:
            b = torch.randint(0, 256, (B, res, res, 3), dtype=torch.uint8, generator=g)
            seen.append(b)
            yield {"image": b}

    class Metric:
        name = "mean_abs"

        @staticmethod
        def function(samples, batch):
            assert samples.shape == (4, res, res, 3) and samples.is_cuda
            return samples.abs().mean().item()

    tr.eval_metrics = [Metric]
    p0 = tr.state.params.flat.clone()
    st = tr.fit({"train": gen, "local_batch_size": B}, training_steps_per_epoch=3, epochs=2, val_steps_per_epoch=1,
                sampler_class=EulerSampler,
                sampling_noise_schedule=KarrasVENoiseScheduler(1, sigma_max=80, rho=7, sigma_data=0.5),
                verbose=False, val_diffusion_steps=3)
    assert st.step == 6 and tr.latest_step == 6 and not torch.equal(st.params.flat, p0)
    assert tr.last_val_samples is not None and torch.isfinite(tr.last_val_samples).all()
    assert tr.last_val_samples.abs().max() <= 1.0
    assert "val/mean_abs" in tr.best_val_metrics and tr.best_loss < 1e9
    assert len(seen) >= 6
    assert tr.best_state.step == 6 or tr.best_state.step == 3


def test_device_prefetcher_yields_device_batches_in_order():
    batches = [{"image": torch.full((2, 4, 4, 3), i, dtype=torch.uint8), "tag": i} for i in range(5)]
    out = list(DevicePrefetcher(iter(batches), "image", dev, depth=2))
    assert [b["tag"] for b in out] == list(range(5))
    for i, b in enumerate(out):
        assert b["image"].is_cuda and b["image"].dtype == torch.uint8 and int(b["image"][0, 0, 0, 0]) == i


@pytest.mark.parametrize("graph", [False, True])
def test_train_step_parameters_and_ema_vs_oracle(graph, monkeypatch):
    """compare the EMA and the MAGNITUDE of the parameter update with the oracle, not only
    the update direction."""
    torch.manual_seed(0)
    res, B = 16, 4
    model, tr = _trainer(res, graph)
    P = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in tr.state.params.named.items()}
    p0 = {k: v.detach().clone() for k, v in P.items()}
    ema = {k: v.detach().clone() for k, v in P.items()}
    freqs = model._fourier_freqs(dev).cpu()
    img = torch.randint(0, 256, (B, res, res, 3), dtype=torch.uint8)
    step = tr._define_train_step(B)
    opt = {}
    for it in range(3):
        noise, t = torch.randn(B, res, res, 3), torch.randn(B)
        monkeypatch.setattr(utils, "device_normal", lambda key, shape, device, dtype=torch.float32, _n=noise, _t=t:
                            (_t if len(shape) == 1 else _n).to(device))
        tr.state, loss, tr.rngstate = step(tr.state, tr.rngstate, {"image": img.clone()}, 0)
        want = train_ref.edm_train_step(P, opt, img, noise, t, freqs, lr=1e-3, wd=1e-4, ema=ema, step=it + 1)
        assert abs(loss.item() - want.item()) / want.item() < 2e-2
    n_got = n_ref = dot = 0.0
    e_num = e_den = 0.0
    for k in P:
        dg = tr.state.params.named[k].cpu() - p0[k]
        dr = P[k].detach() - p0[k]
        n_got += dg.pow(2).sum().item(); n_ref += dr.pow(2).sum().item(); dot += (dg * dr).sum().item()
        eg = tr.state.ema_params.named[k].cpu() - p0[k]
        er = ema[k] - p0[k]
        e_num += (eg - er).pow(2).sum().item(); e_den += er.pow(2).sum().item()
    assert dot / (n_got * n_ref) ** 0.5 > 0.9
    assert abs((n_got / n_ref) ** 0.5 - 1) < 0.05            # update magnitude within 5 %
    assert (e_num / e_den) ** 0.5 < 0.45                     # EMA displacement tracks the oracle's
    full = tr.state.ema_params
    for k in [k for k in P if k.endswith("kernel")][:6]:     # (biases start at 0: their EMA IS the update)
        assert rel(full.named[k], ema[k]) < 1e-3             # absolute EMA values: tight
