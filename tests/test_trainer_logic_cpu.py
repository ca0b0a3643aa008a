"""Host logic of the trainer that needs no GPU: optimiser chain descriptions, bucket bounds of the
gradient exchange, the oracle's DynamicScale / lamb / clip restatements on hand-computable inputs."""
import numpy as np

from flaxdiff_b200.trainer import (Optimizer, adam, adamw, bucket_bounds, chain, clip_by_global_norm, lamb,
                                   warmup_cosine_decay_schedule)
from oracle import diffusion_ref as R


def test_optax_chain_descriptions():
    o = chain(clip_by_global_norm(1.5), adamw(1e-3, weight_decay=0.01))
    assert (o.kind, o.clip_norm, o.weight_decay, o.lr_at(0)) == ("adamw", 1.5, 0.01, 1e-3)
    assert adam(1e-3).weight_decay == 0.0 and adam(1e-3).kind == "adam"
    l = lamb(2e-3)
    assert (l.kind, l.eps, l.weight_decay) == ("lamb", 1e-6, 0.0)        # optax.lamb defaults
    sched = warmup_cosine_decay_schedule(0.0, 1.0, 10, 110, 0.1)
    assert Optimizer(sched).lr_at(5) == 0.5 and abs(Optimizer(sched).lr_at(110) - 0.1) < 1e-12


def test_bucket_bounds_cover_and_align():
    for total, n in ((32040704 + 64, 4), (1000, 4), (5000, 1), (8192, 16)):
        b = bucket_bounds(total, n)
        assert b[0][0] == 0 and b[-1][1] == total
        assert all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
        assert all(lo % 1024 == 0 for lo, _ in b) and all(hi > lo for lo, hi in b)
        assert len(b) <= n


def test_oracle_dynamic_scale_rule():
    s, f = 65536.0, 0
    s, f = R.dynamic_scale_update(s, f, True, growth_interval=2)
    assert (s, f) == (65536.0, 1)
    s, f = R.dynamic_scale_update(s, f, True, growth_interval=2)
    assert (s, f) == (65536.0, 2)
    s, f = R.dynamic_scale_update(s, f, True, growth_interval=2)       # fin_steps == interval -> grow
    assert (s, f) == (131072.0, 0)
    s, f = R.dynamic_scale_update(s, 1, False, growth_interval=2)      # non-finite -> halve, reset
    assert (s, f) == (65536.0, 0)


def test_oracle_clip_and_lamb_closed_forms():
    g = [np.array([3.0, 0.0], np.float32), np.array([4.0], np.float32)]          # ||g|| = 5
    out, nrm = R.clip_by_global_norm(g, 1.0)
    assert abs(nrm - 5) < 1e-6 and np.allclose(out[0], [0.6, 0]) and np.allclose(out[1], [0.8])
    out, _ = R.clip_by_global_norm(g, 10.0)
    assert np.array_equal(out[0], g[0])
    # first lamb step: u = sign(g) (bias-corrected m / sqrt(v)), trust ratio = ||p|| / ||u||
    p = [np.array([3.0, 4.0], np.float32)]
    gg = [np.array([0.5, -2.0], np.float32)]
    z = [np.zeros(2, np.float32)]
    p2, _, _ = R.lamb_step(p, gg, z, [z[0].copy()], 1, lr=0.1, eps=0.0)
    assert np.allclose(p2[0], p[0] - 0.1 * (5 / np.sqrt(2)) * np.array([1, -1]), atol=1e-5)
    # zero parameter norm -> ratio 1
    p2, _, _ = R.lamb_step([np.zeros(2, np.float32)], gg, z, [z[0].copy()], 1, lr=0.1, eps=0.0)
    assert np.allclose(p2[0], -0.1 * np.array([1, -1]), atol=1e-6)
