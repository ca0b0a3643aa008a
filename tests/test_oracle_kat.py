"""Oracle (oracle/diffusion_ref.py) against the self-derived known answers of SURVEY.md Appendix C.
The reference has no tests or golden vectors; these values were computed from the reference
formulas in float64/float32 NumPy and pin the restatement's constants."""
import numpy as np

from oracle import diffusion_ref as R


def test_linear_schedule_tables():
    t = R.linear_tables(1000)
    acp = t["alpha_cumprod"]
    np.testing.assert_allclose(acp[[0, 1, 499, 500, 998, 999]],
                               [0.9999, 0.99978006, 0.07858724, 0.07779666, 4.1181964e-05, 4.0358325e-05], rtol=2e-6)
    np.testing.assert_allclose(t["sqrt_alpha_cumprod"][999], 0.0063528204, rtol=2e-6)
    np.testing.assert_allclose(t["sqrt_one_minus_alpha_cumprod"][0], 0.01000083, rtol=2e-5)
    assert t["posterior_variance"][0] == 0
    np.testing.assert_allclose(t["posterior_log_variance_clipped"][0], -46.0517, rtol=1e-6)
    np.testing.assert_allclose(t["posterior_variance"][[1, 999]], [5.453269e-05, 0.019999983], rtol=2e-5)
    np.testing.assert_allclose(t["p2_loss_weights"][[0, 500, 999]], [1.00016594e-04, 0.922203362, 0.999959588], rtol=2e-4)


def test_discrete_index_clamps_and_truncates():
    assert R.discrete_index([0, 999, 1000, 1500, 12.9, -1]).tolist() == [0, 999, 999, 999, 12, 999]


def test_karras_schedule():
    s = R.karras_sigma([1, 0.999, 0.5, 0.001, 0])
    np.testing.assert_allclose(s, [80, 79.5642604, 2.51521898, 0.00205014647, 0.002], rtol=3e-5)
    np.testing.assert_allclose(R.karras_model_time(s), [1.09550666, 1.09414125, 0.230589967, -1.54746101, -1.55365202],
                               rtol=3e-5)
    np.testing.assert_allclose(R.karras_weight(s), [4.00015625, 4.00015796, 4.15806699, 121907.175, 125002.0], rtol=2e-4)


def test_edm_sigma():
    np.testing.assert_allclose(R.edm_sigma([-2, 0, 1]), [0.0273237224, 0.301194212, 1.0], rtol=1e-6)


def test_get_steps_linear_bit_exact():
    want = {18: ([1000, 941, 882, 823, 764], [117, 58, 0]), 30: ([1000, 965, 931, 896, 862], [68, 34, 0]),
            50: ([1000, 979, 959, 938, 918], [40, 20, 0]), 100: ([1000, 989, 979, 969, 959], [20, 10, 0]),
            200: ([1000, 994, 989, 984, 979], [10, 5, 0])}
    for n, (head, tail) in want.items():
        s = R.get_steps_linear(1000, 0, n)
        assert s.dtype == np.int16 and len(s) == n
        assert s[:5].tolist() == head and s[-3:].tolist() == tail


def test_karras_preconditioning():
    c_in, c_out, c_skip = R.karras_coeffs([0.5], 0.5)
    np.testing.assert_allclose([c_in[0], c_skip[0], c_out[0]], [1.41421354, 0.49999999, 0.353553386], rtol=1e-6)


def test_euler_step_on_ve_schedule():
    rng = np.random.default_rng(0)
    x, x0 = rng.standard_normal((2, 4, 4, 3), dtype=np.float32), rng.standard_normal((2, 4, 4, 3), dtype=np.float32)
    cs, r = np.float32(3.0), np.float32(0.25)
    out = R.euler_step(x, x0, [1, 1], [cs, cs], [1, 1], [cs * r, cs * r])
    np.testing.assert_allclose(out, x0 + r * (x - x0), rtol=2e-5, atol=2e-6)


def test_oracle_ddpm_train_step_configs0():
    """BASELINE configs[0] on the CPU oracle: LinearNoiseSchedule + epsilon prediction.  At t = 0 the input is
    almost the clean image and the p2 weight (1 + acp/(1-acp))^-1 = 1 - acp = beta_0 = 1e-4 nearly cancels the loss."""
    import torch
    from oracle import train_ref, unet_ref  # noqa: F401
    from flaxdiff_b200 import utils
    from flaxdiff_b200.models.simple_unet import Unet
    torch.manual_seed(0)
    model = Unet(attention_configs=(None,) * 4)
    fp = model.init(utils.PRNGKey(4), device=torch.device("cpu"))
    P = {k: v.clone().requires_grad_(True) for k, v in fp.named.items()}
    freqs = model._fourier_freqs("cpu")
    img = torch.randint(0, 256, (2, 8, 8, 3), dtype=torch.uint8)
    noise = torch.randn(2, 8, 8, 3)
    before = {k: v.detach().clone() for k, v in P.items()}
    l_hi = train_ref.ddpm_train_step(P, {}, img, noise, torch.tensor([999, 500]), freqs, lr=1e-3)
    assert torch.isfinite(l_hi) and l_hi > 0
    assert any(not torch.equal(before[k], P[k].detach()) for k in P)          # AdamW moved the parameters
    l_lo = train_ref.ddpm_train_step(P, {}, img, noise, torch.tensor([0, 0]), freqs, lr=0.0, wd=0.0)
    assert l_lo < 1e-3 * l_hi
