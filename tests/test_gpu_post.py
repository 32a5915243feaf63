"""GPU parity tests of the consumers right after the density hot path, through the wrapper API
(-> torch.ops.flowchain.* -> C-ABI):
  * denormalize (fc_to_trajectories)   vs the UNMODIFIED TP_metrics.denormalize outputs (tests/golden/post_metrics.npz);
  * predict_best_of_n (fc_best_of_n)   vs the reference's displacement_error / evaluate_helper numbers in the same fixture
                                        and vs the oracle at N = 2000 with NaN candidates;
  * prob_to_grid (fc_grid_density_maps) vs the fp64 oracle on an injected noise tape: position-space density maps on a
                                        metric lattice + the exact log p of the ground truth.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, load_golden
from oracle import flowchain_oracle as orc
from flowchain_iccv2023_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def model():
    from flowchain_iccv2023_b200.model import fastpredNF_TP
    spec, params, g = load_golden("default_eth")
    m = fastpredNF_TP(encoder=None, pred_len=spec.pred_len, n_blocks=spec.n_blocks, n_hidden=spec.n_hidden,
                      hidden_size=spec.hidden_size, cond_size=spec.cond_size)
    m.flow.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
    return m.cuda().eval(), spec, params, g


@pytest.fixture(scope="module")
def post():
    return dict(np.load(os.path.join(GOLDEN_DIR, "post_metrics.npz")))


@pytest.mark.parametrize("tag,state", [("v", "state_v"), ("p", "state_p")])
def test_denormalize_matches_reference_metrics(model, post, tag, state):
    m, spec, p, _ = model
    g = post
    m.pred_state = state
    std, dt, obs, gt_st, pred_st = g["std"], g["dt"], g["obs"], g["gt_st"], g["pred_st"]
    gt_un = gt_st * std if tag == "v" else gt_st * std + obs[:, -1:, 0:2]
    for j in range(pred_st.shape[0]):
        dd = {"obs": T(obs), "gt": T(gt_un), "gt_st": T(gt_st), "dt": T(dt), ("pred_st", 0): T(pred_st[j])}
        dd = m.denormalize(dd, std)
        np.testing.assert_allclose(dd[("pred", 0)].cpu().numpy(), g[f"pred_{tag}"][j], atol=2e-6, rtol=1e-6)
        np.testing.assert_allclose(dd["gt"].cpu().numpy(), g[f"gt_pos_{tag}"], atol=2e-6, rtol=1e-6)
    # sample clouds, batched over the three single-agent dictionaries the reference had to process one by one
    dd = {"obs": T(obs[:3]), "gt": T(gt_un[:3]), "gt_st": T(gt_st[:3]), "dt": T(dt[:3]), ("pred_st", 0): T(pred_st[0, :3]),
          ("prob_st", 0): T(g["prob_st_0"]), ("prob_st", 3): T(g["prob_st_3"])}
    dd = m.denormalize(dd, std)
    np.testing.assert_allclose(dd[("prob", 0)].cpu().numpy(), g[f"prob_0_{tag}"], atol=2e-6, rtol=2e-6)
    np.testing.assert_allclose(dd[("prob", 3)].cpu().numpy(), g[f"prob_3_{tag}"], atol=2e-6, rtol=2e-6)
    m.pred_state = "state_v"


@pytest.mark.parametrize("tag,mode", [("v", 1), ("p", 0)])
def test_best_of_n_matches_reference_metrics(model, post, tag, mode):
    m, spec, p, _ = model
    g = post
    h = m.flow.engine()
    cand = g["pred_st"].transpose(1, 0, 2, 3)                      # (B, N, T, 2)
    ade, fde, ia, jf = torch.ops.flowchain.best_of_n(h.id, T(cand), T(g[f"gt_pos_{tag}"]), T(g["obs"][:, -1, 0:2]),
                                                    T(g["dt"]) if mode == 1 else None, None, float(g["std"][0]), float(g["std"][1]), mode)
    np.testing.assert_allclose(ade.cpu().numpy(), g[f"ade_trials_{tag}"].min(axis=0), atol=2e-6, rtol=2e-6)
    np.testing.assert_allclose(fde.cpu().numpy(), g[f"fde_trials_{tag}"].min(axis=0), atol=2e-6, rtol=2e-6)
    np.testing.assert_array_equal(ia.cpu().numpy(), g[f"ade_trials_{tag}"].argmin(axis=0))
    np.testing.assert_array_equal(jf.cpu().numpy(), g[f"fde_trials_{tag}"].argmin(axis=0))
    np.testing.assert_allclose(ade.sum().item(), g[f"ade_sum_{tag}"], rtol=1e-5)     # TP_metrics.__call__'s "ade" / "fde"
    np.testing.assert_allclose(fde.sum().item(), g[f"fde_sum_{tag}"], rtol=1e-5)


def test_best_of_n_large_with_nan_vs_oracle(model):
    m, spec, p, _ = model
    h = m.flow.engine()
    A, N, Tn = 130, 2000, spec.pred_len
    cand = (0.4 * synth.normal(1, "cand", (A, N, Tn, 2))).astype(np.float32)
    cand[3, 7, 2, 1] = np.nan
    cand[5, :, 0, 0] = np.nan                                       # a whole agent needs the fallback
    gt = (synth.normal(2, "gt", (A, Tn, 2))).astype(np.float32)
    last = synth.normal(3, "last", (A, 2)).astype(np.float32)
    dt = np.full((A,), 0.4, np.float32)
    fb = (0.1 * synth.normal(4, "fb", (A, 2))).astype(np.float32)
    std = np.array([1.1, 0.9], np.float32)
    ade, fde, ia, jf = torch.ops.flowchain.best_of_n(h.id, T(cand), T(gt), T(last), T(dt), T(fb), float(std[0]), float(std[1]), 1)
    ra, rf, ri, rj = orc.best_of_n(cand, gt, last, dt, fb, std, 1, np.float64)
    np.testing.assert_allclose(ade.cpu().numpy(), ra, atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(fde.cpu().numpy(), rf, atol=1e-5, rtol=1e-5)
    assert (ia.cpu().numpy() == ri).mean() > 0.99 and (jf.cpu().numpy() == rj).mean() > 0.99    # fp32 near-ties may flip


def test_predict_best_of_n_end_to_end(model):
    """Wrapper: one sampler launch with S = n_trial on an injected tape, then the on-device reduction; the candidates are
    the sampler's trajectories and the minima equal the oracle's on exactly those candidates."""
    m, spec, p, g = model
    B, N, Tn = 4, 20, spec.pred_len
    inp = synth.make_inputs(spec, B, seed=11)
    m.encoder = lambda d: T(inp["cond"])
    obs_st = torch.zeros(B, 8, 6, device="cuda")
    obs_st[:, -1, 2:4] = T(inp["base_pos"])
    obs = synth.normal(12, "obs", (B, 8, 6)).astype(np.float32)
    std = np.array([1.7, 2.3], np.float32)
    dt = np.full((B,), 0.4, np.float32)
    gt = (obs[:, -1:, 0:2] + 0.5 * synth.normal(13, "gtp", (B, Tn, 2))).astype(np.float32)
    z0 = synth.normal(14, "z0", (B * N, 2)).astype(np.float32)
    eps = synth.make_eps(spec, B * N, 15)
    dd = {"obs_st": obs_st, "obs": T(obs), "gt": T(gt), "dt": T(dt)}
    res = m.predict_best_of_n(dd, std, n_trial=N, z0=T(z0), eps=T(eps))
    cand = res["candidates_st"].cpu().numpy()
    assert cand.shape == (B, N, Tn, 2)
    rs, _, _, _ = orc.sample_with_log_prob(p, np.repeat(inp["base_pos"], N, 0), np.repeat(inp["cond"], N, 0), z0, eps, np.float64)
    np.testing.assert_allclose(cand.reshape(rs.shape), rs, atol=2e-4, rtol=1e-4)
    ra, rf, ri, rj = orc.best_of_n(cand, gt, obs[:, -1, 0:2], dt, obs_st[:, 0, 2:4].cpu().numpy(), std, 1, np.float64)
    np.testing.assert_allclose(res["ade_per_agent"].cpu().numpy(), ra, atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(res["fde_per_agent"].cpu().numpy(), rf, atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(res["ade"].item(), ra.sum(), rtol=1e-5)
    assert res["nsample"] == B


@pytest.mark.parametrize("prec,atol,rtol", [(_lib.FC_PREC_FP32, 1e-4, 3e-6), (_lib.FC_PREC_F16X3, 5e-3, 3e-6)])
def test_prob_to_grid_vs_oracle(model, prec, atol, rtol):
    """Position-space density maps on a metric lattice (what TP_visualizer.prob_to_grid hands the plots and the EMD /
    log-prob metrics) and the exact log p of the ground truth, against the fp64 oracle on an injected tape."""
    m, spec, p, g = model
    B, Tn = 3, spec.pred_len
    inp = synth.make_inputs(spec, B, seed=21)
    m.encoder = lambda d: T(inp["cond"])
    obs_st = torch.zeros(B, 8, 6, device="cuda")
    obs_st[:, -1, 2:4] = T(inp["base_pos"])
    obs = synth.normal(22, "obs", (B, 8, 6)).astype(np.float32)
    std = np.array([1.7, 2.3], np.float32)
    dt = np.array([0.4, 0.4, 0.5], np.float32)
    gt_st = (inp["base_pos"][:, None, :] + 0.2 * synth.normal(23, "gt_st", (B, Tn, 2))).astype(np.float32)
    n = 12
    xs = np.linspace(-4.0, 4.0, n, dtype=np.float32) + obs[:, -1, 0].mean()
    ys = np.linspace(-4.0, 4.0, n, dtype=np.float32) + obs[:, -1, 1].mean()
    xx, yy = np.meshgrid(xs, ys)                                    # TP_visualizer.get_grid's layout
    eps = synth.make_eps(spec, B * n * n, 24)
    eps_gt = synth.make_eps(spec, B, 25)
    dd = {"obs_st": obs_st, "obs": T(obs), "gt_st": T(gt_st), "dt": T(dt)}
    m.flow.precision = prec
    dd = m.prob_to_grid(dd, xx, yy, std, as_prob=False, eps=T(eps), eps_gt=T(eps_gt))
    m.flow.precision = _lib.FC_PREC_FP32
    out = dd[("prob", 0)].cpu().numpy()
    assert out.shape == (B, n * n, Tn) and dd["grid"][0] is xx
    # oracle inputs: the inverse of TP_metrics.unstandardize / output_to_trajectories (TP_metrics.py:68-112)
    step = gt_st.astype(np.float64) * std * dt[:, None, None]
    origin = obs[:, -1, None, 0:2] + np.cumsum(step, axis=1) - step
    inv_scale = 1.0 / (std[None].astype(np.float64) * dt[:, None])
    log_jac = np.log(inv_scale).sum(-1)
    grid_xy = np.stack([xx.flatten(), yy.flatten()], -1)
    ref = orc.density_maps_metric(p, inp["base_pos"], inp["cond"], gt_st, grid_xy, origin, inv_scale, log_jac, eps, np.float64, as_prob=False)
    err = np.abs(out - ref)
    print(f"prob_to_grid log-density: max|err| {err.max():.3e} at max|ref| {np.abs(ref).max():.1f}")
    # the density's own bar + the fp32 rounding of the fused query transform z = (X - origin) * inv_scale (|X| ~ 4:
    # dz ~ 4e-7, amplified by |w - prev| / sigma^2 ~ 300 in the base term)
    assert np.all(err <= atol + 5e-4 + rtol * np.abs(ref))
    ref_gt = orc.log_prob(p, inp["base_pos"], gt_st, inp["cond"], eps_gt, np.float64) + log_jac[:, None]
    np.testing.assert_allclose(dd[("gt_traj_log_prob", 0)].cpu().numpy(), ref_gt, atol=atol, rtol=1e-5)
    # as_prob: the same numbers through the fused exp, a proper density (Riemann sum over a wide lattice ~ 1 at step 0)
    dd2 = m.prob_to_grid({"obs_st": obs_st, "obs": T(obs), "gt_st": T(gt_st), "dt": T(dt)}, xx, yy, std, eps=T(eps), eps_gt=T(eps_gt))
    np.testing.assert_allclose(dd2[("prob", 0)].cpu().numpy(), np.exp(ref), rtol=2e-2 if prec != _lib.FC_PREC_FP32 else 2e-3, atol=1e-30)
