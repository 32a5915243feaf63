"""Training path (SURVEY §8b "Registration", §8f rank 4): with the flow in train() mode and grad enabled the
reference-API methods run plain differentiable PyTorch (flow.py) so that `fastpredNF_TP.update()` -- Adam on the mean
negative log-density, fastpredNF.py:186-202 -- keeps working.  CPU tests:
  * the PyTorch forwards against the numpy oracle (eval statistics, injected noise tape, fp64 1e-9);
  * train() semantics of the invertible BatchNorm (batch statistics, the reference's running update);
  * one whole update() step (loss, gradients, Adam, running statistics) against the UNMODIFIED reference classes when
    /root/reference is present in this container;
  * update() lowers the loss and invalidates the packed weights of the fused kernels.
"""
import copy
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from oracle import flowchain_oracle as orc
from flowchain_iccv2023_b200 import synth
from flowchain_iccv2023_b200.flow import fastpredNF_CIF_separate_cond
from flowchain_iccv2023_b200.model import fastpredNF_TP


def build(spec, params, dtype=torch.float64):
    f = fastpredNF_CIF_separate_cond(2, spec.n_blocks, spec.hidden_size, spec.n_hidden, spec.cond_size,
                                     flow_architecture="realNVP", pred_len=spec.pred_len)
    f.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
    return f.to(dtype)


def t64(a):
    return torch.from_numpy(np.asarray(a, dtype=np.float64))


@pytest.mark.parametrize("case", ["default_eth", "odd_small", "deep_blocks"])
def test_torch_path_matches_oracle(case):
    """Eval statistics + injected tape: every chain-level PyTorch method equals the fp64 oracle / golden outputs."""
    spec, p, g = load_golden(case)
    f = build(spec, p).eval()
    with torch.no_grad():
        out = f._log_prob_separate_torch(t64(g["base_pos"]), t64(g["x"]), t64(g["cond"]), t64(g["eps"])).numpy()
        np.testing.assert_allclose(out, g["log_prob_64"], atol=1e-9, rtol=1e-10)
        out = f._log_prob_sequential_torch(t64(g["base_pos"]), t64(g["x"]), t64(g["cond"]), None, t64(g["eps_seq"])).numpy()
        np.testing.assert_allclose(out, g["log_prob_seq_64"], atol=1e-9, rtol=1e-10)
        ns = min(3, spec.pred_len - 1)
        out = f._log_prob_sequential_torch(t64(g["base_pos"]), t64(g["x"]), t64(g["cond"]), ns, t64(g["eps_seq"])).numpy()
        np.testing.assert_allclose(out, g["log_prob_seq_n3_64"], atol=1e-9, rtol=1e-10)
        B, S, D = g["z0"].shape
        bp = t64(g["base_pos"][:B])[:, None].expand(-1, S, -1)
        cd = t64(g["cond"][:B])[:, None].expand(-1, S, -1, -1)
        seq, lp, ldjs, base = f._sample_with_log_prob_torch(bp, cd, t64(g["z0"]), t64(g["eps_sample"]).reshape(B, S, -1))
        np.testing.assert_allclose(seq.numpy(), g["sample_seq_64"], atol=1e-8, rtol=1e-8)
        np.testing.assert_allclose(ldjs.numpy(), g["sample_ldjs_64"], atol=1e-6, rtol=1e-8)
        np.testing.assert_allclose(lp.numpy(), g["sample_log_prob_64"], atol=1e-6, rtol=1e-8)
        np.testing.assert_allclose(base.numpy(), g["sample_base_64"], atol=1e-10)


def test_dispatch_train_mode_is_differentiable_torch():
    """train() + grad -> PyTorch (works on CPU, gradients reach every parameter); the same call in eval() refuses the CPU."""
    spec, p, g = load_golden("odd_small")
    f = build(spec, p, torch.float32).train()
    lp = f.log_prob(torch.from_numpy(g["base_pos"]), torch.from_numpy(g["x"]), torch.from_numpy(g["cond"]))
    assert lp.requires_grad and lp.shape == g["log_prob_64"].shape
    (-lp.mean()).backward()
    missing = [n for n, q in f.named_parameters() if q.grad is None]
    assert not missing, missing
    assert all(torch.isfinite(q.grad).all() for q in f.parameters())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        f.eval().log_prob(torch.from_numpy(g["base_pos"]), torch.from_numpy(g["x"]), torch.from_numpy(g["cond"]))


def test_batchnorm_train_semantics():
    """TFCondARFlow.py:286-315: batch mean / UNBIASED batch variance, running <- 0.9 running + 0.1 batch, inverse in
    train() mode reuses the statistics of the last forward."""
    from flowchain_iccv2023_b200.flow import BatchNorm
    bn = BatchNorm(2).double().train()
    with torch.no_grad():
        bn.log_gamma.copy_(torch.tensor([0.3, -0.2]))
        bn.beta.copy_(torch.tensor([0.1, 0.5]))
    x = torch.randn(64, 3, 2, dtype=torch.float64, generator=torch.Generator().manual_seed(0)) * 2 + 1
    y, ldj = bn(x)
    flat = x.reshape(-1, 2)
    mean, var = flat.mean(0), flat.var(0, unbiased=True)
    np.testing.assert_allclose(y.detach().numpy(), (torch.exp(bn.log_gamma) * (x - mean) / torch.sqrt(var + 1e-5) + bn.beta).detach().numpy(), atol=1e-12)
    np.testing.assert_allclose(ldj[0, 0].detach().numpy(), (bn.log_gamma - 0.5 * torch.log(var + 1e-5)).detach().numpy(), atol=1e-12)
    np.testing.assert_allclose(bn.running_mean.numpy(), 0.1 * mean.numpy(), atol=1e-12)
    np.testing.assert_allclose(bn.running_var.numpy(), 0.9 + 0.1 * var.numpy(), atol=1e-12)
    xb, ldjb = bn.inverse(y)
    np.testing.assert_allclose(xb.detach().numpy(), x.numpy(), atol=1e-10)
    np.testing.assert_allclose((ldj + ldjb).detach().numpy(), 0.0, atol=1e-12)
    bn.eval()
    y2, _ = bn(x)
    assert not torch.allclose(y2, y)          # eval uses the running statistics


def make_wrapper(spec, params, cond, lr=1e-3, wd=1e-6):
    m = fastpredNF_TP(encoder=lambda d: cond, pred_len=spec.pred_len, n_blocks=spec.n_blocks, n_hidden=spec.n_hidden,
                      hidden_size=spec.hidden_size, cond_size=spec.cond_size, lr=lr, weight_decay=wd)
    m.flow.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
    return m


def test_update_lowers_loss_and_invalidates_packed_weights():
    spec = synth.FlowSpec(pred_len=3, cond_size=4, hidden_size=16, n_hidden=1, n_blocks=2)
    params = synth.make_params(spec, 5)
    B = 32
    inp = synth.make_inputs(spec, B, seed=6)
    cond = torch.from_numpy(inp["cond"])
    m = make_wrapper(spec, params, cond, lr=3e-3).train()
    obs = torch.zeros(B, 8, 6)
    obs[:, -1, 2:4] = torch.from_numpy(inp["base_pos"])
    gt = torch.from_numpy((inp["base_pos"][:, None, :] + 0.1 * synth.normal(7, "gt", (B, 3, 2))).astype(np.float32))
    torch.manual_seed(0)
    m.flow._packed_version = 123            # pretend the kernels hold a packed copy
    before = copy.deepcopy(m.flow.state_dict())
    losses = [m.update({"obs_st": obs, "gt_st": gt.clone()})["loss"] for _ in range(60)]
    assert m.flow._packed_version is None   # re-pack forced on the next inference call
    assert np.mean(losses[-10:]) < np.mean(losses[:10]) - 0.5, (losses[:3], losses[-3:])
    after = m.flow.state_dict()
    assert any(not torch.equal(before[k], after[k]) for k in before if "weight" in k)
    assert not torch.equal(before["net.0.bijection.1.running_mean"], after["net.0.bijection.1.running_mean"])
    assert m.optimizers == [m.optimizer] and len(m.optimizer.param_groups) == 2
    assert m.optimizer.param_groups[1]["weight_decay"] == 0.0       # biases are not decayed (fastpredNF.py:72-86)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present")
def test_update_step_matches_unmodified_reference():
    """One training step, train() mode, same weights / batch / noise tape: loss, every gradient-updated parameter and the
    BatchNorm running statistics equal the UNMODIFIED reference's (fastpredNF.py:186-202 via object.__new__)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    ref = make_golden.import_reference()
    spec = synth.FlowSpec(pred_len=4, cond_size=6, hidden_size=32, n_hidden=2, n_blocks=3)
    params = synth.make_params(spec, 11)
    B = 16
    inp = synth.make_inputs(spec, B, seed=12)
    cond = torch.from_numpy(inp["cond"])
    obs = torch.zeros(B, 8, 6)
    obs[:, -1, 2:4] = torch.from_numpy(inp["base_pos"])
    gt = torch.from_numpy((inp["base_pos"][:, None, :] + 0.2 * synth.normal(13, "gt", (B, spec.pred_len, 2))).astype(np.float32))
    eps = torch.from_numpy(synth.make_eps(spec, B, 14))

    rflow = make_golden.build_flow(ref, spec, 11).train()
    rmodel = object.__new__(ref.fastpredNF_TP)
    torch.nn.Module.__init__(rmodel)
    rmodel.flow, rmodel.pred_len, rmodel.pred_state, rmodel.dequantize = rflow, spec.pred_len, "state_v", False
    rmodel.encoder = lambda d: cond
    decay = [n for n, _ in rmodel.named_parameters() if "bias" not in n]
    rmodel.optimizer = torch.optim.Adam(
        [{"params": [q for n, q in rmodel.named_parameters() if n in decay], "weight_decay": 1e-6},
         {"params": [q for n, q in rmodel.named_parameters() if n not in decay], "weight_decay": 0.0}], 1e-3)
    with make_golden.NoiseTape(make_golden.sep_tape(spec, eps)):
        rloss = rmodel.update({"obs_st": obs, "gt_st": gt.clone()})["loss"]

    m = make_wrapper(spec, params, cond, lr=1e-3, wd=1e-6).train()
    real_randn_like = torch.randn_like
    tape = make_golden.sep_tape(spec, eps)
    it = iter(tape)
    torch.randn_like = lambda t, *a, **k: next(it).to(t.dtype)      # our modules draw randn_like in the same order
    try:
        loss = m.update({"obs_st": obs, "gt_st": gt.clone()})["loss"]
    finally:
        torch.randn_like = real_randn_like
    assert abs(loss - rloss) <= 1e-5 * max(1.0, abs(rloss)), (loss, rloss)
    rs, os_ = rflow.state_dict(), m.flow.state_dict()
    assert list(rs.keys()) == list(os_.keys())
    rgrad = {n: q.grad for n, q in rflow.named_parameters()}
    ograd = {n: q.grad for n, q in m.flow.named_parameters()}
    scale = max(float(v.abs().max()) for v in rgrad.values())
    for k in rs:
        if k in rgrad:      # gradients (left in .grad by both update() calls) agree to fp32 rounding of the backward pass
            np.testing.assert_allclose(ograd[k].numpy(), rgrad[k].numpy(), atol=2e-6 * scale, rtol=1e-4, err_msg="grad " + k)
            # Adam's first step is lr * g / (|g| + 1e-8): where |g| is at the rounding floor the step itself is noise,
            # so the updated parameters are compared where the gradient is resolved
            ok = (rgrad[k].abs() > 1e-4 * scale).numpy()
            np.testing.assert_allclose(os_[k].numpy()[ok], rs[k].numpy()[ok], atol=2e-6, rtol=1e-5, err_msg=k)
        else:               # buffers: masks, BatchNorm running statistics after the train-mode forward
            np.testing.assert_allclose(os_[k].numpy(), rs[k].numpy(), atol=1e-6, rtol=1e-5, err_msg=k)
