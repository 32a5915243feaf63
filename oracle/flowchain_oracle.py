"""CPU oracle for the FlowChain analytical-density hot path.  TEST INFRASTRUCTURE ONLY.

This is a numpy restatement of the reference's algorithm (PyTorch eager, fp32) for the
path SURVEY.md §8 names.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product path (flowchain_iccv2023_b200) never does and fails loudly without its CUDA library.

Parity status: PINNED.  The reference has no tests or golden vectors of its own
(SURVEY §4), so the pin is the reference code itself: tests/golden/make_golden.py
imports the unmodified reference classes from /root/reference, injects a recorded
noise tape, and commits inputs + fp32/fp64 reference outputs under tests/golden/;
tests/test_oracle_golden.py checks every function below against them.

Conventions: ``p`` is a dict of numpy arrays keyed like the reference state_dict of
``fastpredNF_CIF_separate_cond``; ``dtype`` selects float32 (what the reference computes
in) or float64 (ground truth for tolerances).  All noise is INJECTED (the reference draws
``torch.randn_like`` -- log_prob is stochastic).

Noise-tape layouts (rows N = agents x points):
  separate / inverse : eps (N, sum_i C_i); step i owns columns [off_i, off_i + C_i).
  sequential         : eps (N, sum_k (T-k) C_k); block k is the (N, T-k, C_k) draw the reference
                       makes at loop step k (result index j sits at j-k inside the block).
  inverse            : additionally z0 (N, 2), the base draw u0 = base_pos + std0 * z0.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np

LOG_2PI = math.log(2.0 * math.pi)
BN_EPS = 1e-5  # TFCondARFlow.py:275


class Spec:
    """Shape of a chain inferred from the parameter dictionary."""

    def __init__(self, p: Dict[str, np.ndarray]):
        self.T, self.D = p["var"].shape
        self.n_blocks = 0
        while f"net.0.bijection.{2 * self.n_blocks}.mask" in p:
            self.n_blocks += 1
        n_lin = 0
        while f"net.0.bijection.0.s_net.{2 * n_lin}.weight" in p:
            n_lin += 1
        self.n_hidden = n_lin - 2
        w0 = p["net.0.bijection.0.s_net.0.weight"]
        self.H = w0.shape[0]
        self.C0 = w0.shape[1] - self.D

    def C(self, i: int) -> int:
        return self.C0 + i * self.D

    def eps_offsets(self) -> List[int]:
        off, o = [], 0
        for i in range(self.T):
            off.append(o)
            o += self.C(i)
        off.append(o)
        return off

    def seq_eps_offsets(self) -> List[int]:
        off, o = [], 0
        for k in range(self.T):
            off.append(o)
            o += (self.T - k) * self.C(k)
        off.append(o)
        return off


def cast_params(p: Dict[str, np.ndarray], dtype) -> Dict[str, np.ndarray]:
    return {k: np.asarray(v, dtype=dtype) for k, v in p.items()}


# ----------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------

def _linear(p, key, x):
    """nn.Linear: x @ W^T + b."""
    return x @ p[key + ".weight"].T + p[key + ".bias"]


def _coupling_nets(p, pre: str, n_lin: int, inp: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """s_net (Tanh) and t_net (ReLU) of one LinearMaskedCoupling (TFCondARFlow.py:341-358)."""
    s = inp
    t = inp
    for l in range(n_lin):
        s = _linear(p, f"{pre}.s_net.{2 * l}", s)
        t = _linear(p, f"{pre}.t_net.{2 * l}", t)
        if l < n_lin - 1:
            s = np.tanh(s)
            t = np.maximum(t, 0)
    return s, t


def coupling_forward(p, pre, n_lin, x, y):
    """LinearMaskedCoupling.forward (TFCondARFlow.py:360-381): returns (u, per-dim ldj)."""
    mask = p[pre + ".mask"]
    mx = x * mask
    s, t = _coupling_nets(p, pre, n_lin, np.concatenate([y, mx], axis=-1))
    t = t * (1 - mask)
    log_s = np.tanh(s) * (1 - mask)
    u = x * np.exp(log_s) + t
    return u, log_s


def coupling_inverse(p, pre, n_lin, u, y):
    """LinearMaskedCoupling.inverse (TFCondARFlow.py:383-402)."""
    mask = p[pre + ".mask"]
    mu = u * mask
    s, t = _coupling_nets(p, pre, n_lin, np.concatenate([y, mu], axis=-1))
    t = t * (1 - mask)
    log_s = np.tanh(s) * (1 - mask)
    x = (u - t) * np.exp(-log_s)
    return x, -log_s


def batchnorm_forward(p, pre, x):
    """BatchNorm.forward, eval branch (TFCondARFlow.py:303-315)."""
    mean, var = p[pre + ".running_mean"], p[pre + ".running_var"]
    x_hat = (x - mean) / np.sqrt(var + x.dtype.type(BN_EPS))
    y = np.exp(p[pre + ".log_gamma"]) * x_hat + p[pre + ".beta"]
    ldj = p[pre + ".log_gamma"] - x.dtype.type(0.5) * np.log(var + x.dtype.type(BN_EPS))
    return y, np.broadcast_to(ldj, x.shape)


def batchnorm_inverse(p, pre, y):
    """BatchNorm.inverse, eval branch (TFCondARFlow.py:317-330)."""
    mean, var = p[pre + ".running_mean"], p[pre + ".running_var"]
    x_hat = (y - p[pre + ".beta"]) * np.exp(-p[pre + ".log_gamma"])
    x = x_hat * np.sqrt(var + y.dtype.type(BN_EPS)) + mean
    ldj = y.dtype.type(0.5) * np.log(var + y.dtype.type(BN_EPS)) - p[pre + ".log_gamma"]
    return x, np.broadcast_to(ldj, y.shape)


def bijection_forward(p, spec: Spec, i: int, x, y):
    """FlowSequential.forward over [coupling, BN] x n_blocks (TFCondARFlow.py:257-262)."""
    total = 0
    for b in range(spec.n_blocks):
        x, l = coupling_forward(p, f"net.{i}.bijection.{2 * b}", spec.n_hidden + 2, x, y)
        total = total + l
        x, l = batchnorm_forward(p, f"net.{i}.bijection.{2 * b + 1}", x)
        total = total + l
    return x, np.sum(total, axis=-1)


def bijection_inverse(p, spec: Spec, i: int, u, y):
    """FlowSequential.inverse: modules reversed (TFCondARFlow.py:264-269)."""
    total = 0
    for b in range(spec.n_blocks - 1, -1, -1):
        u, l = batchnorm_inverse(p, f"net.{i}.bijection.{2 * b + 1}", u)
        total = total + l
        u, l = coupling_inverse(p, f"net.{i}.bijection.{2 * b}", spec.n_hidden + 2, u, y)
        total = total + l
    return u, np.sum(total, axis=-1)


def density_params(p, pre, cond_inputs):
    """DiagonalGaussianConditionalDensity._means_and_log_stddevs (fastpredNF.py:709-710, 736-753)."""
    out = []
    for net in ("shift_net", "log_scale_net"):
        h = cond_inputs
        for l in range(3):
            h = _linear(p, f"{pre}.{net}.{2 * l}", h)
            if l < 2:
                h = np.maximum(h, 0)
        out.append(h)
    return out[0], out[1]


def diag_gauss_log_prob(w, means, log_std):
    """diagonal_gaussian_log_prob (fastpredNF.py:713-724)."""
    var = np.exp(log_std) ** 2
    dim = w.shape[-1]
    const = w.dtype.type(-0.5 * dim * LOG_2PI)
    return const + (-np.sum(log_std, axis=-1)) + w.dtype.type(-0.5) * np.sum((w - means) ** 2 / var, axis=-1)


def diag_gauss_sample(means, log_std, eps):
    """diagonal_gaussian_sample with the noise injected (fastpredNF.py:727-733)."""
    samples = np.exp(log_std) * eps + means
    return samples, diag_gauss_log_prob(samples, means, log_std)


def cif_forward(p, spec: Spec, i: int, z, y, eps):
    """CIF_step.forward (fastpredNF.py:769-774): K = 1 auxiliary sample."""
    mu, ls = density_params(p, f"net.{i}.r_u_given_z", np.concatenate([z, y], axis=-1))
    u, log_r = diag_gauss_sample(mu, ls, eps)
    w, ldj = bijection_forward(p, spec, i, z, u)
    mu, ls = density_params(p, f"net.{i}.q_u_given_w", np.concatenate([w, y], axis=-1))
    log_q = diag_gauss_log_prob(u, mu, ls)
    return w, ldj + log_q - log_r


def cif_inverse(p, spec: Spec, i: int, w, y, eps):
    """CIF_step.inverse (fastpredNF.py:776-781)."""
    mu, ls = density_params(p, f"net.{i}.q_u_given_w", np.concatenate([w, y], axis=-1))
    u, log_q = diag_gauss_sample(mu, ls, eps)
    z, ldj = bijection_inverse(p, spec, i, w, u)
    mu, ls = density_params(p, f"net.{i}.r_u_given_z", np.concatenate([z, y], axis=-1))
    log_r = diag_gauss_log_prob(u, mu, ls)
    return z, ldj + log_q - log_r


def normal_log_prob(value, loc, scale):
    """torch.distributions.Normal.log_prob: -(v-loc)^2/(2 var) - log(scale) - log(sqrt(2 pi))."""
    var = scale ** 2
    return -((value - loc) ** 2) / (2 * var) - np.log(scale) - value.dtype.type(math.log(math.sqrt(2 * math.pi)))


def base_std(p, step):
    """fastpredNF.base_dist's scale: clamp(var[step], min=1e-4) (fastpredNF.py:401-402)."""
    return np.maximum(p["var"][step], p["var"].dtype.type(1e-4))


# ----------------------------------------------------------------------------------------------
# chain-level entry points
# ----------------------------------------------------------------------------------------------

def forward_separate(p, seq, cond, eps):
    """fastpredNF_separate_cond.forward_separate (fastpredNF.py:571-584)."""
    spec = Spec(p)
    N, T = seq.shape[0], seq.shape[1]
    off = spec.eps_offsets()
    seq_cond = seq.reshape(N, T * spec.D)
    us, ldjs = [], []
    for step in range(T):
        y = np.concatenate([cond[:, step], seq_cond[:, :step * spec.D]], axis=-1)
        u, l = cif_forward(p, spec, step, seq[:, step], y, eps[:, off[step]:off[step + 1]])
        us.append(u)
        ldjs.append(l)
    return np.stack(us, axis=1), np.stack(ldjs, axis=1)


def log_prob(p, base_pos, x, cond, eps, dtype=np.float32):
    """fastpredNF.log_prob -> log_prob_separate (fastpredNF.py:450-452, 460-465): (N, T)."""
    p = cast_params(p, dtype)
    base_pos, x, cond, eps = (np.asarray(a, dtype=dtype) for a in (base_pos, x, cond, eps))
    u, ldj = forward_separate(p, x, cond, eps)
    prev = np.concatenate([base_pos[:, None], x[:, :-1]], axis=1)
    std = base_std(p, list(range(prev.shape[1])))
    base = np.sum(normal_log_prob(u, prev, std), axis=-1)
    return base + ldj


def forward_sequential(p, seq, cond, eps, n_step: Optional[int] = None):
    """fastpredNF_separate_cond.forward_sequential (fastpredNF.py:548-569)."""
    spec = Spec(p)
    N, T = seq.shape[0], seq.shape[1]
    off = spec.seq_eps_offsets()
    x = seq[:, -1:].copy()
    seq_cond = seq.reshape(N, T * spec.D)[:, None]
    cum = 0
    start = T - 1
    stop = -1 if n_step is None else start - n_step
    for step in range(start, stop, -1):
        C = spec.C(step)
        pre = np.broadcast_to(seq_cond[..., :step * spec.D], (N, T - step, step * spec.D))
        y = np.concatenate([cond[:, step:], pre], axis=-1)
        e = eps[:, off[step]:off[step + 1]].reshape(N, T - step, C)
        x, l = cif_forward(p, spec, step, x, y, e)
        cum = cum + l
        if step != 0:
            x = np.concatenate([seq[:, step - 1:step], x], axis=1)
            cum = np.concatenate([np.zeros_like(cum[:, 0:1]), cum], axis=1)
    cum = cum / np.cumsum(np.ones_like(cum), axis=1)
    return x, cum


def log_prob_sequential(p, base_pos, x, cond, eps, n_step: Optional[int] = None, dtype=np.float32):
    """fastpredNF.log_prob_sequential (fastpredNF.py:454-458): (N, T) or (N, n_step+1)."""
    p = cast_params(p, dtype)
    base_pos, x, cond, eps = (np.asarray(a, dtype=dtype) for a in (base_pos, x, cond, eps))
    u, cum = forward_sequential(p, x, cond, eps, n_step)
    base = np.sum(normal_log_prob(u, base_pos[:, None], base_std(p, 0)), axis=-1)
    return base + cum


def inverse(p, u, cond, eps):
    """fastpredNF_separate_cond.inverse (fastpredNF.py:586-607) on flattened rows:
    u (N,2), cond (N,T,C0), eps (N, sum C_i) -> seq (N,T,2), ldj cumulative mean (N,T), ldjs (N,T)."""
    spec = Spec(p)
    off = spec.eps_offsets()
    N = u.shape[0]
    seq, ldjs = [], []
    for step in range(cond.shape[1]):
        if step != 0:
            y = np.concatenate([cond[:, step], np.stack(seq, axis=1).reshape(N, -1)], axis=-1)
        else:
            y = cond[:, step]
        u, l = cif_inverse(p, spec, step, u, y, eps[:, off[step]:off[step + 1]])
        seq.append(u)
        ldjs.append(l)
    seq = np.stack(seq, axis=1)
    ldjs = np.stack(ldjs, axis=1)
    cum = np.cumsum(ldjs, axis=1)
    cum = cum / np.cumsum(np.ones_like(cum), axis=1)
    return seq, cum, ldjs


def sample_with_log_prob(p, base_pos, cond, z0, eps, dtype=np.float32):
    """fastpredNF.sample_with_log_prob (fastpredNF.py:472-477) on flattened rows N = B*S.
    Returns (seq (N,T,2), log_prob (N,T), ldjs (N,T), base_log_prob (N,1))."""
    p = cast_params(p, dtype)
    base_pos, cond, z0, eps = (np.asarray(a, dtype=dtype) for a in (base_pos, cond, z0, eps))
    std0 = base_std(p, 0)
    u0 = z0 * std0 + base_pos          # Normal.sample(): normal_() * std + mean
    seq, cum, ldjs = inverse(p, u0, cond, eps)
    base = np.sum(normal_log_prob(u0, base_pos, std0), axis=-1)[..., None]
    return seq, base + cum, ldjs, base


def predict_from_new_obs(p, prob_st_0, ldjs, normalizer, new_pos, time_step: int, dtype=np.float32):
    """fastpredNF_TP.predict_from_new_obs for ONE agent (B = 1; fastpredNF.py:142-164).
    prob_st_0 (S,T,3), ldjs (S,T), normalizer scalar = sum_s exp(base_lp), new_pos (2,) = gt_st[t-1].
    Returns prob_st_t (S, T-t, 3)."""
    p = cast_params(p, dtype)
    prob_st_0, ldjs, new_pos = (np.asarray(a, dtype=dtype) for a in (prob_st_0, ldjs, new_pos))
    normalizer = np.asarray(normalizer, dtype=dtype)
    t = time_step
    assert 0 < t < p["var"].shape[0]
    out = prob_st_0[:, t:].copy()
    lp = np.sum(normal_log_prob(prob_st_0[:, t - 1, :2], new_pos, base_std(p, t)), axis=-1)
    bp = np.exp(lp)
    with np.errstate(divide="ignore", invalid="ignore"):
        lp = np.log(bp / bp.sum() * normalizer)
    out[..., -1] = lp[:, None] + ldjs[:, t:]
    return out


def predict_from_new_obs_batched(p, prob_st_0, ldjs, normalizer, new_pos, time_step, dtype=np.float32):
    """vmap of the B = 1 reference call over agents: prob_st_0 (A,S,T,3)."""
    return np.stack([predict_from_new_obs(p, prob_st_0[a], ldjs[a], normalizer[a], new_pos[a], time_step, dtype)
                     for a in range(prob_st_0.shape[0])])


# ----------------------------------------------------------------------------------------------
# dense-grid evaluation: rows = agents x grid points
# ----------------------------------------------------------------------------------------------

def grid_rows(base_pos, cond, grid):
    """Expand per-agent inputs to rows n = a*G + g for prefix mode 'self' (x = g repeated over T)."""
    A, T = cond.shape[0], cond.shape[1]
    G = grid.shape[0]
    bp = np.repeat(base_pos, G, axis=0)
    cd = np.repeat(cond, G, axis=0)
    x = np.broadcast_to(np.tile(grid, (A, 1))[:, None, :], (A * G, T, grid.shape[1])).copy()
    return bp, x, cd


def grid_log_prob_self(p, base_pos, cond, grid, eps, dtype=np.float32, sequential=False, n_step=None):
    """a13-(i): ONE reference call flow.log_prob(base_pos, g.expand(T), cond) with rows = A*G. -> (A,G,T)."""
    bp, x, cd = grid_rows(base_pos, cond, grid)
    if sequential:
        out = log_prob_sequential(p, bp, x, cd, eps, n_step, dtype)
    else:
        out = log_prob(p, bp, x, cd, eps, dtype)
    return out.reshape(cond.shape[0], grid.shape[0], -1)


def grid_log_prob_shared(p, base_pos, cond, cond_traj, grid, eps, dtype=np.float32):
    """a13-(ii): prefix = per-agent conditioning trajectory, query = g at step i only:
    log p_i(g) = sum_d Normal(prev_i, std_i).log_prob(net[i](g; [cond[a,i], cond_traj[a,:i]])) + l_i,
    prev_0 = base_pos, prev_i = cond_traj[a,i-1].  -> (A,G,T)."""
    p = cast_params(p, dtype)
    base_pos, cond, cond_traj, grid, eps = (np.asarray(a, dtype=dtype) for a in (base_pos, cond, cond_traj, grid, eps))
    spec = Spec(p)
    A, T = cond.shape[0], cond.shape[1]
    G = grid.shape[0]
    off = spec.eps_offsets()
    g_rows = np.tile(grid, (A, 1))
    out = np.empty((A * G, T), dtype=dtype)
    for i in range(T):
        y_a = np.concatenate([cond[:, i], cond_traj[:, :i].reshape(A, -1)], axis=-1)
        y = np.repeat(y_a, G, axis=0)
        w, l = cif_forward(p, spec, i, g_rows, y, eps[:, off[i]:off[i + 1]])
        prev = base_pos if i == 0 else cond_traj[:, i - 1]
        prev = np.repeat(prev, G, axis=0)
        out[:, i] = np.sum(normal_log_prob(w, prev, base_std(p, i)), axis=-1) + l
    return out.reshape(A, G, T)


def log_prob_importance(p, base_pos, x, cond, eps_k, dtype=np.float32):
    """Extension: K importance samples of the CIF auxiliary variable,
    log p ~= logsumexp_k(log_prob with eps_k) - log K.  eps_k (K, N, sum C_i).  K = 1 is the reference."""
    vals = np.stack([log_prob(p, base_pos, x, cond, e, dtype) for e in eps_k])
    m = vals.max(axis=0)
    return m + np.log(np.exp(vals - m).sum(axis=0)) - np.log(vals.dtype.type(len(eps_k)))


# ----------------------------------------------------------------------------------------------
# consumers right after the hot path
# ----------------------------------------------------------------------------------------------

def to_trajectories(state_st, offset, dt, std, state_mode, dtype=np.float32):
    """TP_metrics.unstandardize + output_to_trajectories (src/metrics/TP_metrics.py:68-112) for one tensor
    state_st (A,S,Tk,ncol), ncol = 2 (standardised state) or 3 (+ log p): positions x std, then 'state_v' (mode 1):
    cumsum over the step axis * dt + offset, 'state_p' (mode 0): + offset; the last column becomes exp(log p).
    offset (A,2) = obs[:, -1, :2] for key 0, metric gt[:, k-1, :2] for an update key k; dt (A) (batched = vmap of the
    reference's batch-size-1 broadcasting)."""
    x = np.array(state_st, dtype=dtype, copy=True)
    std = np.asarray(std, dtype=dtype)
    offset = np.asarray(offset, dtype=dtype)
    pos = x[..., :2] * std
    if state_mode == 1:
        pos = np.cumsum(pos, axis=2, dtype=dtype) * np.asarray(dt, dtype=dtype)[:, None, None, None] + offset[:, None, None, :]
    else:
        pos = pos + offset[:, None, None, :]
    x[..., :2] = pos
    if x.shape[-1] == 3:
        x[..., 2] = np.exp(x[..., 2])
    return x


def best_of_n(cand_st, gt, last_obs, dt, fallback_st, std, state_mode, dtype=np.float32):
    """The ADE/FDE best-of-N evaluation of src/main.py:162-180 for one batch: cand_st (A,N,T,2) standardised candidates
    (N = TEST.N_TRIAL predictions), NaN coordinates replaced by fallback_st (fastpredNF.py:131-136), un-standardised and
    integrated (TP_metrics.py:68-112), then displacement_error / final_displacement_error per candidate and the minimum
    over candidates per agent (TP_metrics.py:34-46, 163-191: stack over trials, min, the caller sums over agents).
    gt (A,T,2) metric positions.  Returns ade_min (A), fde_min (A), ade_arg (A), fde_arg (A)."""
    c = np.array(cand_st, dtype=dtype, copy=True)
    if fallback_st is not None:
        fb = np.broadcast_to(np.asarray(fallback_st, dtype=dtype)[:, None, None, :], c.shape)
        c = np.where(np.isnan(c), fb, c)
    pos = to_trajectories(c, last_obs, dt, std, state_mode, dtype)
    err = np.linalg.norm(pos - np.asarray(gt, dtype=dtype)[:, None], axis=-1)       # (A, N, T)
    ade = err.mean(axis=-1)
    fde = err[..., -1]
    return ade.min(axis=1), fde.min(axis=1), ade.argmin(axis=1), fde.argmin(axis=1)


def density_maps_metric(p, base_pos, cond, cond_traj, grid_xy, origin, inv_scale, log_jac, eps, dtype=np.float32, as_prob=True):
    """What replaces TP_visualizer.prob_to_grid (src/visualization/TP_visualizer.py:118-179) for FlowChain: the analytic
    per-step density (a13-ii, prefix = cond_traj) evaluated on a METRIC lattice grid_xy (G,2).  Agent a at step t maps
    the lattice into its standardised state space with z = (X - origin[a,t]) * inv_scale[a] (the inverse of
    TP_metrics.unstandardize / output_to_trajectories, TP_metrics.py:68-112) and the density picks up the Jacobian
    exp(log_jac[a]).  Returns (A,G,T): p, or log p with as_prob=False."""
    p = cast_params(p, dtype)
    base_pos, cond, cond_traj, grid_xy, origin, inv_scale, log_jac, eps = (
        np.asarray(a, dtype=dtype) for a in (base_pos, cond, cond_traj, grid_xy, origin, inv_scale, log_jac, eps))
    spec = Spec(p)
    A, T = cond.shape[0], cond.shape[1]
    G = grid_xy.shape[0]
    off = spec.eps_offsets()
    out = np.empty((A * G, T), dtype=dtype)
    for i in range(T):
        z = ((grid_xy[None] - origin[:, i, None]) * inv_scale[:, None]).reshape(A * G, 2)
        y_a = np.concatenate([cond[:, i], cond_traj[:, :i].reshape(A, -1)], axis=-1)
        w, l = cif_forward(p, spec, i, z, np.repeat(y_a, G, axis=0), eps[:, off[i]:off[i + 1]])
        prev = base_pos if i == 0 else cond_traj[:, i - 1]
        out[:, i] = np.sum(normal_log_prob(w, np.repeat(prev, G, axis=0), base_std(p, i)), axis=-1) + l
    out = out.reshape(A, G, T) + log_jac[:, None, None]
    return np.exp(out) if as_prob else out
