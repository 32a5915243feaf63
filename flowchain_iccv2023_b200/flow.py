"""Host-side mirror of the reference flow interface for the density hot path.

Class, method and state_dict names follow the reference (src/models/TP/fastpredNF.py:380-826,
src/models/TP/TFCondARFlow.py:254-402) so a reference checkpoint loads unchanged and callers
(`fastpredNF_TP`, src/main.py) need no edits.  In eval() mode -- inference, the hot path -- the modules
only OWN the parameters: all arithmetic runs in the fused sm_100a kernels behind ``torch.ops.flowchain.*``
and there is no PyTorch/CPU fallback (without the CUDA library every density call raises).  In train()
mode with grad enabled the same methods run the modules' plain differentiable PyTorch forwards instead
(SURVEY §8b "Registration"): that is the reference's training path (`update()`), not an inference fallback.

Noise: the reference draws ``torch.randn_like`` inside every CIF step (log_prob is stochastic,
SURVEY §0 D7).  Here each call takes ``eps=`` (a noise tape, layouts in include/flowchain_b200.h)
for parity runs, or ``seed=`` for the in-kernel Philox generator; with neither, a fresh seed is
drawn from torch's default generator so repeated calls differ like the reference's.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
from torch.distributions import Normal

from . import _lib, ops, packing
from .synth import FlowSpec


# ------------------------------------------------------------------------------------------------
# modules (same registration order / keys as the reference modules)
#
# Inference (module in eval() mode) never calls these forwards: all arithmetic of the path runs in the fused
# sm_100a kernels.  They exist for the TRAINING path (SURVEY §8b "Registration": with grad enabled the accelerated
# methods defer to the PyTorch modules so `update()` -- Adam on -log_prob, fastpredNF.py:186-202 -- keeps working):
# plain differentiable PyTorch, used only when the flow is in train() mode with grad enabled.
# ------------------------------------------------------------------------------------------------

def _mlp(dims, act):
    layers = []
    for l in range(len(dims) - 1):
        layers.append(nn.Linear(dims[l], dims[l + 1]))
        if l < len(dims) - 2:
            layers.append(act())
    return nn.Sequential(*layers)


class LinearMaskedCoupling(nn.Module):
    """Masked affine coupling (TFCondARFlow.py:336-402): mask, s_net (Tanh), t_net (ReLU); conditioner input [y, x*mask]."""

    def __init__(self, input_size, hidden_size, n_hidden, mask, cond_label_size):
        super().__init__()
        self.register_buffer("mask", mask)
        dims = [input_size + cond_label_size] + [hidden_size] * (n_hidden + 1) + [input_size]
        self.s_net = _mlp(dims, nn.Tanh)
        self.t_net = _mlp(dims, nn.ReLU)
        self.t_net.load_state_dict(self.s_net.state_dict())   # the reference deep-copies s_net into t_net

    def _st(self, kept, y):
        free = 1 - self.mask
        h = torch.cat([y, kept], dim=-1)
        return torch.tanh(self.s_net(h)) * free, self.t_net(h) * free

    def forward(self, x, y):
        log_s, t = self._st(x * self.mask, y)
        return x * torch.exp(log_s) + t, log_s

    def inverse(self, u, y):
        log_s, t = self._st(u * self.mask, y)
        return (u - t) * torch.exp(-log_s), -log_s


class BatchNorm(nn.Module):
    """Invertible batch-norm layer (TFCondARFlow.py:272-330).  train(): batch statistics (unbiased variance) and the
    reference's running update `running <- 0.9 running + 0.1 batch`; eval(): running statistics."""

    def __init__(self, input_size, momentum=0.9, eps=1e-5):
        super().__init__()
        self.momentum, self.eps = momentum, eps
        self.log_gamma = nn.Parameter(torch.zeros(input_size))
        self.beta = nn.Parameter(torch.zeros(input_size))
        self.register_buffer("running_mean", torch.zeros(input_size))
        self.register_buffer("running_var", torch.ones(input_size))

    def _stats(self, x=None):
        if not self.training:
            return self.running_mean, self.running_var
        if x is not None:
            flat = x.reshape(-1, x.shape[-1])
            self.batch_mean, self.batch_var = flat.mean(0), flat.var(0)
            with torch.no_grad():
                self.running_mean.mul_(self.momentum).add_(self.batch_mean.detach() * (1 - self.momentum))
                self.running_var.mul_(self.momentum).add_(self.batch_var.detach() * (1 - self.momentum))
        return self.batch_mean, self.batch_var

    def forward(self, x, y=None):
        mean, var = self._stats(x)
        out = torch.exp(self.log_gamma) * (x - mean) / torch.sqrt(var + self.eps) + self.beta
        return out, (self.log_gamma - 0.5 * torch.log(var + self.eps)).expand_as(x)

    def inverse(self, out, y=None):
        mean, var = self._stats(None)
        x = (out - self.beta) * torch.exp(-self.log_gamma) * torch.sqrt(var + self.eps) + mean
        return x, (0.5 * torch.log(var + self.eps) - self.log_gamma).expand_as(out)


class FlowSequential(nn.Sequential):
    """Container [coupling, BN] x n_blocks (TFCondARFlow.py:254-269): per-dim log-dets summed over modules, then over dims."""

    def forward(self, x, y):
        total = 0
        for module in self:
            x, ldj = module(x, y)
            total = total + ldj
        return x, total.sum(-1)

    def inverse(self, u, y):
        total = 0
        for module in reversed(self):
            u, ldj = module.inverse(u, y)
            total = total + ldj
        return u, total.sum(-1)


_LOG_2PI = 1.8378770664093453


class DiagonalGaussianConditionalDensity(nn.Module):
    """r(u|z,y) / q(u|w,y) (fastpredNF.py:690-733): diagonal Gaussian whose mean / log-std come from two MLPs."""

    def __init__(self, cond_dim, out_dim):
        super().__init__()
        self.shift_net = _mlp([cond_dim, 2 * cond_dim, 2 * cond_dim, out_dim], nn.ReLU)
        self.log_scale_net = _mlp([cond_dim, 2 * cond_dim, 2 * cond_dim, out_dim], nn.ReLU)

    @staticmethod
    def _lp(u, mean, log_std):
        return -0.5 * u.shape[-1] * _LOG_2PI - log_std.sum(-1) - 0.5 * ((u - mean) ** 2 / torch.exp(log_std) ** 2).sum(-1)

    def log_prob(self, u, cond_inputs):
        return self._lp(u, self.shift_net(cond_inputs), self.log_scale_net(cond_inputs))

    def sample(self, cond_inputs, eps=None):
        mean, log_std = self.shift_net(cond_inputs), self.log_scale_net(cond_inputs)
        eps = torch.randn_like(mean) if eps is None else eps
        u = torch.exp(log_std) * eps + mean
        return u, self._lp(u, mean, log_std)


class CIF_step(nn.Module):
    """One continuously-indexed flow step (fastpredNF.py:756-781), K = 1 auxiliary sample."""

    def __init__(self, bijection, input_size, cond_label_size):
        super().__init__()
        self.bijection = bijection
        self.cond_label_size = cond_label_size
        self.r_u_given_z = DiagonalGaussianConditionalDensity(input_size + cond_label_size, cond_label_size)
        self.q_u_given_w = DiagonalGaussianConditionalDensity(input_size + cond_label_size, cond_label_size)

    def forward(self, z, y, eps=None):
        u, log_r = self.r_u_given_z.sample(torch.cat([z, y], dim=-1), eps)
        w, ldj = self.bijection(z, u)
        return w, ldj + self.q_u_given_w.log_prob(u, torch.cat([w, y], dim=-1)) - log_r

    def inverse(self, w, y, eps=None):
        u, log_q = self.q_u_given_w.sample(torch.cat([w, y], dim=-1), eps)
        z, ldj = self.bijection.inverse(w, u)
        return z, ldj + log_q - self.r_u_given_z.log_prob(u, torch.cat([z, y], dim=-1))


def create_RealNVP_step(n_blocks, input_size, hidden_size, n_hidden, cond_label_size):
    """fastpredNF.py:610-628: alternating masks, BatchNorm after every coupling."""
    modules = []
    mask = torch.arange(input_size).float() % 2
    for _ in range(n_blocks):
        modules += [LinearMaskedCoupling(input_size, hidden_size, n_hidden, mask, cond_label_size)]
        mask = 1 - mask
        modules += [BatchNorm(input_size)]
    return FlowSequential(*modules)


# ------------------------------------------------------------------------------------------------
# the flow
# ------------------------------------------------------------------------------------------------

class fastpredNF_CIF_separate_cond(nn.Module):
    """Drop-in for the reference class of the same name (fastpredNF.py:814-826 + base classes :380-607)."""

    def __init__(self, input_size: int, n_blocks: int, hidden_size: int, n_hidden: int, cond_label_size: int,
                 **kwargs):
        super().__init__()
        arch = kwargs.get("flow_architecture", "realNVP")
        if arch != "realNVP":
            raise NotImplementedError(f"flow_architecture {arch!r}: only 'realNVP' is on the accelerated path "
                                      "(no shipped config selects MAF, SURVEY §2)")
        if input_size != 2:
            raise NotImplementedError("the fused kernels require input_size == 2")
        pred_len = kwargs["pred_len"]
        self.input_size = input_size
        self.spec = FlowSpec(pred_len=pred_len, input_size=input_size, cond_size=cond_label_size,
                             hidden_size=hidden_size, n_hidden=n_hidden, n_blocks=n_blocks)
        self.net = nn.ModuleList([
            CIF_step(create_RealNVP_step(n_blocks, input_size, hidden_size, n_hidden, cond_label_size + i * input_size),
                     input_size, cond_label_size + i * input_size)
            for i in range(pred_len)])
        self.var = nn.Parameter(torch.ones(pred_len, input_size) / 10, requires_grad=True)
        self.precision = _lib.FC_PREC_FP32       # default arithmetic path of the MLPs
        self.check_versions = True               # re-pack automatically when parameters change in place
        self._handle: Optional[ops.Handle] = None
        self._packed_version = None
        self._tensors = None

    # -- native handle management ---------------------------------------------------------------
    def _param_version(self):
        if self._tensors is None:
            self._tensors = list(self.parameters()) + list(self.buffers())
        v = 0
        for t in self._tensors:
            v += t._version
        return v

    def invalidate(self):
        """Force a re-pack of the weights on the next call (after training `update()` / load)."""
        self._packed_version = None
        self._tensors = None

    def _load_from_state_dict(self, *a, **k):
        super()._load_from_state_dict(*a, **k)
        self.invalidate()

    def engine(self) -> ops.Handle:
        dev = self.var.device
        if dev.type != "cuda":
            raise RuntimeError("flowchain_b200: move the flow to a CUDA device first (.cuda()); there is no CPU fallback")
        if self._handle is None or self._handle.device != dev:
            if self._handle is not None:
                self._handle.close()
            self._handle = ops.Handle(self.spec, dev)
            self.invalidate()
        if self._packed_version is None or (self.check_versions and self._packed_version != self._param_version()):
            ver = self._param_version()
            blob = packing.flatten_state(self.spec, self.state_dict())
            self._handle.load_weights(blob, ver)
            self._packed_version = ver
        return self._handle

    @staticmethod
    def _seed(seed):
        if seed is not None:
            return int(seed) & 0x7FFFFFFFFFFFFFFF
        return int(torch.randint(0, 2 ** 62, (1,)).item())

    def _dev(self, t, like=None):
        if t is None:
            return None
        return t.to(device=self.var.device, dtype=torch.float32, non_blocking=True).contiguous()

    # -- reference API ----------------------------------------------------------------------------
    def base_dist(self, base_pos, step=0):
        """fastpredNF.base_dist (fastpredNF.py:401-402)."""
        return Normal(base_pos, torch.clamp(self.var[step], min=1e-4))

    def log_prob(self, base_pos, x, cond, eps=None, seed=None, precision=None):
        """flow.log_prob == log_prob_separate (fastpredNF.py:450-452, 460-465): (N,2),(N,T,2),(N,T,C0) -> (N,T)."""
        if self._training_path():
            return self._log_prob_separate_torch(base_pos, x, cond, eps)
        with torch.no_grad():
            h = self.engine()
            src = x.device
            out = torch.ops.flowchain.log_prob(
                h.id, self._dev(base_pos), self._dev(x), self._dev(cond), self._dev(eps),
                0 if eps is not None else self._seed(seed), self.precision if precision is None else precision)
            return out.to(src)

    log_prob_separate = log_prob

    def log_prob_sequential(self, base_pos, x, cond, n_step=None, eps=None, seed=None, precision=None):
        """fastpredNF.log_prob_sequential (fastpredNF.py:454-458): full chain, (N,T) or (N,n_step+1)."""
        if self._training_path():
            return self._log_prob_sequential_torch(base_pos, x, cond, n_step, eps)
        with torch.no_grad():
            h = self.engine()
            src = x.device
            out = torch.ops.flowchain.log_prob_sequential(
                h.id, self._dev(base_pos), self._dev(x), self._dev(cond), self._dev(eps),
                0 if eps is not None else self._seed(seed), -1 if n_step is None else int(n_step),
                self.precision if precision is None else precision)
            return out.to(src)

    def sample_with_log_prob(self, base_pos, cond, z0=None, eps=None, seed=None, precision=None):
        if self._training_path():
            return self._sample_with_log_prob_torch(base_pos, cond, z0, eps)
        with torch.no_grad():
            return self._sample_with_log_prob_cuda(base_pos, cond, z0, eps, seed, precision)

    def _sample_with_log_prob_cuda(self, base_pos, cond, z0=None, eps=None, seed=None, precision=None):
        """fastpredNF.sample_with_log_prob (fastpredNF.py:472-477).

        Accepts the reference's expanded arguments base_pos (B,S,2), cond (B,S,T,C0) (stride-0
        expands are fine: only [:,0] is read -- the reference passes per-agent values expanded over
        S, fastpredNF.py:106-108).  Returns (seq (B,S,T,2), log_prob (B,S,T), ldjs (B,S,T), base_lp (B,S,1)).
        `precision`: FC_PREC_FP32 (CUDA-core sampler) or FC_PREC_F16X3 / FC_PREC_F16 (tensor-core sampler)."""
        h = self.engine()
        src = base_pos.device
        if base_pos.dim() != 3 or cond.dim() != 4:
            raise RuntimeError("sample_with_log_prob expects base_pos (B,S,2) and cond (B,S,T,C0)")
        B, S = base_pos.shape[0], base_pos.shape[1]
        bp = self._dev(base_pos[:, 0])
        cd = self._dev(cond[:, 0])
        z0 = None if z0 is None else self._dev(z0).reshape(B * S, 2)
        eps = None if eps is None else self._dev(eps).reshape(B * S, -1)
        prec = self.precision if precision is None else precision
        outs = torch.ops.flowchain.sample_with_log_prob(h.id, bp, cd, S, z0, eps, self._seed(seed), prec)
        return tuple(o.to(src) for o in outs)

    @torch.no_grad()
    def grid_log_prob(self, base_pos, cond, grid, cond_traj=None, sequential=False, K=1, eps=None, seed=None,
                      map_layout=False, precision=None):
        """Working version of the reference's dead `predict_forward` (fastpredNF.py:166-184): log p for every
        agent x grid point x step.  base_pos (A,2), cond (A,T,C0), grid (G,2) -> (A,G,T) (or (A,T,G) maps).
        cond_traj=None: prefix mode "self" (x[:, i] = g, one literal reference call); cond_traj (A,T,2):
        prefix mode "shared".  K > 1: importance-sampled CIF bound (extension; K = 1 is the reference)."""
        h = self.engine()
        src = grid.device
        out = torch.ops.flowchain.grid_log_prob(
            h.id, self._dev(base_pos), self._dev(cond), self._dev(cond_traj), self._dev(grid), self._dev(eps),
            0 if eps is not None else self._seed(seed),
            _lib.FC_PREFIX_SELF if cond_traj is None else _lib.FC_PREFIX_SHARED, bool(sequential), int(K),
            bool(map_layout), self.precision if precision is None else precision)
        return out.to(src)

    @torch.no_grad()
    def grid_log_prob_to_host(self, base_pos, cond, grid, out=None, chunks=4, cond_traj=None, sequential=False, seed=None,
                              precision=None):
        """Density maps (A, T, G) delivered into (pinned) HOST memory with the device->host copy overlapped with the
        evaluation: the agents are processed in `chunks` launches on the current stream and every finished chunk is copied
        on a side stream while the next one computes (201 MB of maps at config 2 = 3.6 ms of PCIe time, all but the last
        chunk hidden).  Inputs may live on the host (they are tiny).  The in-kernel noise of chunk i is keyed by
        (seed + i, row), so the result is a deterministic function of (seed, chunks) but differs from the unchunked call's
        draw.  Returns `out` (allocated pinned if None); it is complete when the current stream has caught up (the side
        stream is joined back into it), e.g. after `torch.cuda.current_stream().synchronize()`."""
        h = self.engine()
        dev = h.device
        A, G, T = base_pos.shape[0], grid.shape[0], self.spec.pred_len
        if out is None:
            out = torch.empty((A, T, G), dtype=torch.float32).pin_memory()
        if tuple(out.shape) != (A, T, G) or out.dtype != torch.float32 or out.device.type != "cpu":
            raise RuntimeError(f"out must be a float32 host tensor of shape {(A, T, G)}")
        bp = base_pos.to(dev, non_blocking=True)
        cd = cond.to(dev, non_blocking=True)
        gr = grid.to(dev, non_blocking=True)
        ct = None if cond_traj is None else cond_traj.to(dev, non_blocking=True)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        main, side = torch.cuda.current_stream(dev), self._copy_stream
        base_seed = self._seed(seed)
        chunks = max(1, min(int(chunks), A))
        keep = []
        for i in range(chunks):
            lo, hi = (A * i) // chunks, (A * (i + 1)) // chunks
            if hi == lo:
                continue
            maps = self.grid_log_prob(bp[lo:hi], cd[lo:hi], gr, cond_traj=None if ct is None else ct[lo:hi],
                                      sequential=sequential, seed=base_seed + i, map_layout=True, precision=precision)
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
            with torch.cuda.stream(side):
                out[lo:hi].copy_(maps, non_blocking=True)
            maps.record_stream(side)
            keep.append(maps)
        main.wait_stream(side)
        return out

    @torch.no_grad()
    def grid_log_prob_multicast(self, base_pos, cond, grid, mc_ptr, offset_elems, strides, cond_traj=None,
                                sequential=False, seed=None, precision=None):
        """`grid_log_prob` of this rank's shard with the results stored through an NVSwitch multicast address
        (`dist.MulticastMaps`): element (a, g, t) lands at `offset_elems + a*strides[0] + g*strides[1] + t*strides[2]` of
        EVERY rank's symmetric map buffer.  Returns nothing; call the buffer's barrier before reading."""
        from . import ops
        h = self.engine()
        ops.grid_log_prob_multicast(
            h.id, self._dev(base_pos), self._dev(cond), self._dev(cond_traj), self._dev(grid), self._seed(seed),
            _lib.FC_PREFIX_SELF if cond_traj is None else _lib.FC_PREFIX_SHARED, bool(sequential), int(mc_ptr),
            int(offset_elems), int(strides[0]), int(strides[1]), int(strides[2]),
            self.precision if precision is None else precision)

    def forward(self, *a, **k):
        raise RuntimeError("call log_prob / log_prob_sequential / sample_with_log_prob / grid_log_prob")

    # -- training path: the modules' plain differentiable PyTorch forwards ---------------------------------------------
    def _training_path(self) -> bool:
        """train() mode with grad enabled = the reference's eager training path (what `update()` differentiates).
        eval() mode always runs the fused kernels and raises without a B200: there is no CPU inference fallback."""
        return self.training and torch.is_grad_enabled()

    def forward_separate(self, seq, cond, eps=None):
        """fastpredNF_separate_cond.forward_separate (fastpredNF.py:571-584): teacher-forced per-step transform.
        seq (N,T,2), cond (N,T,C0), eps (N, sum_i C_i) | None -> u (N,T,2), ldj (N,T)."""
        N, T = seq.shape[0], seq.shape[1]
        flat = seq.reshape(N, T * self.input_size)
        us, ldjs, off = [], [], 0
        for i, net in enumerate(self.net):
            C = self.spec.cond_width(i)
            y = torch.cat([cond[:, i], flat[:, :i * self.input_size]], dim=-1)
            u, l = net(seq[:, i], y, None if eps is None else eps[:, off:off + C])
            off += C
            us.append(u)
            ldjs.append(l)
        return torch.stack(us, dim=1), torch.stack(ldjs, dim=1)

    def forward_sequential(self, seq, cond, n_step=None, eps=None):
        """fastpredNF_separate_cond.forward_sequential (fastpredNF.py:548-569): result index j carries
        net[0] o ... o net[j] (seq[:, j]); ldj_j = mean over its chain steps.  eps: the sequential tape layout."""
        N, T, D = seq.shape[0], seq.shape[1], self.input_size
        flat = seq.reshape(N, 1, T * D)
        x, cum = seq[:, -1:], 0
        offs, o = [], 0
        for k in range(T):
            offs.append(o)
            o += (T - k) * self.spec.cond_width(k)
        last = -1 if n_step is None else T - 1 - n_step
        for step in range(T - 1, last, -1):
            C = self.spec.cond_width(step)
            y = torch.cat([cond[:, step:], flat[..., :step * D].expand(N, T - step, step * D)], dim=-1)
            e = None if eps is None else eps[:, offs[step]:offs[step] + (T - step) * C].reshape(N, T - step, C)
            x, l = self.net[step](x, y, e)
            cum = cum + l
            if step != 0:
                x = torch.cat([seq[:, step - 1:step], x], dim=1)
                cum = torch.cat([torch.zeros_like(cum[:, :1]), cum], dim=1)
        return x, cum / torch.arange(1, cum.shape[1] + 1, device=cum.device, dtype=cum.dtype)

    def inverse(self, u, cond, eps=None):
        """fastpredNF_separate_cond.inverse (fastpredNF.py:586-607): autoregressive sampling direction.
        u (..., 2), cond (..., T, C0) -> seq (..., T, 2), running mean of the step terms (..., T), step terms (..., T)."""
        seq, ldjs, off = [], [], 0
        lead = u.shape[:-1]
        for step, net in enumerate(self.net):
            C = self.spec.cond_width(step)
            y = cond[..., step, :]
            if step:
                y = torch.cat([y, torch.stack(seq, dim=-2).reshape(*lead, -1)], dim=-1)
            u, l = net.inverse(u, y, None if eps is None else eps[..., off:off + C])
            off += C
            seq.append(u)
            ldjs.append(l)
        ldjs = torch.stack(ldjs, dim=-1)
        counts = torch.arange(1, ldjs.shape[-1] + 1, device=ldjs.device, dtype=ldjs.dtype)
        return torch.stack(seq, dim=-2), torch.cumsum(ldjs, dim=-1) / counts, ldjs

    def _log_prob_separate_torch(self, base_pos, x, cond, eps=None):
        u, ldj = self.forward_separate(x, cond, eps)
        prev = torch.cat([base_pos[:, None], x[:, :-1]], dim=1)
        return self.base_dist(prev, list(range(prev.shape[1]))).log_prob(u).sum(-1) + ldj

    def _log_prob_sequential_torch(self, base_pos, x, cond, n_step=None, eps=None):
        u, ldj = self.forward_sequential(x, cond, n_step, eps)
        return self.base_dist(base_pos[:, None]).log_prob(u).sum(-1) + ldj

    def _sample_with_log_prob_torch(self, base_pos, cond, z0=None, eps=None):
        dist = self.base_dist(base_pos)
        u0 = dist.sample() if z0 is None else z0.reshape(base_pos.shape) * dist.scale + dist.loc
        seq, cum, ldjs = self.inverse(u0, cond, None if eps is None else eps.reshape(*base_pos.shape[:-1], -1))
        base = dist.log_prob(u0).sum(-1, keepdim=True)
        return seq, base + cum, ldjs, base
