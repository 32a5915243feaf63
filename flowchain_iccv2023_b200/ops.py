"""PyTorch custom ops (``torch.ops.flowchain.*``) on the thin C-ABI layer.

PyTorch is plumbing here: it owns device memory and streams.  Each op validates shapes/dtypes
(the C side only sees pointers), passes ``data_ptr()``s and the current CUDA stream to
libflowchain_b200.so and turns a non-zero status into ``RuntimeError`` carrying ``fc_last_error``.
Handles are referred to by integer id so that op schemas stay tensor/int only.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .synth import FlowSpec

_HANDLES: Dict[int, "Handle"] = {}
_NEXT_ID = [1]


class Handle:
    """Owns one fc_handle (packed weights of one flow on one device)."""

    def __init__(self, spec: FlowSpec, device: torch.device):
        lib = _lib.load()
        if device.type != "cuda":
            raise RuntimeError("flowchain_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        self.spec = spec
        self.device = device
        self.index = device.index if device.index is not None else torch.cuda.current_device()
        self.desc = _lib.ModelDesc(spec.pred_len, spec.input_size, spec.cond_size, spec.hidden_size,
                                   spec.n_hidden, spec.n_blocks)
        ptr = ctypes.c_void_p()
        rc = lib.fc_create(ctypes.byref(ptr), ctypes.byref(self.desc), self.index)
        if rc != 0:
            raise RuntimeError("fc_create failed: " + _lib.last_error(None))
        self.ptr = ptr
        self.id = _NEXT_ID[0]
        _NEXT_ID[0] += 1
        _HANDLES[self.id] = self
        self.E = spec.total_eps
        self.Eseq = sum((spec.pred_len - k) * spec.cond_width(k) for k in range(spec.pred_len))

    def check(self, rc: int, what: str):
        if rc != 0:
            raise RuntimeError(f"{what} failed: {_lib.last_error(self.ptr)}")

    def load_weights(self, blob: np.ndarray, version: int = 0):
        lib = _lib.load()
        blob = np.ascontiguousarray(blob, dtype=np.float32)
        with torch.cuda.device(self.index):
            stream = torch.cuda.current_stream().cuda_stream
            self.check(lib.fc_load_weights(self.ptr, blob.ctypes.data, blob.size, version, stream), "fc_load_weights")

    @property
    def launch_count(self) -> int:
        return int(_lib.load().fc_launch_count(self.ptr))

    def close(self):
        if self.ptr:
            _lib.load().fc_destroy(self.ptr)
            self.ptr = None
            _HANDLES.pop(self.id, None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _h(hid: int) -> Handle:
    try:
        return _HANDLES[hid]
    except KeyError:
        raise RuntimeError(f"unknown flowchain handle id {hid}")


def _chk(t: Optional[torch.Tensor], shape, name: str, h: Handle) -> int:
    """Validate a float32 contiguous CUDA tensor on the handle's device; returns its data_ptr (0 for None)."""
    if t is None:
        return 0
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name}: expected float32, got {t.dtype}")
    if not t.is_cuda or t.device.index != h.index:
        raise RuntimeError(f"{name}: expected a tensor on cuda:{h.index}, got {t.device}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise RuntimeError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
    return t.data_ptr()


def _stream(h: Handle) -> int:
    return torch.cuda.current_stream(h.index).cuda_stream


# ------------------------------------------------------------------------------------------------
# op definitions
# ------------------------------------------------------------------------------------------------

@torch.library.custom_op("flowchain::log_prob", mutates_args=())
def log_prob(hid: int, base_pos: torch.Tensor, x: torch.Tensor, cond: torch.Tensor,
             eps: Optional[torch.Tensor], seed: int, precision: int) -> torch.Tensor:
    h = _h(hid)
    s = h.spec
    N = x.shape[0]
    out = torch.empty((N, s.pred_len), dtype=torch.float32, device=x.device)
    rc = _lib.load().fc_log_prob(
        h.ptr, _chk(base_pos, (N, 2), "base_pos", h), _chk(x, (N, s.pred_len, 2), "x", h),
        _chk(cond, (N, s.pred_len, s.cond_size), "cond", h), _chk(eps, (N, h.E), "eps", h),
        seed, out.data_ptr(), N, precision, _stream(h))
    h.check(rc, "fc_log_prob")
    return out


@log_prob.register_fake
def _(hid, base_pos, x, cond, eps, seed, precision):
    return x.new_empty((x.shape[0], x.shape[1]))


@torch.library.custom_op("flowchain::log_prob_sequential", mutates_args=())
def log_prob_sequential(hid: int, base_pos: torch.Tensor, x: torch.Tensor, cond: torch.Tensor,
                        eps: Optional[torch.Tensor], seed: int, n_step: int, precision: int) -> torch.Tensor:
    h = _h(hid)
    s = h.spec
    N = x.shape[0]
    n_out = s.pred_len if n_step < 0 else n_step + 1
    out = torch.empty((N, n_out), dtype=torch.float32, device=x.device)
    rc = _lib.load().fc_log_prob_sequential(
        h.ptr, _chk(base_pos, (N, 2), "base_pos", h), _chk(x, (N, s.pred_len, 2), "x", h),
        _chk(cond, (N, s.pred_len, s.cond_size), "cond", h), _chk(eps, (N, h.Eseq), "eps", h),
        seed, n_step, out.data_ptr(), N, precision, _stream(h))
    h.check(rc, "fc_log_prob_sequential")
    return out


@log_prob_sequential.register_fake
def _(hid, base_pos, x, cond, eps, seed, n_step, precision):
    return x.new_empty((x.shape[0], x.shape[1] if n_step < 0 else n_step + 1))


@torch.library.custom_op("flowchain::grid_log_prob", mutates_args=())
def grid_log_prob(hid: int, base_pos: torch.Tensor, cond: torch.Tensor, cond_traj: Optional[torch.Tensor],
                  grid: torch.Tensor, eps: Optional[torch.Tensor], seed: int, prefix_mode: int,
                  sequential: bool, K: int, map_layout: bool, precision: int) -> torch.Tensor:
    """Returns (A, G, T) (reference row order) or, with map_layout, (A, T, G) density maps."""
    h = _h(hid)
    s = h.spec
    A, G, T = base_pos.shape[0], grid.shape[0], s.pred_len
    width = h.Eseq if sequential else h.E
    if map_layout:
        out = torch.empty((A, T, G), dtype=torch.float32, device=grid.device)
        sa, sp, st = T * G, 1, G
    else:
        out = torch.empty((A, G, T), dtype=torch.float32, device=grid.device)
        sa, sp, st = G * T, T, 1
    rc = _lib.load().fc_grid_log_prob(
        h.ptr, _chk(base_pos, (A, 2), "base_pos", h), _chk(cond, (A, T, s.cond_size), "cond", h),
        _chk(cond_traj, (A, T, 2), "cond_traj", h), _chk(grid, (G, 2), "grid", h),
        _chk(eps, (A * G * K, width), "eps", h), seed, prefix_mode, int(sequential), K, out.data_ptr(),
        sa, sp, st, A, G, precision, _stream(h))
    h.check(rc, "fc_grid_log_prob")
    return out


@grid_log_prob.register_fake
def _(hid, base_pos, cond, cond_traj, grid, eps, seed, prefix_mode, sequential, K, map_layout, precision):
    A, G, T = base_pos.shape[0], grid.shape[0], cond.shape[1]
    return grid.new_empty((A, T, G) if map_layout else (A, G, T))


def grid_log_prob_multicast(hid: int, base_pos: torch.Tensor, cond: torch.Tensor, cond_traj: Optional[torch.Tensor],
                            grid: torch.Tensor, seed: int, prefix_mode: int, sequential: bool, mc_ptr: int,
                            offset_elems: int, sa: int, sp: int, st: int, precision: int) -> None:
    """fc_grid_log_prob_multicast: the (a, g, t) results are stored to `mc_ptr + 4*(offset_elems + a*sa + g*sp + t*st)`,
    an NVSwitch multicast address, so every rank's symmetric map buffer receives them (no all-gather afterwards).
    Not a torch custom op: the destination is a raw multicast address, not a tensor; the caller owns the barrier."""
    h = _h(hid)
    s = h.spec
    A, G, T = base_pos.shape[0], grid.shape[0], s.pred_len
    if mc_ptr == 0:
        raise RuntimeError("multicast pointer is null: the symmetric buffer has no NVSwitch multicast mapping")
    rc = _lib.load().fc_grid_log_prob_multicast(
        h.ptr, _chk(base_pos, (A, 2), "base_pos", h), _chk(cond, (A, T, s.cond_size), "cond", h),
        _chk(cond_traj, (A, T, 2), "cond_traj", h), _chk(grid, (G, 2), "grid", h), None, seed, prefix_mode,
        int(sequential), mc_ptr + 4 * offset_elems, sa, sp, st, A, G, precision, _stream(h))
    h.check(rc, "fc_grid_log_prob_multicast")


@torch.library.custom_op("flowchain::sample_with_log_prob", mutates_args=())
def sample_with_log_prob(hid: int, base_pos: torch.Tensor, cond: torch.Tensor, S: int,
                         z0: Optional[torch.Tensor], eps: Optional[torch.Tensor], seed: int,
                         precision: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    h = _h(hid)
    s = h.spec
    B, T = base_pos.shape[0], s.pred_len
    dev = base_pos.device
    seq = torch.empty((B, S, T, 2), dtype=torch.float32, device=dev)
    lp = torch.empty((B, S, T), dtype=torch.float32, device=dev)
    ldjs = torch.empty((B, S, T), dtype=torch.float32, device=dev)
    base = torch.empty((B, S, 1), dtype=torch.float32, device=dev)
    rc = _lib.load().fc_sample_with_log_prob(
        h.ptr, _chk(base_pos, (B, 2), "base_pos", h), _chk(cond, (B, T, s.cond_size), "cond", h),
        _chk(z0, (B * S, 2), "z0", h), _chk(eps, (B * S, h.E), "eps", h), seed,
        seq.data_ptr(), lp.data_ptr(), ldjs.data_ptr(), base.data_ptr(), B, S, precision, _stream(h))
    h.check(rc, "fc_sample_with_log_prob")
    return seq, lp, ldjs, base


@sample_with_log_prob.register_fake
def _(hid, base_pos, cond, S, z0, eps, seed, precision):
    B, T = base_pos.shape[0], cond.shape[1]
    return (base_pos.new_empty((B, S, T, 2)), base_pos.new_empty((B, S, T)), base_pos.new_empty((B, S, T)),
            base_pos.new_empty((B, S, 1)))


@torch.library.custom_op("flowchain::base_normalizer", mutates_args=())
def base_normalizer(hid: int, base_lp: torch.Tensor) -> torch.Tensor:
    h = _h(hid)
    A, S = base_lp.shape[0], base_lp.shape[1]
    out = torch.empty((A, 1), dtype=torch.float32, device=base_lp.device)
    rc = _lib.load().fc_base_normalizer(h.ptr, _chk(base_lp, None, "base_lp", h), out.data_ptr(), A, S, _stream(h))
    h.check(rc, "fc_base_normalizer")
    return out


@base_normalizer.register_fake
def _(hid, base_lp):
    return base_lp.new_empty((base_lp.shape[0], 1))


@torch.library.custom_op("flowchain::update", mutates_args=())
def update(hid: int, new_pos: torch.Tensor, time_step: int, prob_st_0: torch.Tensor, ldjs: torch.Tensor,
           normalizer: torch.Tensor) -> torch.Tensor:
    h = _h(hid)
    T = h.spec.pred_len
    A, S = prob_st_0.shape[0], prob_st_0.shape[1]
    if not (0 < time_step < T):
        raise AssertionError(f"time_step for update need to be 0~{T - 1}, got {time_step}")
    out = torch.empty((A, S, T - time_step, 3), dtype=torch.float32, device=prob_st_0.device)
    rc = _lib.load().fc_update(
        h.ptr, _chk(new_pos, (A, 2), "new_pos", h), time_step, _chk(prob_st_0, (A, S, T, 3), "prob_st_0", h),
        _chk(ldjs, (A, S, T), "ldjs", h), _chk(normalizer.reshape(-1), (A,), "normalizer", h),
        out.data_ptr(), A, S, _stream(h))
    h.check(rc, "fc_update")
    return out


@update.register_fake
def _(hid, new_pos, time_step, prob_st_0, ldjs, normalizer):
    A, S, T, _ = prob_st_0.shape
    return prob_st_0.new_empty((A, S, T - time_step, 3))


@torch.library.custom_op("flowchain::update_lazy", mutates_args=())
def update_lazy(hid: int, new_pos: torch.Tensor, time_step: int, prob_st_0: torch.Tensor,
                normalizer: torch.Tensor) -> torch.Tensor:
    """base_lp_t (A, S): the per-sample term of the rapid update; prob_st_t[..., 2] == base_lp_t[..., None] + ldjs[:, :, t:]."""
    h = _h(hid)
    T = h.spec.pred_len
    A, S = prob_st_0.shape[0], prob_st_0.shape[1]
    if not (0 < time_step < T):
        raise AssertionError(f"time_step for update need to be 0~{T - 1}, got {time_step}")
    out = torch.empty((A, S), dtype=torch.float32, device=prob_st_0.device)
    rc = _lib.load().fc_update_lazy(
        h.ptr, _chk(new_pos, (A, 2), "new_pos", h), time_step, _chk(prob_st_0, (A, S, T, 3), "prob_st_0", h),
        _chk(normalizer.reshape(-1), (A,), "normalizer", h), out.data_ptr(), A, S, _stream(h))
    h.check(rc, "fc_update_lazy")
    return out


@update_lazy.register_fake
def _(hid, new_pos, time_step, prob_st_0, normalizer):
    return prob_st_0.new_empty((prob_st_0.shape[0], prob_st_0.shape[1]))


# ------------------------------------------------------------------------------------------------
# consumers right after the hot path
# ------------------------------------------------------------------------------------------------

@torch.library.custom_op("flowchain::grid_density_maps", mutates_args=())
def grid_density_maps(hid: int, base_pos: torch.Tensor, cond: torch.Tensor, cond_traj: torch.Tensor, grid_xy: torch.Tensor,
                      origin: torch.Tensor, inv_scale: torch.Tensor, log_jac: Optional[torch.Tensor],
                      eps: Optional[torch.Tensor], seed: int, as_prob: bool, map_layout: bool, precision: int) -> torch.Tensor:
    """Per-step density on a METRIC lattice: out[a,g,t] = p_t((grid_xy[g] - origin[a,t]) * inv_scale[a] | cond_traj[a,:t])
    * exp(log_jac[a]) (or its log).  (A, G, T), or (A, T, G) with map_layout."""
    h = _h(hid)
    s = h.spec
    A, G, T = base_pos.shape[0], grid_xy.shape[0], s.pred_len
    if map_layout:
        out = torch.empty((A, T, G), dtype=torch.float32, device=grid_xy.device)
        sa, sp, st = T * G, 1, G
    else:
        out = torch.empty((A, G, T), dtype=torch.float32, device=grid_xy.device)
        sa, sp, st = G * T, T, 1
    rc = _lib.load().fc_grid_density_maps(
        h.ptr, _chk(base_pos, (A, 2), "base_pos", h), _chk(cond, (A, T, s.cond_size), "cond", h),
        _chk(cond_traj, (A, T, 2), "cond_traj", h), _chk(grid_xy, (G, 2), "grid_xy", h), _chk(origin, (A, T, 2), "origin", h),
        _chk(inv_scale, (A, 2), "inv_scale", h), _chk(log_jac, (A,), "log_jac", h), _chk(eps, (A * G, h.E), "eps", h),
        seed, int(as_prob), out.data_ptr(), sa, sp, st, A, G, precision, _stream(h))
    h.check(rc, "fc_grid_density_maps")
    return out


@grid_density_maps.register_fake
def _(hid, base_pos, cond, cond_traj, grid_xy, origin, inv_scale, log_jac, eps, seed, as_prob, map_layout, precision):
    A, G, T = base_pos.shape[0], grid_xy.shape[0], cond.shape[1]
    return grid_xy.new_empty((A, T, G) if map_layout else (A, G, T))


@torch.library.custom_op("flowchain::to_trajectories", mutates_args=())
def to_trajectories(hid: int, state_st: torch.Tensor, offset: torch.Tensor, dt: Optional[torch.Tensor], std0: float, std1: float,
                    state_mode: int) -> torch.Tensor:
    """TP_metrics.unstandardize + output_to_trajectories for one (A, S, Tk, 2|3) tensor (see fc_to_trajectories)."""
    h = _h(hid)
    if state_st.dim() != 4 or state_st.shape[-1] not in (2, 3):
        raise RuntimeError(f"state_st: expected (A, S, Tk, 2|3), got {tuple(state_st.shape)}")
    A, S, Tk, ncol = state_st.shape
    out = torch.empty_like(state_st)
    rc = _lib.load().fc_to_trajectories(
        h.ptr, _chk(state_st, None, "state_st", h), _chk(offset, (A, 2), "offset", h), _chk(dt, (A,), "dt", h),
        float(std0), float(std1), state_mode, out.data_ptr(), A, S, Tk, ncol, _stream(h))
    h.check(rc, "fc_to_trajectories")
    return out


@to_trajectories.register_fake
def _(hid, state_st, offset, dt, std0, std1, state_mode):
    return torch.empty_like(state_st)


@torch.library.custom_op("flowchain::best_of_n", mutates_args=())
def best_of_n(hid: int, cand_st: torch.Tensor, gt: torch.Tensor, last_obs: torch.Tensor, dt: Optional[torch.Tensor],
              fallback_st: Optional[torch.Tensor], std0: float, std1: float,
              state_mode: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Per-agent best-of-N ADE / FDE with the winning candidate indices (see fc_best_of_n)."""
    h = _h(hid)
    A, N, T, _ = cand_st.shape
    dev = cand_st.device
    ade = torch.empty((A,), dtype=torch.float32, device=dev)
    fde = torch.empty((A,), dtype=torch.float32, device=dev)
    ia = torch.empty((A,), dtype=torch.int32, device=dev)
    jf = torch.empty((A,), dtype=torch.int32, device=dev)
    rc = _lib.load().fc_best_of_n(
        h.ptr, _chk(cand_st, (A, N, T, 2), "cand_st", h), _chk(gt, (A, T, 2), "gt", h), _chk(last_obs, (A, 2), "last_obs", h),
        _chk(dt, (A,), "dt", h), _chk(fallback_st, (A, 2), "fallback_st", h), float(std0), float(std1), state_mode,
        ade.data_ptr(), fde.data_ptr(), ia.data_ptr(), jf.data_ptr(), A, N, T, _stream(h))
    h.check(rc, "fc_best_of_n")
    return ade, fde, ia, jf


@best_of_n.register_fake
def _(hid, cand_st, gt, last_obs, dt, fallback_st, std0, std1, state_mode):
    A = cand_st.shape[0]
    return (cand_st.new_empty((A,)), cand_st.new_empty((A,)), cand_st.new_empty((A,), dtype=torch.int32),
            cand_st.new_empty((A,), dtype=torch.int32))
