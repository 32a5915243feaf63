"""Deterministic synthetic weights and inputs for the FlowChain density hot path.

There is no network for checkpoints or datasets, so every test and benchmark
runs on random-init weights of the reference architecture and synthetic
ETH-shaped inputs (BASELINE.json configs).  Everything here is a pure function
of (spec, seed): a counter-based splitmix64 hash -> uniforms -> Box-Muller, all
in numpy integer/float64 arithmetic, so the container that generates golden
vectors and the GPU box that replays them agree without torch's RNG.

The parameter dictionary uses the reference's ``state_dict`` keys of
``fastpredNF_CIF_separate_cond`` (reference: src/models/TP/fastpredNF.py:814-826,
690-710, 756-767; src/models/TP/TFCondARFlow.py:272-284, 336-358) so the same
dictionary can be loaded into the unmodified reference module with
``load_state_dict`` (tests/golden/make_golden.py does that).
"""
from __future__ import annotations

import dataclasses
import zlib
from typing import Dict

import numpy as np


@dataclasses.dataclass(frozen=True)
class FlowSpec:
    """Shape knobs of the chain (reference: src/default_params.py:30-35, fastpredNF.py:58-70)."""
    pred_len: int = 12          # T, number of chained CIF steps
    input_size: int = 2         # D, flow dimension (kernels require 2)
    cond_size: int = 16         # C0 = MODEL.FLOW.CONDITIONING_LENGTH
    hidden_size: int = 64       # H  = MODEL.FLOW.HIDDEN_SIZE
    n_hidden: int = 2           # MODEL.FLOW.N_HIDDEN
    n_blocks: int = 3           # MODEL.FLOW.N_BLOCKS

    def cond_width(self, step: int) -> int:
        """C_i: width of the conditioning vector of chain step i (fastpredNF.py:823)."""
        return self.cond_size + step * self.input_size

    @property
    def total_eps(self) -> int:
        return sum(self.cond_width(i) for i in range(self.pred_len))

    def macs_per_step(self, step: int) -> int:
        """MACs of every nn.Linear of one CIF step as shaped in the reference."""
        C, D, H = self.cond_width(step), self.input_size, self.hidden_size
        c = C + D
        net = (c * H) + self.n_hidden * H * H + H * D
        dens = 2 * (c * 2 * c + (2 * c) * (2 * c) + 2 * c * C)
        return 2 * dens + 2 * self.n_blocks * net

    def flops_separate_row(self) -> int:
        return 2 * sum(self.macs_per_step(i) for i in range(self.pred_len))

    def flops_sequential_row(self) -> int:
        return 2 * sum((self.pred_len - i) * self.macs_per_step(i) for i in range(self.pred_len))


DEFAULT_SPEC = FlowSpec()

_MASK64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _MASK64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _MASK64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _MASK64
        return z ^ (z >> np.uint64(31))


def _stream(seed: int, tag: str) -> np.uint64:
    h = zlib.crc32(tag.encode()) & 0xFFFFFFFF
    return _splitmix64(np.array([(seed * 0x100000001B3 + h) & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))[0]


def uniform01(seed: int, tag: str, n: int, lane: int = 0) -> np.ndarray:
    """n float64 uniforms in (0,1), a pure function of (seed, tag, lane, index)."""
    base = _stream(seed, tag)
    idx = np.arange(n, dtype=np.uint64) * np.uint64(2) + np.uint64(lane)
    with np.errstate(over="ignore"):
        bits = _splitmix64((idx * np.uint64(0xD1342543DE82EF95) + base) & _MASK64)
    return ((bits >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def normal(seed: int, tag: str, shape) -> np.ndarray:
    """float64 standard normals (Box-Muller over two hashed uniform lanes)."""
    n = int(np.prod(shape)) if len(shape) else 1
    u1 = uniform01(seed, tag, n, 0)
    u2 = uniform01(seed, tag, n, 1)
    z = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    return z.reshape(shape)


def uniform(seed: int, tag: str, shape, lo: float, hi: float) -> np.ndarray:
    n = int(np.prod(shape)) if len(shape) else 1
    return (lo + (hi - lo) * uniform01(seed, tag, n, 0)).reshape(shape)


def _linear(params: Dict[str, np.ndarray], key: str, n_in: int, n_out: int, seed: int) -> None:
    bound = 1.0 / np.sqrt(n_in)  # same scale as nn.Linear's default init
    params[key + ".weight"] = uniform(seed, key + ".weight", (n_out, n_in), -bound, bound).astype(np.float32)
    params[key + ".bias"] = uniform(seed, key + ".bias", (n_out,), -bound, bound).astype(np.float32)


def make_params(spec: FlowSpec = DEFAULT_SPEC, seed: int = 0, perturb_bn: bool = True) -> Dict[str, np.ndarray]:
    """Random-init parameters in the reference state_dict layout (float32 numpy).

    BatchNorm statistics / log_gamma / beta and ``var`` are perturbed away from
    their constructor values (0/1/0/0 and 0.1) because those would hide bugs
    (SURVEY §4 item 4).  s_net and t_net are drawn independently.
    """
    T, D, H = spec.pred_len, spec.input_size, spec.hidden_size
    p: Dict[str, np.ndarray] = {}
    if perturb_bn:
        p["var"] = uniform(seed, "var", (T, D), 0.06, 0.16).astype(np.float32)
    else:
        p["var"] = np.full((T, D), 0.1, dtype=np.float32)
    mask0 = (np.arange(D) % 2).astype(np.float32)
    for i in range(T):
        C = spec.cond_width(i)
        c = C + D
        mask = mask0.copy()
        for b in range(spec.n_blocks):
            pre = f"net.{i}.bijection.{2 * b}"
            p[pre + ".mask"] = mask.copy()
            for net in ("s_net", "t_net"):
                dims = [c] + [H] * (spec.n_hidden + 1) + [D]
                for l in range(len(dims) - 1):
                    _linear(p, f"{pre}.{net}.{2 * l}", dims[l], dims[l + 1], seed)
            bn = f"net.{i}.bijection.{2 * b + 1}"
            if perturb_bn:
                p[bn + ".log_gamma"] = (0.1 * normal(seed, bn + ".log_gamma", (D,))).astype(np.float32)
                p[bn + ".beta"] = (0.1 * normal(seed, bn + ".beta", (D,))).astype(np.float32)
                p[bn + ".running_mean"] = (0.2 * normal(seed, bn + ".running_mean", (D,))).astype(np.float32)
                p[bn + ".running_var"] = uniform(seed, bn + ".running_var", (D,), 0.5, 1.5).astype(np.float32)
            else:
                p[bn + ".log_gamma"] = np.zeros(D, np.float32)
                p[bn + ".beta"] = np.zeros(D, np.float32)
                p[bn + ".running_mean"] = np.zeros(D, np.float32)
                p[bn + ".running_var"] = np.ones(D, np.float32)
            mask = 1.0 - mask
        for dens in ("r_u_given_z", "q_u_given_w"):
            for net in ("shift_net", "log_scale_net"):
                dims = [c, 2 * c, 2 * c, C]
                for l in range(3):
                    _linear(p, f"net.{i}.{dens}.{net}.{2 * l}", dims[l], dims[l + 1], seed)
    return p


def make_inputs(spec: FlowSpec, agents: int, seed: int = 1) -> Dict[str, np.ndarray]:
    """ETH-shaped synthetic batch: per-agent encoder output ``cond`` repeated over
    the T steps (mirrors fastpredNF.py:374-375), last observed velocity ``base_pos`` and a
    conditioning trajectory ``cond_traj`` in standardised-velocity units."""
    T, D, C0 = spec.pred_len, spec.input_size, spec.cond_size
    cond1 = normal(seed, "cond", (agents, 1, C0)).astype(np.float32)
    return {
        "cond": np.ascontiguousarray(np.broadcast_to(cond1, (agents, T, C0))),
        "base_pos": (0.3 * normal(seed, "base_pos", (agents, D))).astype(np.float32),
        "cond_traj": (0.3 * normal(seed, "cond_traj", (agents, T, D))).astype(np.float32),
    }


def make_grid(n_per_dim: int, lo: float = -3.0, hi: float = 3.0) -> np.ndarray:
    """cartesian_prod(linspace(lo,hi,n), linspace(lo,hi,n)) as (n*n, 2) float32
    (the query lattice of fastpredNF.py:177-178 / TP_visualizer.py:53)."""
    ax = np.linspace(lo, hi, n_per_dim, dtype=np.float32)
    gx, gy = np.meshgrid(ax, ax, indexing="ij")
    return np.stack([gx.ravel(), gy.ravel()], axis=-1).astype(np.float32)


def make_eps(spec: FlowSpec, rows: int, seed: int = 2, tag: str = "eps") -> np.ndarray:
    """Noise tape for the separate density path: (rows, sum_i C_i) float32, step i occupying
    columns [off_i, off_i + C_i) -- the K=1 auxiliary draw of CIF_step.forward (fastpredNF.py:727-733)."""
    return normal(seed, tag, (rows, spec.total_eps)).astype(np.float32)
