"""state_dict -> canonical fp32 weight blob consumed by fc_load_weights (csrc/capi.cu).

The PyTorch modules stay the owners of the parameters (reference state_dict layout);
the native side holds a re-laid-out read-only copy.  Blob order:

    var[T*D]
    for step i in 0..T-1:
        for block b in 0..n_blocks-1:
            bijection.{2b}.mask[D]
            s_net: for l in 0..n_hidden+1: weight[out][in] (row-major), bias[out]
            t_net: likewise
            bijection.{2b+1}: log_gamma[D], beta[D], running_mean[D], running_var[D]
        for dens in (r_u_given_z, q_u_given_w): for net in (shift_net, log_scale_net):
            for l in 0..2: weight[out][in], bias[out]
"""
from __future__ import annotations

from typing import Callable, Mapping

import numpy as np

from .synth import FlowSpec


def blob_keys(spec: FlowSpec):
    yield "var"
    for i in range(spec.pred_len):
        for b in range(spec.n_blocks):
            pre = f"net.{i}.bijection.{2 * b}"
            yield pre + ".mask"
            for net in ("s_net", "t_net"):
                for l in range(spec.n_hidden + 2):
                    yield f"{pre}.{net}.{2 * l}.weight"
                    yield f"{pre}.{net}.{2 * l}.bias"
            bn = f"net.{i}.bijection.{2 * b + 1}"
            for k in ("log_gamma", "beta", "running_mean", "running_var"):
                yield f"{bn}.{k}"
        for dens in ("r_u_given_z", "q_u_given_w"):
            for net in ("shift_net", "log_scale_net"):
                for l in range(3):
                    yield f"net.{i}.{dens}.{net}.{2 * l}.weight"
                    yield f"net.{i}.{dens}.{net}.{2 * l}.bias"


def _expected_shape(spec: FlowSpec, key: str):
    parts = key.split(".")
    D, H = spec.input_size, spec.hidden_size
    if key == "var":
        return (spec.pred_len, D)
    i = int(parts[1])
    C = spec.cond_width(i)
    c = C + D
    leaf = parts[-1]
    if parts[2] == "bijection":
        if leaf in ("mask", "log_gamma", "beta", "running_mean", "running_var"):
            return (D,)
        l = int(parts[5]) // 2
        n_in = c if l == 0 else H
        n_out = D if l == spec.n_hidden + 1 else H
    else:
        l = int(parts[4]) // 2
        dims = [c, 2 * c, 2 * c, C]
        n_in, n_out = dims[l], dims[l + 1]
    return (n_out, n_in) if leaf == "weight" else (n_out,)


def flatten_state(spec: FlowSpec, state: Mapping, prefix: str = "") -> np.ndarray:
    """Concatenate the tensors of `state` (numpy arrays or torch tensors keyed like the reference
    state_dict, optionally under `prefix`, e.g. "flow.") into the canonical float32 blob."""
    chunks = []
    for key in blob_keys(spec):
        t = state[prefix + key]
        if hasattr(t, "detach"):
            t = t.detach().to("cpu").float().numpy()
        a = np.asarray(t, dtype=np.float32)
        want = _expected_shape(spec, key)
        if tuple(a.shape) != want:
            raise ValueError(f"{prefix + key}: shape {tuple(a.shape)} does not match the model spec {want}")
        chunks.append(a.reshape(-1))
    return np.ascontiguousarray(np.concatenate(chunks))
