"""Generate tests/golden/*.npz by running the UNMODIFIED reference classes from /root/reference.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

The reference has no tests or golden vectors; its own code, executed with an
injected noise tape, is the pin.  For each case we store the inputs, the noise tape and the
reference outputs in fp32 (what the reference computes) and fp64 (ground truth for tolerances).
Weights are NOT stored: they are a pure function of (spec, seed) via flowchain_iccv2023_b200.synth.

Reference entry points exercised (src/models/TP/fastpredNF.py):
  flow.log_prob :450 | flow.log_prob_sequential(n_step) :454 | flow.sample_with_log_prob :472
  fastpredNF_TP.predict / predict_from_new_obs :94-164 (via object.__new__ + stub encoder)
  CIF_step.forward / .inverse :769-781 per step (used for prefix mode "shared")
"""
from __future__ import annotations

import copy
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/src"

from flowchain_iccv2023_b200 import synth  # noqa: E402


def import_reference():
    if "yacs" not in sys.modules:  # annotations only (fastpredNF.py:1, utils.py:4)
        yacs = types.ModuleType("yacs")
        cfgmod = types.ModuleType("yacs.config")

        class CfgNode(dict):
            pass
        cfgmod.CfgNode = CfgNode
        yacs.config = cfgmod
        sys.modules["yacs"] = yacs
        sys.modules["yacs.config"] = cfgmod
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import models.TP.fastpredNF as ref  # noqa
    return ref


class NoiseTape:
    """Replays injected noise in the reference's draw order: patches torch.randn_like
    (CIF auxiliary eps) and torch.distributions.Normal.sample (base draw)."""

    def __init__(self, randn_list, z0=None):
        self.randn_list = list(randn_list)
        self.z0 = z0
        self.i = 0

    def __enter__(self):
        self._randn_like = torch.randn_like
        self._sample = torch.distributions.Normal.sample
        tape = self

        def randn_like(t, *a, **k):
            e = tape.randn_list[tape.i]
            tape.i += 1
            assert tuple(e.shape) == tuple(t.shape), (e.shape, t.shape)
            return e.to(t.dtype)

        def sample(dist, sample_shape=torch.Size()):
            z = tape.z0.to(dist.loc.dtype)
            return z * dist.scale.expand(z.shape) + dist.loc.expand(z.shape)   # normal_()*std + mean

        torch.randn_like = randn_like
        torch.distributions.Normal.sample = sample
        return self

    def __exit__(self, *exc):
        torch.randn_like = self._randn_like
        torch.distributions.Normal.sample = self._sample
        assert self.i == len(self.randn_list), "tape not fully consumed"


def build_flow(ref, spec: synth.FlowSpec, seed: int):
    flow = ref.fastpredNF_CIF_separate_cond(
        input_size=spec.input_size, n_blocks=spec.n_blocks, n_hidden=spec.n_hidden,
        hidden_size=spec.hidden_size, cond_label_size=spec.cond_size,
        flow_architecture="realNVP", pred_len=spec.pred_len)
    params = synth.make_params(spec, seed)
    missing = flow.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    flow.eval()
    return flow


def sep_tape(spec, eps):
    out, o = [], 0
    for i in range(spec.pred_len):
        C = spec.cond_width(i)
        out.append(eps[:, o:o + C])
        o += C
    return out


def seq_tape(spec, eps, n_step=None):
    T = spec.pred_len
    offs, o = [], 0
    for k in range(T):
        offs.append(o)
        o += (T - k) * spec.cond_width(k)
    stop = -1 if n_step is None else T - 1 - n_step
    return [eps[:, offs[k]:offs[k] + (T - k) * spec.cond_width(k)].reshape(eps.shape[0], T - k, spec.cond_width(k))
            for k in range(T - 1, stop, -1)]


def make_case(ref, name, spec, seed, agents, grid_n, S, lo=-3.0, hi=3.0):
    flow32 = build_flow(ref, spec, seed)
    flow64 = copy.deepcopy(flow32).double()
    T, D = spec.pred_len, spec.input_size
    inp = synth.make_inputs(spec, agents, seed + 1)
    out = {"spec": np.array([spec.pred_len, spec.input_size, spec.cond_size, spec.hidden_size,
                             spec.n_hidden, spec.n_blocks], dtype=np.int64),
           "seed": np.array(seed)}
    out.update({k: v for k, v in inp.items()})

    # ---- (1) general rows: x = arbitrary trajectories near base_pos -------------------------
    N = agents
    x = (inp["base_pos"][:, None, :] + 0.5 * synth.normal(seed + 3, "x", (N, T, D))).astype(np.float32)
    eps = synth.make_eps(spec, N, seed + 2)
    eps_seq = synth.normal(seed + 4, "eps_seq", (N, sum((T - k) * spec.cond_width(k) for k in range(T)))).astype(np.float32)
    out.update(x=x, eps=eps, eps_seq=eps_seq)
    for tag, flow, dt in (("32", flow32, torch.float32), ("64", flow64, torch.float64)):
        tb, tx, tc = (torch.from_numpy(a).to(dt) for a in (inp["base_pos"], x, inp["cond"]))
        with torch.no_grad():
            with NoiseTape(sep_tape(spec, torch.from_numpy(eps))):
                out["log_prob_" + tag] = flow.log_prob(tb, tx, tc).numpy()
            with NoiseTape(seq_tape(spec, torch.from_numpy(eps_seq))):
                out["log_prob_seq_" + tag] = flow.log_prob_sequential(tb, tx, tc).numpy()
            ns = min(3, T - 1)
            with NoiseTape(seq_tape(spec, torch.from_numpy(eps_seq), ns)):
                out["log_prob_seq_n3_" + tag] = flow.log_prob_sequential(tb, tx, tc, n_step=ns).numpy()

    # ---- (2) dense grid, prefix modes self / shared -----------------------------------------
    grid = synth.make_grid(grid_n, lo, hi)
    G = grid.shape[0]
    eps_g = synth.make_eps(spec, agents * G, seed + 5, "eps_grid")
    out.update(grid=grid, eps_grid=eps_g)
    for tag, flow, dt in (("32", flow32, torch.float32), ("64", flow64, torch.float64)):
        tb = torch.from_numpy(np.repeat(inp["base_pos"], G, axis=0)).to(dt)
        tc = torch.from_numpy(np.repeat(inp["cond"], G, axis=0)).to(dt)
        tg = torch.from_numpy(np.tile(grid, (agents, 1))).to(dt)
        tx = tg[:, None, :].expand(-1, T, -1)
        with torch.no_grad():
            with NoiseTape(sep_tape(spec, torch.from_numpy(eps_g))):
                out["grid_self_" + tag] = flow.log_prob(tb, tx, tc).numpy().reshape(agents, G, T)
            # shared: loop i<T over net[i] + base_dist(prev_i, step=i)   
            res = []
            ct = torch.from_numpy(inp["cond_traj"]).to(dt)
            ca = torch.from_numpy(inp["cond"]).to(dt)
            ba = torch.from_numpy(inp["base_pos"]).to(dt)
            with NoiseTape(sep_tape(spec, torch.from_numpy(eps_g))):
                for i in range(T):
                    y = torch.cat([ca[:, i], ct[:, :i].reshape(agents, -1)], dim=-1).repeat_interleave(G, dim=0)
                    w, l = flow.net[i](tg, y)
                    prev = (ba if i == 0 else ct[:, i - 1]).repeat_interleave(G, dim=0)
                    res.append(flow.base_dist(prev, step=i).log_prob(w).sum(-1) + l)
            out["grid_shared_" + tag] = torch.stack(res, dim=1).numpy().reshape(agents, G, T)

    # ---- (3) sampler + rapid update through the reference wrapper ----------------------------
    B = min(agents, 3)
    z0 = synth.normal(seed + 6, "z0", (B, S, D)).astype(np.float32)
    eps_s = synth.make_eps(spec, B * S, seed + 7, "eps_sample")
    gt_st = (inp["base_pos"][:B, None, :] + 0.15 * synth.normal(seed + 8, "gt", (B, T, D))).astype(np.float32)
    out.update(z0=z0, eps_sample=eps_s, gt_st=gt_st)
    for tag, flow, dt in (("32", flow32, torch.float32), ("64", flow64, torch.float64)):
        seqs, lps, ldjs, bases = [], [], [], []
        upd = {1: [], T // 2: [], T - 1: []}
        for b in range(B):   # the reference wrapper only works at B = 1
            model = object.__new__(ref.fastpredNF_TP)
            torch.nn.Module.__init__(model)
            model.flow = flow
            model.pred_len = T
            model.pred_state = "state_v"
            cond_b = torch.from_numpy(inp["cond"][b:b + 1]).to(dt)
            model.encoder = lambda d, c=cond_b: c
            obs_st = torch.zeros(1, 8, 6, dtype=dt)
            obs_st[0, -1, 2:4] = torch.from_numpy(inp["base_pos"][b]).to(dt)
            dd = {"obs_st": obs_st, "gt_st": torch.from_numpy(gt_st[b:b + 1]).to(dt)}
            # predict_inverse_prob hard-codes sample_num = 10000 (fastpredNF.py:104); run the same
            # statements with S samples so the fixture stays small.
            with torch.no_grad():
                base_pos = model.get_base_pos(dd)[:, None].expand(-1, S, -1).clone()
                dist_args = model.encoder(dd)[:, None].expand(-1, S, -1, -1)
                eb = torch.from_numpy(eps_s[b * S:(b + 1) * S])
                tape = [e[None] for e in sep_tape(spec, eb)]
                with NoiseTape(tape, torch.from_numpy(z0[b:b + 1])):
                    seq, lp, lj, bl = flow.sample_with_log_prob(base_pos, cond=dist_args)
                model.ldjs_tmp = lj
                model.base_prob_normalizer_tmp = bl.exp().sum(dim=1)
                dd[("prob_st", 0)] = torch.cat([seq, lp[..., None]], dim=-1)
                dd[("pred_st", 0)] = seq[:, -1]
                for t in upd:
                    r = model.predict_from_new_obs(dd, t)
                    upd[t].append(r[("prob_st", t)][0].numpy())
            seqs.append(seq[0].numpy()); lps.append(lp[0].numpy()); ldjs.append(lj[0].numpy()); bases.append(bl[0].numpy())
        out["sample_seq_" + tag] = np.stack(seqs)
        out["sample_log_prob_" + tag] = np.stack(lps)
        out["sample_ldjs_" + tag] = np.stack(ldjs)
        out["sample_base_" + tag] = np.stack(bases)
        for t, v in upd.items():
            out[f"update_t{t}_{tag}"] = np.stack(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, {k: v.shape for k, v in out.items() if hasattr(v, "shape")}, os.path.getsize(path) // 1024, "KiB")


CASES = {
    # default ETH model (config 1 shapes, small row counts so the fixture stays small)
    "default_eth": dict(spec=synth.FlowSpec(), seed=0, agents=8, grid_n=6, S=64),
    # odd shapes: exercises padding / genericity of the fp32 path (Optuna ranges, main.py:222-229)
    "odd_small": dict(spec=synth.FlowSpec(pred_len=5, cond_size=9, hidden_size=40, n_hidden=1, n_blocks=2), seed=10, agents=5, grid_n=5, S=48),
    # deeper chain (config 5 knobs at reduced width): n_blocks 6
    "deep_blocks": dict(spec=synth.FlowSpec(pred_len=4, cond_size=16, hidden_size=64, n_hidden=2, n_blocks=6), seed=20, agents=4, grid_n=4, S=32),
    # config 5 stress knobs: 2x coupling layers, 256-wide conditioner MLPs (short chain so the fixture stays small)
    "stress_wide": dict(spec=synth.FlowSpec(pred_len=3, cond_size=16, hidden_size=256, n_hidden=2, n_blocks=6), seed=30, agents=3, grid_n=4, S=32),
    # corner of the reference's own tuning range (src/main.py:222-229): HIDDEN_SIZE 128, CONDITIONING_LENGTH 64, N_HIDDEN 3
    "tuned_h128_c64": dict(spec=synth.FlowSpec(pred_len=12, cond_size=64, hidden_size=128, n_hidden=3, n_blocks=3), seed=40, agents=4, grid_n=4, S=24),
    # BASELINE config 5 at full model size: 12 steps x 6 coupling blocks x 256-wide conditioner MLPs (20 M parameters)
    "config5_full": dict(spec=synth.FlowSpec(pred_len=12, cond_size=16, hidden_size=256, n_hidden=2, n_blocks=6), seed=50, agents=3, grid_n=3, S=16),
}


def main():
    ref = import_reference()
    torch.set_num_threads(4)
    only = sys.argv[1:]
    for name, kw in CASES.items():
        if only and name not in only:
            continue
        make_case(ref, name, kw["spec"], seed=kw["seed"], agents=kw["agents"], grid_n=kw["grid_n"], S=kw["S"])


if __name__ == "__main__":
    main()
