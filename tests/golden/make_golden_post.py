"""Generate tests/golden/post_metrics.npz by running the UNMODIFIED reference post-processing on synthetic outputs:
  TP_metrics.denormalize -> unstandardize + output_to_trajectories  (src/metrics/TP_metrics.py:62-112)
  TP_metrics.__call__ / displacement_error / final_displacement_error / evaluate_helper  (:34-60, 163-191)
Run in the build container only (needs /root/reference):   python tests/golden/make_golden_post.py

TP_metrics.__init__ needs the processed dataset pickle (not shipped) and CUDA, so the instance is built with
object.__new__ and given `state`, `std`, `dataset` directly; `ot` (POT, absent) and the dataset module it imports
`hypers` from are stubbed in sys.modules -- none of the exercised methods touches them.
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
sys.path.insert(0, HERE)

from flowchain_iccv2023_b200 import synth  # noqa: E402
import make_golden  # noqa: E402


def import_metrics():
    make_golden.import_reference()          # yacs stub + sys.path
    for name in ("ot", "data", "data.TP", "data.TP.trajectron_dataset"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["data.TP.trajectron_dataset"].hypers = {}
    import metrics.TP_metrics as tpm  # noqa
    return tpm


def make_metrics(tpm, state, std):
    m = object.__new__(tpm.TP_metrics)
    m.state = state
    m.std = torch.tensor(std, dtype=torch.float32)
    m.dataset = "zara1"          # no dataset-specific rescaling of ADE/FDE
    return m


def main():
    tpm = import_metrics()
    B, T, N, S = 5, 12, 6, 40
    std = np.array([1.7, 2.3], dtype=np.float32)
    dt = np.full((B,), 0.4, dtype=np.float32)
    obs = synth.normal(1, "obs", (B, 8, 6)).astype(np.float32) * 3.0
    gt_st = (0.3 * synth.normal(2, "gt_st", (B, T, 2))).astype(np.float32)
    pred_st = (0.3 * synth.normal(3, "pred_st", (N, B, T, 2))).astype(np.float32)
    out = dict(std=std, dt=dt, obs=obs, gt_st=gt_st, pred_st=pred_st)
    for tag, state in (("v", {"velocity": ["x", "y"]}), ("p", {"position": ["x", "y"]})):
        m = make_metrics(tpm, state, std)
        gt_un = gt_st * std if tag == "v" else gt_st * std + obs[:, -1:, 0:2]      # what the loader provides as `gt`
        dict_list = []
        for j in range(N):
            dict_list.append({"obs": torch.from_numpy(obs), "gt": torch.from_numpy(gt_un.copy()), "gt_st": torch.from_numpy(gt_st),
                              "dt": torch.from_numpy(dt), ("pred_st", 0): torch.from_numpy(pred_st[j])})
        dict_list = m.denormalize(dict_list)
        out[f"pred_{tag}"] = np.stack([d[("pred", 0)].numpy() for d in dict_list])
        out[f"gt_pos_{tag}"] = dict_list[0]["gt"].numpy()
        res = m(dict_list)
        out[f"ade_sum_{tag}"] = np.asarray(res["ade"], dtype=np.float32)
        out[f"fde_sum_{tag}"] = np.asarray(res["fde"], dtype=np.float32)
        ade = torch.stack([tpm.displacement_error(d[("pred", 0)][:, -12:], d["gt"][:, -12:]) for d in dict_list])
        fde = torch.stack([tpm.final_displacement_error(d[("pred", 0)][:, -1], d["gt"][:, -1]) for d in dict_list])
        out[f"ade_trials_{tag}"] = ade.numpy()
        out[f"fde_trials_{tag}"] = fde.numpy()
        # sample clouds ("prob_st", k): the reference's broadcasting only works at batch size 1 (main.py uses a B = 1 loader)
        probs0, probs3 = [], []
        for b in range(3):
            p0 = np.concatenate([0.4 * synth.normal(10 + b, "seq", (1, S, T, 2)), 4.0 * synth.normal(20 + b, "lp", (1, S, T, 1)) - 2.0], -1).astype(np.float32)
            p3 = np.concatenate([0.4 * synth.normal(30 + b, "seq3", (1, S, T - 3, 2)), 4.0 * synth.normal(40 + b, "lp3", (1, S, T - 3, 1)) - 2.0], -1).astype(np.float32)
            d = {"obs": torch.from_numpy(obs[b:b + 1]), "gt": torch.from_numpy(gt_un[b:b + 1].copy()), "gt_st": torch.from_numpy(gt_st[b:b + 1]),
                 "dt": torch.from_numpy(dt[b:b + 1]), ("pred_st", 0): torch.from_numpy(pred_st[0, b:b + 1]),
                 ("prob_st", 0): torch.from_numpy(p0), ("prob_st", 3): torch.from_numpy(p3)}
            d = m.denormalize([d])[0]
            probs0.append(d[("prob", 0)].numpy()[0])
            probs3.append(d[("prob", 3)].numpy()[0])
            if tag == "v":
                out.setdefault("prob_st_0", []).append(p0[0])
                out.setdefault("prob_st_3", []).append(p3[0])
        out[f"prob_0_{tag}"] = np.stack(probs0)
        out[f"prob_3_{tag}"] = np.stack(probs3)
    out["prob_st_0"] = np.stack(out["prob_st_0"])
    out["prob_st_3"] = np.stack(out["prob_st_3"])
    path = os.path.join(HERE, "post_metrics.npz")
    np.savez_compressed(path, **out)
    print({k: np.asarray(v).shape for k, v in out.items()}, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
