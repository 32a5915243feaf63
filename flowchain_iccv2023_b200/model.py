"""Host-side mirror of the reference model wrapper ``fastpredNF_TP`` for the density hot path
(reference: src/models/TP/fastpredNF.py:21-243; model protocol src/models/build_model.py:22-39).

Kept: ``predict(data_dict, return_prob)``, ``predict_from_new_obs(data_dict, time_step)``,
``get_base_pos``, ``save / load / check_saved_path``, the ``data_dict`` keys
(``("prob_st", k)`` (B,S,T-k,3), ``("pred_st", k)`` (B,T-k,2)) and the module-level update cache
(``ldjs_tmp``, ``base_prob_normalizer_tmp``).  Added: a working ``predict_forward`` (dense grid).
The history encoder is NOT part of this path (north_star (d)): pass any callable
``encoder(data_dict) -> (B, T, C0)`` (a PyTorch module in production, a stub in tests/bench).
``update()`` is the reference's Adam training step on the flow's differentiable PyTorch path
(flow.py, train() mode); the fused kernels re-pack the weights automatically afterwards.
"""
from __future__ import annotations

from pathlib import Path
from typing import Callable, Dict, Optional

import torch
import torch.nn as nn

from .flow import fastpredNF_CIF_separate_cond


class fastpredNF_TP(nn.Module):
    def __init__(self, cfg=None, encoder: Optional[Callable] = None, pred_len: int = 12, obs_len: int = 8,
                 pred_state: str = "state_v", n_blocks: int = 3, n_hidden: int = 2, hidden_size: int = 64,
                 cond_size: int = 16, output_path: Optional[str] = None, lr: float = 1e-3, weight_decay: float = 0.0,
                 dequantize: bool = False):
        super().__init__()
        if cfg is not None:   # duck-typed yacs node (src/default_params.py)
            lr = cfg.SOLVER.LR
            weight_decay = cfg.SOLVER.WEIGHT_DECAY
            dequantize = cfg.SOLVER.DEQUANTIZE
            pred_len = cfg.DATA.PREDICT_LENGTH
            obs_len = cfg.DATA.OBSERVE_LENGTH
            pred_state = cfg.DATA.TP.PRED_STATE
            n_blocks = cfg.MODEL.FLOW.N_BLOCKS
            n_hidden = cfg.MODEL.FLOW.N_HIDDEN
            hidden_size = cfg.MODEL.FLOW.HIDDEN_SIZE
            cond_size = cfg.MODEL.FLOW.CONDITIONING_LENGTH
            output_path = cfg.OUTPUT_DIR
        self.output_path = Path(output_path) if output_path is not None else Path("output")
        self.obs_len, self.pred_len, self.pred_state = obs_len, pred_len, pred_state
        self.output_size = 2
        self.encoder = encoder
        self.flow = fastpredNF_CIF_separate_cond(input_size=2, n_blocks=n_blocks, n_hidden=n_hidden,
                                                 hidden_size=hidden_size, cond_label_size=cond_size,
                                                 flow_architecture="realNVP", pred_len=pred_len)
        self.lr, self.weight_decay, self.dequantize = lr, weight_decay, dequantize
        self.optimizer = None            # built on first use so that an nn.Module encoder assigned later is included
        self.optimizers = []
        self.sample_num_prob = 10000     # fastpredNF.py:104
        self.sample_num_ml = 100         # fastpredNF.py:126
        self.ldjs_tmp = None
        self.base_prob_normalizer_tmp = None

    # ---- model protocol -------------------------------------------------------------------------
    def predict(self, data_dict: Dict, return_prob: bool = True, **noise) -> Dict:
        if return_prob:
            return self.predict_inverse_prob(data_dict, **noise)
        return self.predict_inverse_ML(data_dict, **noise)

    def _sample(self, data_dict, sample_num, **noise):
        dist_args = self.encoder(data_dict)                                       # (B, T, C0)
        base_pos = self.get_base_pos(data_dict)[:, None].expand(-1, sample_num, -1)
        dist_args = dist_args[:, None].expand(-1, sample_num, -1, -1)
        return self.flow.sample_with_log_prob(base_pos, cond=dist_args, **noise)

    @torch.no_grad()
    def predict_inverse_prob(self, data_dict: Dict, **noise) -> Dict:
        """fastpredNF.py:101-121."""
        sampled_seq, log_prob, seq_ldjs, base_log_prob = self._sample(data_dict, self.sample_num_prob, **noise)
        self.ldjs_tmp = seq_ldjs                                                  # cache for the rapid update
        h = self.flow.engine()
        self.base_prob_normalizer_tmp = torch.ops.flowchain.base_normalizer(
            h.id, base_log_prob.to(self.flow.var.device).contiguous()).to(base_log_prob.device)
        data_dict[("prob_st", 0)] = torch.cat([sampled_seq, log_prob[..., None]], dim=-1)
        data_dict[("pred_st", 0)] = sampled_seq[:, -1]
        return data_dict

    @torch.no_grad()
    def predict_inverse_ML(self, data_dict: Dict, **noise) -> Dict:
        """fastpredNF.py:123-140."""
        sampled_seq, _, _, _ = self._sample(data_dict, self.sample_num_ml, **noise)
        pred = sampled_seq[:, -1]
        if torch.sum(torch.isnan(pred)):
            pred = torch.where(torch.isnan(pred), data_dict["obs_st"][:, 0, None, 2:4].expand(pred.size()), pred)
        data_dict[("pred_st", 0)] = pred
        return data_dict

    @torch.no_grad()
    def predict_from_new_obs(self, data_dict: Dict, time_step: int) -> Dict:
        """Rapid update (fastpredNF.py:142-164); batched = vmap of the B = 1 reference call over agents."""
        assert 0 < time_step and time_step < self.pred_len, \
            f"time_step for update need to be 0~{self.pred_len-1}, got {time_step}"
        assert ("prob_st", 0) in data_dict.keys(), "Initial prediction needed before updating"
        h = self.flow.engine()
        dev = self.flow.var.device
        prob0 = data_dict[("prob_st", 0)]
        src = prob0.device
        out = torch.ops.flowchain.update(
            h.id, data_dict["gt_st"][:, time_step - 1].to(dev, torch.float32).contiguous(), time_step,
            prob0.to(dev).contiguous(), self.ldjs_tmp.to(dev).contiguous(),
            self.base_prob_normalizer_tmp.to(dev).contiguous())
        data_dict[("prob_st", time_step)] = out.to(src)
        data_dict[("pred_st", time_step)] = data_dict[("pred_st", 0)][:, time_step:]
        return data_dict

    @torch.no_grad()
    def predict_from_new_obs_lazy(self, data_dict: Dict, time_step: int) -> torch.Tensor:
        """Lazy form of the rapid update: returns only base_lp_t (B, S), the per-sample term that the reference adds
        to `ldjs_tmp[:, :, t:]` (fastpredNF.py:150-156); ("prob_st", t)[..., 2] == base_lp_t[..., None] + ldjs_tmp[:, :, t:]
        and the positions are ("prob_st", 0)[:, :, t:, :2] unchanged.  40 KB per agent instead of 1.3 MB."""
        assert 0 < time_step and time_step < self.pred_len, \
            f"time_step for update need to be 0~{self.pred_len-1}, got {time_step}"
        assert ("prob_st", 0) in data_dict.keys(), "Initial prediction needed before updating"
        h = self.flow.engine()
        dev = self.flow.var.device
        prob0 = data_dict[("prob_st", 0)]
        out = torch.ops.flowchain.update_lazy(
            h.id, data_dict["gt_st"][:, time_step - 1].to(dev, torch.float32).contiguous(), time_step,
            prob0.to(dev).contiguous(), self.base_prob_normalizer_tmp.to(dev).contiguous())
        return out.to(prob0.device)

    @torch.no_grad()
    def predict_forward(self, data_dict: Dict, sample_num_per_dim: int = 200, lo: float = -3.0, hi: float = 3.0,
                        cond_traj=None, sequential: bool = False, **kw) -> Dict:
        """Working version of the reference's dead grid evaluation (fastpredNF.py:166-184):
        ("prob_st", 0) = cat[grid, log_prob] of shape (B, G, T, 3) on the lattice linspace(lo,hi,n)^2."""
        dist_args = self.encoder(data_dict)
        base_pos = self.get_base_pos(data_dict)
        dev = self.flow.var.device
        ax = torch.linspace(lo, hi, steps=sample_num_per_dim, device=dev)
        grid = torch.cartesian_prod(ax, ax)
        lp = self.flow.grid_log_prob(base_pos, dist_args, grid, cond_traj=cond_traj, sequential=sequential, **kw)
        B, G, T = lp.shape
        seq = grid[None, :, None, :].expand(B, G, T, 2)
        data_dict[("prob_st", 0)] = torch.cat([seq, lp[..., None]], dim=-1)
        return data_dict

    # ---- consumers right after the hot path -----------------------------------------------
    _STATE_MODE = {"state_p": 0, "state_v": 1}

    def _state_mode(self) -> int:
        if self.pred_state not in self._STATE_MODE:
            raise AssertionError("acceleration states are not supported (the reference asserts the same, TP_metrics.py:70,94)")
        return self._STATE_MODE[self.pred_state]

    def _f32(self, t):
        return None if t is None else torch.as_tensor(t).to(self.flow.var.device, torch.float32).contiguous()

    @torch.no_grad()
    def denormalize(self, data_dict: Dict, std) -> Dict:
        """TP_metrics.denormalize for one data_dict (src/metrics/TP_metrics.py:62-112), fused on the device:
        ("pred", 0) = metric trajectory of ("pred_st", 0); `gt` (un-standardised state) -> metric positions;
        every ("prob_st", k) -> ("prob", k) = [metric positions, p = exp(log p)].  `std` = the dataset's standardisation
        scale (2,), `data_dict["dt"]` (B,), `data_dict["obs"][:, -1, :2]` = last observed metric position."""
        if ("pred", 0) in data_dict:
            return data_dict
        h = self.flow.engine()
        mode = self._state_mode()
        s0, s1 = (float(v) for v in torch.as_tensor(std).flatten()[:2])
        last_obs = self._f32(data_dict["obs"][:, -1, 0:2])
        dt = self._f32(data_dict["dt"]) if mode == 1 else None
        pred = self._f32(data_dict[("pred_st", 0)])
        data_dict[("pred", 0)] = torch.ops.flowchain.to_trajectories(h.id, pred[:, None], last_obs, dt, s0, s1, mode)[:, 0]
        if mode == 1:       # the un-standardised ground-truth velocities become positions (TP_metrics.py:82-83)
            gt_v = self._f32(data_dict["gt"])
            data_dict["gt"] = torch.ops.flowchain.to_trajectories(h.id, gt_v[:, None], last_obs, dt, 1.0, 1.0, 1)[:, 0]
        gt_pos = self._f32(data_dict["gt"])
        for k in [k for k in data_dict if isinstance(k, tuple) and k[0] == "prob_st" and isinstance(data_dict[k], torch.Tensor)]:
            off = last_obs if k[1] == 0 else gt_pos[:, k[1] - 1].contiguous()
            data_dict[("prob", k[1])] = torch.ops.flowchain.to_trajectories(h.id, self._f32(data_dict[k]), off, dt, s0, s1, mode)
        return data_dict

    @torch.no_grad()
    def prob_to_grid(self, data_dict: Dict, xx, yy, std, cond_traj="gt", as_prob: bool = True, **kw) -> Dict:
        """Drop-in for TP_visualizer.prob_to_grid on FlowChain outputs (src/visualization/TP_visualizer.py:118-179): sets
        `grid`, ("prob", 0) = (B, G, T) density on the metric lattice (xx, yy) (flattened like `zz.reshape(-1, T)`) and
        ("gt_traj_log_prob", 0) (B, T).  The reference scatters 10 000 sampled (position, p) pairs per step and
        interpolates them on the CPU (single-linkage clustering + griddata + RBF), then reads the ground truth off the
        lattice with interp2d; here the analytic per-step density is evaluated directly: on the lattice, conditioned on
        a prefix trajectory (`cond_traj`: "gt" = teacher forcing with data_dict["gt_st"], or a (B,T,2) standardised
        tensor such as the ML prediction), and exactly at the ground truth.  Both are proper densities in metric space."""
        h = self.flow.engine()
        dev = self.flow.var.device
        mode = self._state_mode()
        stdv = torch.as_tensor(std, dtype=torch.float32, device=dev).flatten()[:2]
        dist_args = self._f32(self.encoder(data_dict))
        base_pos = self._f32(self.get_base_pos(data_dict))
        traj = self._f32(data_dict["gt_st"] if isinstance(cond_traj, str) else cond_traj)
        B, T = traj.shape[0], traj.shape[1]
        last_obs = self._f32(data_dict["obs"][:, -1, 0:2])
        if mode == 1:
            dt = self._f32(data_dict["dt"])
            step = traj * stdv * dt[:, None, None]                                  # metric displacement of every step
            origin = last_obs[:, None] + torch.cumsum(step, dim=1) - step           # position BEFORE step t
            inv_scale = 1.0 / (stdv[None] * dt[:, None])
        else:
            origin = last_obs[:, None].expand(B, T, 2)
            inv_scale = (1.0 / stdv)[None].expand(B, 2)
        log_jac = torch.log(inv_scale).sum(-1)                                      # log |d state / d position|
        grid_xy = torch.stack([torch.as_tensor(xx, dtype=torch.float32).flatten(),
                               torch.as_tensor(yy, dtype=torch.float32).flatten()], dim=-1).to(dev).contiguous()
        eps, seed = kw.pop("eps", None), kw.pop("seed", None)
        prec = kw.pop("precision", None)
        maps = torch.ops.flowchain.grid_density_maps(
            h.id, base_pos, dist_args, traj, grid_xy, origin.contiguous(), inv_scale.contiguous(), log_jac.contiguous(),
            self._f32(eps), 0 if eps is not None else self.flow._seed(seed), bool(as_prob), False,
            self.flow.precision if prec is None else prec)
        data_dict["grid"] = [xx, yy]
        data_dict[("prob", 0)] = maps
        if "gt_st" in data_dict:
            lp = self.flow.log_prob(base_pos, self._f32(data_dict["gt_st"]), dist_args, eps=kw.pop("eps_gt", None), seed=seed,
                                    precision=prec)
            data_dict[("gt_traj_log_prob", 0)] = lp + log_jac[:, None]
        return data_dict

    @torch.no_grad()
    def predict_best_of_n(self, data_dict: Dict, std, n_trial: int = 20, **noise) -> Dict:
        """The ADE/FDE best-of-N evaluation of src/main.py:162-180 in one go: ONE sampler launch with S = n_trial
        candidates per agent (the reference runs n_trial predict() calls of 100 samples each and keeps the last sample
        of each, fastpredNF.py:123-140) + an on-device un-standardise / cumsum / min-reduction.  `data_dict["gt"]` must
        hold METRIC ground-truth positions (B,T,2).  Returns per-agent minima and their sums (evaluate_helper)."""
        h = self.flow.engine()
        mode = self._state_mode()
        s0, s1 = (float(v) for v in torch.as_tensor(std).flatten()[:2])
        seq, _, _, _ = self._sample(data_dict, n_trial, **noise)                    # (B, n_trial, T, 2)
        fallback = self._f32(data_dict["obs_st"][:, 0, 2:4])
        ade, fde, ia, jf = torch.ops.flowchain.best_of_n(
            h.id, self._f32(seq), self._f32(data_dict["gt"]), self._f32(data_dict["obs"][:, -1, 0:2]),
            self._f32(data_dict["dt"]) if mode == 1 else None, fallback, s0, s1, mode)
        return {"ade": ade.sum(), "fde": fde.sum(), "ade_per_agent": ade, "fde_per_agent": fde, "ade_arg": ia, "fde_arg": jf,
                "nsample": ade.shape[0], "candidates_st": seq}

    def gt_log_prob(self, data_dict: Dict, **kw) -> torch.Tensor:
        """Exact log p of the ground-truth future under the per-step density: flow.log_prob(base_pos, gt_st, cond) -> (B, T).
        Replaces the visualiser's detour of interpolating a KDE of samples on a grid and reading it off at the GT
        (src/visualization/TP_visualizer.py:165-179): the density is analytic, so it is evaluated."""
        dist_args = self.encoder(data_dict)
        base_pos = self.get_base_pos(data_dict)
        return self.flow.log_prob(base_pos, data_dict["gt_st"], dist_args, **kw)

    def build_optimizer(self):
        """Adam with weight decay on everything but biases (fastpredNF.py:72-89)."""
        named = list(self.named_parameters())
        groups = [{"params": [p for n, p in named if "bias" not in n], "weight_decay": self.weight_decay},
                  {"params": [p for n, p in named if "bias" in n], "weight_decay": 0.0}]
        self.optimizer = torch.optim.Adam(groups, self.lr)
        self.optimizers = [self.optimizer]
        return self.optimizer

    def update(self, data_dict: Dict) -> Dict:
        """Training step (fastpredNF.py:186-202): Adam on the mean negative per-step log-density of the ground-truth
        future.  Runs the flow's differentiable PyTorch path (flow.py); BatchNorm uses batch statistics in train() mode
        exactly as the reference's layer does.  The packed weights of the fused kernels are re-built on the next
        inference call."""
        if self.optimizer is None:
            self.build_optimizer()
        with torch.enable_grad():
            dist_args = self.encoder(data_dict)
            gt = data_dict["gt_st"]
            if self.dequantize:
                gt += torch.rand_like(gt) / 100
            base_pos = self.get_base_pos(data_dict)
            loss = -self.flow._log_prob_separate_torch(base_pos, gt, dist_args)
            loss = loss.mean()
            self.optimizer.zero_grad()
            loss.backward()
            self.optimizer.step()
        self.flow.invalidate()
        return {"loss": loss.mean().item()}

    def get_base_pos(self, data_dict: Dict) -> torch.Tensor:
        """fastpredNF.py:204-213."""
        if self.pred_state == "state_p":
            return data_dict["obs_st"][:, -1, :2]
        elif self.pred_state == "state_v":
            return data_dict["obs_st"][:, -1, 2:4]
        elif self.pred_state == "state_a":
            return data_dict["obs_st"][:, -1, 4:6]
        raise ValueError

    def save(self, epoch: int = 0, path: Path = None) -> None:
        path = self.output_path / "ckpt.pt" if path is None else path
        torch.save({"epoch": epoch, "state": self.state_dict(),
                    "optim_state": self.optimizer.state_dict() if self.optimizer is not None else {}}, path)

    def check_saved_path(self, path: Path = None) -> bool:
        path = self.output_path / "ckpt.pt" if path is None else path
        return Path(path).exists()

    def load(self, path: Path = None) -> int:
        path = self.output_path / "ckpt.pt" if path is None else path
        ckpt = torch.load(path, map_location="cpu")
        # the reference loads strictly (fastpredNF.py:237-238).  Here the encoder is a pluggable callable that may or may
        # not be an nn.Module registered on this wrapper, so only `encoder.*` keys may be absent / extra; every `flow.*`
        # key must match, otherwise the kernels would silently run on random-init weights.
        missing, unexpected = self.load_state_dict(ckpt["state"], strict=False)
        bad = [k for k in list(missing) + list(unexpected) if not k.startswith("encoder.")]
        if bad:
            raise RuntimeError(f"checkpoint {path} does not match this model: missing/unexpected keys {bad[:8]}"
                               f"{' ...' if len(bad) > 8 else ''} ({len(bad)} in total)")
        self.flow.invalidate()
        if ckpt.get("optim_state"):      # optimizer state travels with the checkpoint (fastpredNF.py:240-241)
            if self.optimizer is None:
                self.build_optimizer()
            self.optimizer.load_state_dict(ckpt["optim_state"])
        return ckpt["epoch"]
