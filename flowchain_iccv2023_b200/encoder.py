"""History encoder side of the wrapper (north_star (d): the encoder stays in PyTorch).

The reference's production encoder (`Trajectron_encoder`, src/models/TP/fastpredNF.py:311-377 -> mgcvae.get_latent,
src/models/TP/mgcvae.py:962-1016: node-history LSTM, edge LSTMs, additive attention) needs the processed dataset pickle
and is not part of this path; `fastpredNF_TP` accepts any callable ``encoder(data_dict) -> (B, T, C0)``.  This module
provides the piece that matters once the flow is fused and the encoder's ~100 small eager launches become the latency
floor of `predict()`:

* ``GraphedEncoder`` -- runs ANY PyTorch encoder module under a CUDA graph: the first inference call for a given set of
  input shapes captures the encoder's whole launch sequence into a graph with static input / output buffers, later calls
  copy the inputs in and replay it (one launch instead of one per op).  Training (grad enabled, or train() mode) always
  takes the eager module, so `update()` differentiates through the encoder as before.
* ``HistoryEncoder`` -- a small stand-in with the reference encoders' interface and output convention (an LSTM over the
  standardised observation history, projected to C0 and repeated over the pred_len steps like fastpredNF.py:374-375),
  used by the tests and examples; it is not the Trajectron++ CVAE.
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import torch
import torch.nn as nn


class HistoryEncoder(nn.Module):
    """obs_st (B, obs_len, F) -> dist_args (B, pred_len, C0): LSTM over the history, last hidden state projected and
    repeated over the prediction steps (the reference encoders hand the flow one vector per agent, fastpredNF.py:374-375)."""

    def __init__(self, input_size: int = 6, hidden_size: int = 32, output_size: int = 16, pred_len: int = 12):
        super().__init__()
        self.pred_len = pred_len
        self.lstm = nn.LSTM(input_size, hidden_size, batch_first=True)
        self.dist_args_proj = nn.Linear(hidden_size, output_size)

    def forward(self, data_dict: Dict) -> torch.Tensor:
        _, (h, _) = self.lstm(data_dict["obs_st"])
        return self.dist_args_proj(h[-1])[:, None].expand(-1, self.pred_len, -1)


class GraphedEncoder(nn.Module):
    """CUDA-graph replay of a PyTorch encoder for inference.

    `keys` are the data_dict entries the encoder reads (tensors on the encoder's device); one graph is captured per
    distinct tuple of their shapes / dtypes (batch sizes in test mode are few: BATCH_SIZE and the last partial batch).
    The returned tensor is the graph's static output buffer: it is overwritten by the next call with the same shapes,
    exactly like any other CUDA-graph output -- consume (or clone) it before encoding the next batch."""

    def __init__(self, encoder: nn.Module, keys: Sequence[str] = ("obs_st",), warmup: int = 3):
        super().__init__()
        self.encoder = encoder
        self.keys = tuple(keys)
        self.warmup = warmup
        self._graphs: Dict[Tuple, Tuple] = {}

    def forward(self, data_dict: Dict) -> torch.Tensor:
        if self.training or torch.is_grad_enabled():
            return self.encoder(data_dict)                       # training path: eager and differentiable
        ins = [data_dict[k] for k in self.keys]
        if not all(t.is_cuda for t in ins):
            return self.encoder(data_dict)
        sig = tuple((tuple(t.shape), t.dtype, t.device) for t in ins)
        entry = self._graphs.get(sig)
        if entry is None:
            static_in = [torch.empty_like(t).copy_(t) for t in ins]
            static_dict = dict(data_dict)
            static_dict.update(dict(zip(self.keys, static_in)))
            side = torch.cuda.Stream(device=ins[0].device)
            side.wait_stream(torch.cuda.current_stream(ins[0].device))
            with torch.cuda.stream(side):                        # warm-up off the capture: lazy initialisations (cuDNN plans ...)
                for _ in range(self.warmup):
                    self.encoder(static_dict)
            torch.cuda.current_stream(ins[0].device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self.encoder(static_dict)
            entry = (graph, static_in, static_out)
            self._graphs[sig] = entry
        graph, static_in, static_out = entry
        for dst, src in zip(static_in, ins):
            dst.copy_(src, non_blocking=True)
        graph.replay()
        return static_out

    def invalidate(self):
        """Drop the captured graphs (after the encoder's parameters were re-allocated, e.g. by load_state_dict(assign=True))."""
        self._graphs.clear()
