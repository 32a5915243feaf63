"""Encoder under CUDA graphs: GraphedEncoder replays a PyTorch history encoder bit-identically to its
eager run, per input-shape signature, stays eager (and differentiable) in training, and plugs into fastpredNF_TP.predict."""
import numpy as np
import pytest
import torch

from flowchain_iccv2023_b200 import synth

pytestmark = pytest.mark.gpu


def test_graphed_encoder_matches_eager_and_feeds_predict():
    from flowchain_iccv2023_b200.encoder import GraphedEncoder, HistoryEncoder
    from flowchain_iccv2023_b200.model import fastpredNF_TP
    torch.manual_seed(0)
    enc = HistoryEncoder(6, 32, 16, 12).cuda().eval()
    genc = GraphedEncoder(enc).cuda().eval()
    with torch.no_grad():
        for B in (128, 37, 128):                                   # two shape signatures, the first one reused
            obs = torch.randn(B, 8, 6, device="cuda")
            ref = enc({"obs_st": obs})
            out = genc({"obs_st": obs})
            assert out.shape == (B, 12, 16) and torch.equal(out, ref)
            obs2 = torch.randn(B, 8, 6, device="cuda")
            assert torch.equal(genc({"obs_st": obs2}), enc({"obs_st": obs2}))     # replay with new inputs
    assert len(genc._graphs) == 2
    # latency of one encode: graph replay vs eager (informational)
    obs = torch.randn(128, 8, 6, device="cuda")
    def p50(fn):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(200)]
        with torch.no_grad():
            for a, b in ev:
                a.record(); fn({"obs_st": obs}); b.record()
        torch.cuda.synchronize()
        return float(np.median([a.elapsed_time(b) for a, b in ev])) * 1e3
    print(f"encoder p50: eager {p50(enc):.1f} us, CUDA graph {p50(genc):.1f} us")
    # training: eager + differentiable
    genc.train()
    out = genc({"obs_st": obs})
    assert out.requires_grad
    out.sum().backward()
    assert all(q.grad is not None for q in enc.parameters())
    genc.eval()
    # the wrapper takes it like any encoder
    spec = synth.DEFAULT_SPEC
    m = fastpredNF_TP(encoder=genc, pred_len=spec.pred_len)
    m.flow.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_params(spec, 0).items()}, strict=True)
    m = m.cuda().eval()
    m.sample_num_prob = 256
    dd = m.predict({"obs_st": torch.randn(4, 8, 6, device="cuda") * 0.3}, return_prob=True, seed=1)
    assert dd[("prob_st", 0)].shape == (4, 256, spec.pred_len, 3)
