"""Config 3: p50 latency of the rapid density update (predict_from_new_obs) and of the sampler that fills its
cache, CUDA-event timed, update launched from a captured CUDA graph (>= 1000 replays).

    python tools/bench_update.py            # prints one JSON line per case
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowchain_iccv2023_b200 import synth
from flowchain_iccv2023_b200.flow import fastpredNF_CIF_separate_cond


def main():
    spec = synth.DEFAULT_SPEC
    f = fastpredNF_CIF_separate_cond(2, 3, 64, 2, 16, flow_architecture="realNVP", pred_len=12)
    f.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_params(spec, 0).items()})
    f = f.cuda().eval()
    f.check_versions = False
    h = f.engine()
    S, T = 10000, spec.pred_len
    agents = tuple(int(a) for a in sys.argv[1:]) or (1, 64, 1000)
    for A in agents:
        inp = synth.make_inputs(spec, A, seed=3)
        bp = torch.from_numpy(inp["base_pos"]).cuda()
        cd = torch.from_numpy(inp["cond"]).cuda()
        # sampler (fills the cache): B agents x S samples, 12 autoregressive CIF steps
        for _ in range(2):
            seq, lp, ldjs, base = torch.ops.flowchain.sample_with_log_prob(h.id, bp, cd, S, None, None, 7, 2 if A > 100 else 0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3 if A > 100 else 10
        e0.record()
        for _ in range(reps):
            seq, lp, ldjs, base = torch.ops.flowchain.sample_with_log_prob(h.id, bp, cd, S, None, None, 7, 2 if A > 100 else 0)
        e1.record()
        torch.cuda.synchronize()
        sample_ms = e0.elapsed_time(e1) / reps
        prob0 = torch.cat([seq, lp[..., None]], dim=-1).contiguous()
        norm = torch.ops.flowchain.base_normalizer(h.id, base.contiguous())
        new_pos = (bp + 0.05).contiguous()
        res = {"agents": A, "samples": S, "sampler_ms": sample_ms, "sampler_evals_per_s": A * S * T / (sample_ms / 1e3)}
        for t in (1, 6, 11):
            out = torch.ops.flowchain.update(h.id, new_pos, t, prob0, ldjs, norm)   # warm-up / allocation
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for _ in range(3):
                    torch.ops.flowchain.update(h.id, new_pos, t, prob0, ldjs, norm)
                torch.cuda.synchronize()
                with torch.cuda.graph(g, stream=s):
                    out = torch.ops.flowchain.update(h.id, new_pos, t, prob0, ldjs, norm)
            torch.cuda.synchronize()
            times = []
            n_rep = 1000
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_rep)]
            for a, b in evs:
                a.record()
                g.replay()
                b.record()
            torch.cuda.synchronize()
            times = np.array([a.elapsed_time(b) for a, b in evs]) * 1e3
            algo_bytes = A * (8 * S + 8 * S * (T - t))      # read positions + ldjs, write log-probs
            moved = A * S * ((T - t) * 3 * 4 * 2 + 12 + (T - t) * 4)   # what the kernel actually moves (prob_st copy included)
            res[f"update_t{t}_p50_us"] = float(np.median(times))
            res[f"update_t{t}_p99_us"] = float(np.quantile(times, 0.99))
            res[f"update_t{t}_GBps_moved"] = moved / (np.median(times) * 1e-6) / 1e9
        print(json.dumps(res))


if __name__ == "__main__":
    main()
