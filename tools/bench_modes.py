"""Timing of the secondary entry points at config-2 shape: chain mode (log_prob_sequential over the lattice) and the sampler,
tensor-core (f16x3) vs fp32 CUDA-core.  Prints one JSON line."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowchain_iccv2023_b200 import synth, _lib
from flowchain_iccv2023_b200.flow import fastpredNF_CIF_separate_cond


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    spec = synth.FlowSpec(); p = synth.make_params(spec, 0)
    f = fastpredNF_CIF_separate_cond(2, spec.n_blocks, spec.hidden_size, spec.n_hidden, spec.cond_size, flow_architecture="realNVP", pred_len=spec.pred_len)
    f.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()}); f = f.cuda().eval(); f.check_versions = False
    res = {}
    inp = synth.make_inputs(spec, 64, seed=1); grid = torch.from_numpy(synth.make_grid(256)).cuda()
    b = torch.from_numpy(inp["base_pos"]).cuda(); c = torch.from_numpy(inp["cond"]).cuda()
    for prec, name, reps in ((_lib.FC_PREC_F16X3, "f16x3", 3), (_lib.FC_PREC_FP32, "fp32", 1)):
        ms = timed(lambda: f.grid_log_prob(b, c, grid, sequential=True, seed=1, precision=prec, map_layout=True), reps)
        res[f"chain_config2_{name}_ms"] = ms
        res[f"chain_config2_{name}_tflops"] = 64 * 65536 * spec.flops_sequential_row() / ms / 1e9
    for B, S in ((1, 10000), (128, 100), (128, 10000)):
        inp = synth.make_inputs(spec, B, seed=1)
        bp = torch.from_numpy(inp["base_pos"]).cuda()[:, None].expand(-1, S, -1); cd = torch.from_numpy(inp["cond"]).cuda()[:, None].expand(-1, S, -1, -1)
        for prec, name in ((_lib.FC_PREC_F16X3, "f16x3"), (_lib.FC_PREC_FP32, "fp32")):
            ms = timed(lambda: f.sample_with_log_prob(bp, cd, seed=1, precision=prec), 5)
            res[f"sampler_B{B}_S{S}_{name}_ms"] = ms
            res[f"sampler_B{B}_S{S}_{name}_Mevals_s"] = B * S * spec.pred_len / ms / 1e3
    print(json.dumps(res))


if __name__ == "__main__":
    main()
