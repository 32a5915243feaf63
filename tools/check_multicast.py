"""torchrun check (N >= 2 GPUs of one NVSwitch box): density maps assembled by multicast stores (fused all-gather) are
bit-identical to the NCCL all-gather of the same shards; also times both.  Usage:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_multicast.py"""
import os, sys, torch, numpy as np, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowchain_iccv2023_b200 import synth, _lib
from flowchain_iccv2023_b200 import dist as fdist
from flowchain_iccv2023_b200.flow import fastpredNF_CIF_separate_cond


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    spec = synth.FlowSpec(); p = synth.make_params(spec, 0)
    f = fastpredNF_CIF_separate_cond(2, spec.n_blocks, spec.hidden_size, spec.n_hidden, spec.cond_size, flow_architecture="realNVP", pred_len=spec.pred_len)
    f.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()}); f = f.to(dev).eval(); f.precision = _lib.FC_PREC_F16X3
    ok = True
    for A, n in ((6, 40), (1, 64), (16 * world, 128)):          # agents sharded / points sharded / larger
        inp = synth.make_inputs(spec, A, seed=3)
        bp = torch.from_numpy(inp["base_pos"]).to(dev); cd = torch.from_numpy(inp["cond"]).to(dev); g = torch.from_numpy(synth.make_grid(n)).to(dev)
        ref = fdist.sharded_grid_log_prob(f, bp, cd, g, seed=7)
        mm = fdist.MulticastMaps(A, spec.pred_len, n * n, dev)
        out = mm.fill(f, bp, cd, g, seed=7)
        same = torch.equal(out, ref)
        ok &= same
        if rank == 0: print(f"A={A} grid={n}x{n}: multicast == nccl all-gather: {same}", flush=True)
    # timing at config-2 shard size
    A, n = 64 * world, 256
    inp = synth.make_inputs(spec, A, seed=5)
    bp = torch.from_numpy(inp["base_pos"]).to(dev); cd = torch.from_numpy(inp["cond"]).to(dev); g = torch.from_numpy(synth.make_grid(n)).to(dev)
    mm = fdist.MulticastMaps(A, spec.pred_len, n * n, dev)
    for name, fn in (("nccl all-gather", lambda: fdist.sharded_grid_log_prob(f, bp, cd, g, seed=7)), ("multicast stores", lambda: mm.fill(f, bp, cd, g, seed=7))):
        for _ in range(3): fn()
        dist.barrier(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 5], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0: print(f"{name}: {t.item():.2f} ms per pass ({A * n * n * spec.pred_len / t.item() / 1e3:.0f} M evals/s, {world} GPUs)", flush=True)
    if rank == 0: print("CHECK", "OK" if ok else "FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
