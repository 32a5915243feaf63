"""Probe: torch symmetric memory on this box (2+ ranks): allocate, rendezvous, write into a peer buffer, barrier, verify."""
import os, sys, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm

def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    n = 1 << 20
    t = symm.empty(world * n, dtype=torch.float32, device=torch.device("cuda", lr))
    hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "rendezvous ok; ptrs", [hex(p) for p in hdl.buffer_ptrs][:4], "multicast", getattr(hdl, "multicast_ptr", None), flush=True)
    t.zero_()
    hdl.barrier()
    src = torch.full((n,), float(rank + 1), device="cuda")
    for r in range(world):
        peer = hdl.get_buffer(r, (world * n,), torch.float32)
        peer[rank * n:(rank + 1) * n].copy_(src)
    hdl.barrier()
    torch.cuda.synchronize()
    ok = all(float(t[r * n].item()) == r + 1 and float(t[(r + 1) * n - 1].item()) == r + 1 for r in range(world))
    print(rank, "verify", ok, flush=True)
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
