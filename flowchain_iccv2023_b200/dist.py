"""Multi-GPU sharding of the agent x grid batch.

Every (agent, grid point) row is independent (eval-mode BatchNorm has no cross-row statistics), so the
path shards with NO data-path collective: packed weights (4.4 MB) are replicated, agents are split in
contiguous blocks over ranks (grid tiles when there are fewer agents than ranks), and ONE all-gather
assembles the (A, T, G) density maps.  One process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in
the CPU tests) is plumbing only.

On an NVSwitch box the all-gather is fused into the density kernel (`MulticastMaps`): the maps live in a
symmetric buffer with a multicast mapping and the kernel's result stores go to the multicast address, so the
switch replicates every store to all peers while the kernel is still computing; what remains after the kernel
is one cross-rank barrier.  The NCCL all-gather is the fallback (no multicast support, gloo tests).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `n` units owned by `rank`; blocks differ by at most one unit."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def plan_shards(agents: int, points: int, world: int):
    """Returns ('agents', ranges) when every rank gets >= 1 agent, else ('points', ranges):
    the grid is split instead (config 2 at 8 GPUs can go either way)."""
    if agents >= world:
        return "agents", [shard_range(agents, r, world) for r in range(world)]
    return "points", [shard_range(points, r, world) for r in range(world)]


def _gather(local: torch.Tensor, sizes, axis: int, group=None) -> torch.Tensor:
    world = dist.get_world_size(group)
    m = max(sizes)
    if axis != 0:
        local = local.movedim(axis, 0).contiguous()
    if local.shape[0] != m:
        pad = torch.zeros((m - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    local = local.contiguous()
    out = torch.empty((world * m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    if len(set(sizes)) != 1:
        out = torch.cat([out[r * m: r * m + sizes[r]] for r in range(world)], dim=0)
    if axis != 0:
        out = out.movedim(0, axis).contiguous()
    return out


def gather_agent_maps(local: torch.Tensor, sizes, group=None) -> torch.Tensor:
    """(a_r, T, G) shards -> (A, T, G)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    return _gather(local, list(sizes), 0, group)


def gather_point_maps(local: torch.Tensor, sizes, group=None) -> torch.Tensor:
    """(A, T, g_r) shards -> (A, T, G)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    return _gather(local, list(sizes), 2, group)


class MulticastMaps:
    """Symmetric (A, T, G) density-map buffer, one replica per rank, with an NVSwitch multicast address.

    `fill(flow, ...)` evaluates this rank's shard with `fc_grid_log_prob_multicast`: the kernel's stores are replicated
    by the switch into every rank's replica, then one symmetric-memory barrier makes them readable everywhere.
    Raises RuntimeError when the platform has no multicast support (callers fall back to `sharded_grid_log_prob`)."""

    def __init__(self, agents: int, steps: int, points: int, device, group=None):
        import torch.distributed._symmetric_memory as symm
        grp = dist.group.WORLD if group is None else group
        self.group = grp
        self.shape = (agents, steps, points)
        self.maps = symm.empty(self.shape, dtype=torch.float32, device=device)
        self.handle = symm.rendezvous(self.maps, grp.group_name)
        self.mc_ptr = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        if self.mc_ptr == 0:
            raise RuntimeError("symmetric memory has no multicast mapping on this platform")

    def barrier(self):
        self.handle.barrier()

    def fill(self, flow, base_pos, cond, grid, cond_traj=None, **kw) -> torch.Tensor:
        """Inputs are the FULL per-agent tensors (replicated, tiny).  Returns this rank's replica of the full maps."""
        A, T, G = self.shape
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        mode, ranges = plan_shards(A, G, world)
        lo, hi = ranges[rank]
        self.barrier()          # nobody is still reading the previous maps
        if mode == "agents":
            flow.grid_log_prob_multicast(base_pos[lo:hi], cond[lo:hi], grid, self.mc_ptr, lo * T * G, (T * G, 1, G),
                                         cond_traj=None if cond_traj is None else cond_traj[lo:hi], **kw)
        else:
            flow.grid_log_prob_multicast(base_pos, cond, grid[lo:hi].contiguous(), self.mc_ptr, lo, (T * G, 1, G),
                                         cond_traj=cond_traj, **kw)
        self.barrier()          # every rank's stores have landed in every replica
        return self.maps


def sharded_grid_log_prob(flow, base_pos, cond, grid, cond_traj=None, group=None, **kw) -> torch.Tensor:
    """Density maps (A, T, G) for the whole batch, computed on this rank's shard and all-gathered.
    Inputs are the FULL per-agent tensors (replicated, tiny); every rank returns the full maps."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return flow.grid_log_prob(base_pos, cond, grid, cond_traj=cond_traj, map_layout=True, **kw)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    A, G = base_pos.shape[0], grid.shape[0]
    mode, ranges = plan_shards(A, G, world)
    lo, hi = ranges[rank]
    sizes = [b - a for a, b in ranges]
    if "eps" in kw and kw["eps"] is not None:
        raise ValueError("noise tapes are per-row; shard them with the rows before calling the per-rank op")
    if mode == "agents":
        local = flow.grid_log_prob(base_pos[lo:hi], cond[lo:hi], grid,
                                   cond_traj=None if cond_traj is None else cond_traj[lo:hi], map_layout=True, **kw)
        return gather_agent_maps(local, sizes, group)
    local = flow.grid_log_prob(base_pos, cond, grid[lo:hi].contiguous(), cond_traj=cond_traj, map_layout=True, **kw)
    return gather_point_maps(local, sizes, group)
