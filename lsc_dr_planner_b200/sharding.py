"""Agent sharding across ranks: contiguous index blocks, one process per GPU.

Agents are independent inside one replan (the reference's loop is Jacobi-style,
src/multi_sync_simulator.cpp:305-362), so the data path needs no collective; the only exchange
is one all-gather of the solved trajectories per closed-loop step, which is what
MultiSyncSimulator::broadcastMsgs does in-process for the reference.
Works with any torch.distributed backend (nccl on GPUs, gloo in the CPU tests)."""
from __future__ import annotations


def shard_range(n_agents: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of the agents owned by `rank`: blocks of ceil(n/world), the last ones possibly short or empty."""
    per = (n_agents + world - 1) // world
    lo = min(n_agents, rank * per)
    return lo, min(n_agents, lo + per)


def shard_sizes(n_agents: int, world: int) -> list[int]:
    return [shard_range(n_agents, r, world)[1] - shard_range(n_agents, r, world)[0] for r in range(world)]


def allgather_rows(local, n_total: int, group=None):
    """Concatenate per-rank row blocks (shard_range layout) into the full [n_total, ...] tensor on every rank.
    `local` may be shorter than the block size on the last ranks; it is padded for the collective."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    per = (n_total + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    return out[:n_total]


def allgather_handles(payload: bytes, ok: bool, device=None, group=None) -> tuple[bytes, bool]:
    """All-gather one fixed-size opaque handle per rank (the 64-byte CUDA IPC handles of the peer exchange,
    lscqp_exchange_create) together with a per-rank success flag.  Returns (world x len(payload) bytes in rank order,
    every rank succeeded).  Plumbing only: works on any backend (tensors live on `device`; cpu for gloo)."""
    import torch
    import torch.distributed as dist
    n = len(payload)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return payload, bool(ok)
    world = dist.get_world_size(group)
    mine = torch.tensor(list(payload) + [1 if ok else 0], dtype=torch.uint8, device=device)
    allh = torch.empty((world * (n + 1),), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(allh, mine, group=group)
    allh = allh.cpu().numpy().reshape(world, n + 1)
    return allh[:, :n].tobytes(), bool(allh[:, n].min() == 1)


def all_agree(ok: bool, device=None, group=None) -> bool:
    """True only if `ok` holds on every rank (MIN all-reduce): all ranks take the same exchange path"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return bool(ok)
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(int(flag.item()) == 1)
