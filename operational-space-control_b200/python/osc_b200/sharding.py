"""Multi-GPU plumbing: environments are independent, so the hot path has NO collective.
Each rank owns a contiguous range of environments (one process and one osc_handle per
GPU).  After a step the torques and per-rank statistics are gathered by PEER STORES behind the
C-ABI (osc_gather_*: every rank writes its slice into the slabs its peers exported through
CUDA IPC); this module only moves the 64-byte handles between the ranks (`exchange_bytes`,
any torch.distributed backend) and keeps the torch.distributed all-gather / all-reduce used
by the CPU tests and as the cross-check of the peer-store path.  SURVEY.md 8(e)."""
from __future__ import annotations

import torch
import torch.distributed as dist

BLOCK = 256  # synthetic inputs are generated in 256-env blocks (synth.make_inputs)


def shard_range(total_envs: int, rank: int, world: int):
    """(first_env, count) of rank's contiguous shard; shards start on BLOCK multiples."""
    if world == 1:
        return 0, total_envs
    if total_envs % BLOCK:
        raise ValueError(f"total_envs must be a multiple of {BLOCK} to shard")
    blocks = total_envs // BLOCK
    lo = (blocks * rank) // world
    hi = (blocks * (rank + 1)) // world
    return lo * BLOCK, (hi - lo) * BLOCK


def all_gather_rows(local: torch.Tensor, total_rows: int, world: int) -> torch.Tensor:
    """Gather per-rank [rows_r, k] tensors (possibly ragged) into [total_rows, k]."""
    if world == 1:
        return local
    counts = [shard_range(total_rows, r, world)[1] for r in range(world)]
    pad = max(counts)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat([o[:c] for o, c in zip(out, counts)], 0)


def exchange_bytes(local: bytes, world: int, device=None) -> bytes:
    """All ranks' fixed-size byte strings concatenated in rank order (the CUDA IPC handles of
    osc_gather_create).  Works on gloo (CPU tensors) and nccl (pass the CUDA device)."""
    if world == 1:
        return bytes(local)
    t = torch.tensor(list(local), dtype=torch.uint8, device=device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return b"".join(bytes(o.cpu().tolist()) for o in out)


def setup_peer_gather(osc, rank: int, world: int, device=None):
    """osc_gather_create + handle exchange + osc_gather_attach for one-process-per-GPU jobs."""
    mine = osc.gather_create(rank, world)
    if world > 1:
        osc.gather_attach(ipc_handles=exchange_bytes(mine, world, device))
        dist.barrier()  # nobody pushes before every slab is mapped everywhere


def reduce_stats(stats: dict, world: int, device=None) -> dict:
    if world == 1:
        return dict(stats)
    keys = sorted(stats)
    t = torch.tensor([float(stats[k]) for k in keys], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: type(stats[k])(v) for k, v in zip(keys, t.tolist())}


def max_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
