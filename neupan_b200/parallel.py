"""Multi-GPU plumbing of the hot path: environments are independent, so the batch is
split into contiguous blocks, one process per GPU; the single exchange of a control step is the
gather of the packed per-environment result [S 3(T+1) | U 2T | D T | min_distance 1].

Works with any torch.distributed backend (NCCL over NVLink on the B200 box, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total_envs: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of environments owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(total_envs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_results(S: torch.Tensor, U: torch.Tensor, D: torch.Tensor, min_distance: torch.Tensor) -> torch.Tensor:
    B = S.shape[0]
    return torch.cat([S.reshape(B, -1), U.reshape(B, -1), D.reshape(B, -1), min_distance.reshape(B, 1)], dim=1).contiguous()


def unpack_results(packed: torch.Tensor, T: int):
    B = packed.shape[0]
    a, b, c = 3 * (T + 1), 3 * (T + 1) + 2 * T, 3 * (T + 1) + 3 * T
    return packed[:, :a].reshape(B, 3, T + 1), packed[:, a:b].reshape(B, 2, T), packed[:, b:c].reshape(B, 1, T), packed[:, c]


def gather_results(packed: torch.Tensor, total_envs: int, group=None) -> torch.Tensor:
    """all_gather of the per-rank packed results into (total_envs, width), rank-major = env order.
    Ranks may own different numbers of envs (shard_range); shorter shards are padded for the collective."""
    world = dist.get_world_size(group)
    if world == 1:
        return packed
    sizes = [shard_range(total_envs, r, world) for r in range(world)]
    longest = max(hi - lo for lo, hi in sizes)
    buf = packed
    if packed.shape[0] < longest:
        buf = torch.zeros((longest, packed.shape[1]), dtype=packed.dtype, device=packed.device)
        buf[: packed.shape[0]] = packed
    out = torch.empty((world * longest, packed.shape[1]), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    if all(hi - lo == longest for lo, hi in sizes):
        return out
    return torch.cat([out[r * longest: r * longest + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)
