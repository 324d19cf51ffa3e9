"""Frame sharding across ranks (one process per GPU) and the single all-gather of
per-frame result rows.

The hot path shards by frames: the spatial entropy of frame f depends on frame f
only (SA:129-161); transition row r depends on frames r and r+1 (TA:143-172), so a
rank also reads ONE halo frame that it loads itself -- no sample ever crosses NVLink.
Users are never split (the literal transition bookkeeping depends on global user
order, EU:259-294).  The only collective is one all-gather of the packed per-frame
rows [entropy | hist0 | ...]; tile assignments stay sharded on the owning rank.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def frame_range(num_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous frame range [begin, end) of `rank`; the first F % world ranks get one extra frame."""
    base, extra = divmod(num_frames, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def transition_range(num_frames: int, rank: int, world: int) -> Tuple[int, int, int]:
    """Transition rows [r0, r1) owned by `rank` (row r pairs frames r and r+1) and the
    frame range [r0, r1 + 1) it must read: its own rows plus one halo frame."""
    r0, r1 = frame_range(max(num_frames - 1, 0), rank, world)
    return r0, r1, (r1 + 1 if r1 > r0 else r1)


def all_gather_rows(local: torch.Tensor, counts: List[int], group=None) -> torch.Tensor:
    """Concatenates per-rank row blocks [n_rank, D] in rank order with ONE collective.
    Equal shards use all_gather_into_tensor directly; ragged shards are padded to the
    largest block and trimmed afterwards."""
    world = len(counts)
    if world == 1:
        return local
    n_max = max(counts)
    width = local.shape[1:]
    if local.shape[0] != n_max:
        pad = torch.zeros((n_max - local.shape[0],) + tuple(width), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], 0)
    out = torch.empty((world * n_max,) + tuple(width), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    if all(c == n_max for c in counts):
        return out
    return torch.cat([out[r * n_max: r * n_max + c] for r, c in enumerate(counts)], 0)


@dataclass
class ShardedResult:
    entropy: torch.Tensor            # [F] (or [F-1]) full, on every rank
    rows: torch.Tensor               # [F, D] the gathered per-frame payload (hist0 / prev_count0 as float64)
    local_begin: int                 # first row owned by this rank
    local_end: int


def run_sharded(num_frames: int, compute_local: Callable[[int, int], Tuple[torch.Tensor, torch.Tensor]],
                transition: bool = False, group=None, rank: Optional[int] = None,
                world: Optional[int] = None) -> ShardedResult:
    """Runs `compute_local(frame_begin, frame_end)` on this rank's share and gathers.

    compute_local returns (entropy[n_rows], payload[n_rows, D]) for the rows of the given
    frame range: n_rows = frames for the spatial stage, frames - 1 for the transition stage
    (whose frame range already includes the halo frame)."""
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if transition:
        r0, r1, fend = transition_range(num_frames, rank, world)
        counts = [transition_range(num_frames, r, world)[1] - transition_range(num_frames, r, world)[0] for r in range(world)]
        ent, payload = compute_local(r0, fend)
    else:
        r0, r1 = frame_range(num_frames, rank, world)
        counts = [frame_range(num_frames, r, world)[1] - frame_range(num_frames, r, world)[0] for r in range(world)]
        ent, payload = compute_local(r0, r1)
    if payload.dim() == 1:
        payload = payload[:, None]
    packed = torch.cat([ent.to(torch.float64)[:, None], payload.to(torch.float64)], 1)
    full = all_gather_rows(packed, counts, group)
    return ShardedResult(entropy=full[:, 0].contiguous(), rows=full[:, 1:], local_begin=r0, local_end=r1)
