"""One rollout sharded over the GPUs of one NVSwitch box (one process per GPU, torch.distributed).

What shards on this path and the exchange step each sharding needs:
  * classifier-free-guidance branches — batch rows never interact inside MMDiT.forward
    (sample-id mask, mmdit.py:1414-1429): rank r runs branch r % n_branch with batch 1 and the
    [1,38,1,h,w] predictions are all-gathered (<= 233 KB) before the fused CFG + Euler step,
    which every rank then runs redundantly so the latents stay replicated;
  * VAE decode — the 6 spatial tiles x {rgb, disparity} of one iteration are independent
    (vae.py:994-1000): work items are dealt round-robin, every decoded tile is broadcast from its
    owner, and each rank blends (the blend needs all tiles, vae.py:1002-1011).
Ulysses sequence parallelism around the attention is the next sharding (latency-bound at
L <= 2237) and is not part of round 1.

The functions that decide who does what are pure (tested on CPU); the exchange helpers work on
any backend (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist


def branch_of_rank(rank: int, n_branch: int) -> int:
    return rank % n_branch


def branch_sources(world: int, n_branch: int) -> List[int]:
    """Rank that supplies each branch's prediction (the first rank running that branch)."""
    if world < n_branch:
        raise ValueError(f"{n_branch} CFG branches need at least {n_branch} ranks, got {world}")
    return list(range(n_branch))


def decode_items(n_modalities: int, n_tiles: int) -> List[Tuple[int, int]]:
    """(modality, tile) work items, tile-major so that neighbouring ranks hold different tiles of
    the same modality when the world is small."""
    return [(m, t) for t in range(n_tiles) for m in range(n_modalities)]


def owner_of_item(index: int, world: int) -> int:
    return index % world


def items_of_rank(n_items: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_items, world))


@dataclass
class Shard:
    rank: int
    world: int
    group: object = None

    @staticmethod
    def current(group=None) -> "Shard":
        if dist.is_available() and dist.is_initialized():
            return Shard(dist.get_rank(group), dist.get_world_size(group), group)
        return Shard(0, 1, None)

    @staticmethod
    def grouped(group_size: int) -> "Shard":
        """Split the world into consecutive groups of `group_size` ranks (one rollout per group) and
        return this rank's shard of its group.  Every rank must call this (new_group is collective)."""
        if not (dist.is_available() and dist.is_initialized()):
            return Shard(0, 1, None)
        world, rank = dist.get_world_size(), dist.get_rank()
        if group_size <= 1:
            return Shard(0, 1, None)
        if world % group_size != 0:
            raise ValueError(f"world size {world} is not a multiple of the rollout group size {group_size}")
        mine = None
        for g in range(world // group_size):
            ranks = list(range(g * group_size, (g + 1) * group_size))
            grp = dist.new_group(ranks)
            if rank in ranks:
                mine = grp
        return Shard(dist.get_rank(mine), group_size, mine)

    def _global(self, group_rank: int) -> int:
        return dist.get_global_rank(self.group, group_rank) if self.group is not None else group_rank

    @property
    def active(self) -> bool:
        return self.world > 1

    # ---- CFG branches -----------------------------------------------------------------------
    def my_branch(self, n_branch: int) -> int:
        return branch_of_rank(self.rank, n_branch)

    def gather_branches(self, pred_local: torch.Tensor, n_branch: int) -> torch.Tensor:
        """pred_local: this rank's branch prediction [1, ...] -> [n_branch, ...] ordered
        (uncond, text[, text+history]) on every rank."""
        if not self.active:
            return pred_local
        bufs = [torch.empty_like(pred_local) for _ in range(self.world)]
        dist.all_gather(bufs, pred_local.contiguous(), group=self.group)
        return torch.cat([bufs[r] for r in branch_sources(self.world, n_branch)], dim=0)

    # ---- VAE tiles -------------------------------------------------------------------------------
    def my_items(self, n_items: int) -> List[int]:
        return items_of_rank(n_items, self.rank, self.world)

    def exchange_tiles(self, tile_buffers: Sequence[torch.Tensor]) -> None:
        """tile_buffers[i] lives on every rank and was filled by owner_of_item(i); broadcast each
        from its owner so that every rank can blend."""
        if not self.active:
            return
        works = [dist.broadcast(buf, src=self._global(owner_of_item(i, self.world)), group=self.group, async_op=True)
                 for i, buf in enumerate(tile_buffers)]
        for w in works:
            w.wait()
