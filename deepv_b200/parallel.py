"""One rollout sharded over the GPUs of one NVSwitch box (one process per GPU, torch.distributed).

What shards on this path and the exchange step each sharding needs:
  * classifier-free-guidance branches — batch rows never interact inside MMDiT.forward
    (sample-id mask, mmdit.py:1414-1429): rank r runs branch r % n_branch with batch 1 and the
    [1,38,1,h,w] predictions are all-gathered (<= 233 KB) before the fused CFG + Euler step,
    which every rank then runs redundantly so the latents stay replicated;
  * VAE decode — the 6 spatial tiles x {rgb, disparity} of one iteration are independent
    (vae.py:994-1000): work items are dealt round-robin, every decoded tile is broadcast from its
    owner, and each rank blends (the blend needs all tiles, vae.py:1002-1011).
  * Ulysses sequence parallelism inside a branch: a branch group of `sp` ranks
    shards the video tokens; around every attention heads <-> tokens are exchanged, either as
    epilogue stores into the owning rank's buffers over NVLink (CUDA IPC peer memory, the default)
    or as a staged NCCL all-to-all (`DV_SP_STAGED=1`); see csrc/mmdit.cu / csrc/comm.cu.
PRECONDITION of all three: the latents (and therefore every random draw) are REPLICATED over the
ranks of a rollout group.  `Shard.shared_seed()` broadcasts one seed so that each rank's device
generator produces the same stream; `B200Rollout` uses it whenever it has to make its own noise
source under a shard, and refuses an unseeded one.

The functions that decide who does what are pure (tested on CPU); the exchange helpers work on
any backend (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist


class NcclExchange:
    """The equal-block all-to-all of Ulysses sequence parallelism, run by NCCL INSIDE
    libdeepv_b200.so (csrc/comm.cu): `fn_ptr` / `user_ptr` are handed to dv_mmdit_plan_set_sp, so a
    sequence-parallel forward stays one C call.  The communicator is created from a unique id that
    rank 0 of `group` makes and torch.distributed broadcasts."""

    def __init__(self, group=None, device=None, peer_memory=True):
        import ctypes as C
        self.group, self.device = group, device
        if not peer_memory:
            self.peer_pointers = None   # (instance attribute shadows the method: staged all-to-all)

        from . import _lib
        lib = _lib.load()
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        src = dist.get_global_rank(group, 0) if group is not None else 0
        path = self._nccl_path()
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_char * 128)()
            _lib.check(lib.dv_comm_unique_id(path, C.cast(buf, C.c_void_p)), "dv_comm_unique_id")
            ident = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        backend = dist.get_backend(group)
        carrier = ident.to(device) if backend == "nccl" else ident
        dist.broadcast(carrier, src=src, group=group)
        raw = bytes(carrier.cpu().numpy().tobytes())
        handle = C.c_void_p()
        _lib.check(lib.dv_comm_create(path, raw, rank, world, C.byref(handle)), "dv_comm_create")
        self._lib, self._handle = lib, handle
        self.rank, self.world = rank, world
        self.fn_ptr = C.cast(lib.dv_comm_exchange, C.c_void_p)
        self.user_ptr = handle
        self._opened = []     # peers' buffers mapped through CUDA IPC (closed in close())
        self._peer_cache = {}
        # DV_SP_NCCL_BARRIER=1: keep NCCL's one-word all-reduce as the barrier of the peer-memory variant
        self.device_barrier = os.environ.get("DV_SP_NCCL_BARRIER", "") == ""

    def peer_pointers(self, _key, ptrs):
        """Map the listed device buffers of every rank into this process (CUDA IPC): `ptrs` = this rank's
        [qkv, attention, fp32 stream, barrier flags]; returns one list per buffer with every rank's pointer
        (own entry = own pointer).  The forward then stores q|k|v heads and attention rows straight into
        their owner's buffer over NVLink and synchronises through the flag words.  Collective over the group."""
        import ctypes as C

        from . import _lib
        lib = self._lib
        n = len(ptrs)
        ck = (_key, tuple(ptrs))
        if ck in self._peer_cache:      # the same plan re-sharded the same way: its peers are already mapped
            return self._peer_cache[ck]  # (every rank of the group takes this branch together)
        mine = (C.c_char * (64 * n))()
        for i, ptr in enumerate(ptrs):
            _lib.check(lib.dv_ipc_get_handle(ptr, C.cast(C.byref(mine, 64 * i), C.c_void_p)), "dv_ipc_get_handle")
        t = torch.frombuffer(bytearray(mine.raw), dtype=torch.uint8).clone()
        on_dev = dist.get_backend(self.group) == "nccl"
        t = t.to(self.device) if on_dev else t
        bufs = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(bufs, t, group=self.group)
        out = [[] for _ in range(n)]
        for r in range(self.world):
            if r == self.rank:
                for i in range(n):
                    out[i].append(ptrs[i])
                continue
            raw = bytes(bufs[r].cpu().numpy().tobytes())
            for i in range(n):
                ptr = C.c_void_p()
                _lib.check(lib.dv_ipc_open_handle(raw[64 * i:64 * (i + 1)], C.byref(ptr)), "dv_ipc_open_handle")
                self._opened.append(ptr.value)
                out[i].append(ptr.value)
        self._peer_cache[ck] = out
        return out

    @staticmethod
    def _nccl_path():
        import glob
        import os
        try:
            import nvidia  # the wheel torch depends on
            for base in nvidia.__path__:
                hits = glob.glob(os.path.join(base, "nccl", "lib", "libnccl.so.2"))
                if hits:
                    return hits[0].encode()
        except Exception:
            pass
        return b""

    def close(self):
        """Unmap the peers' buffers and destroy the communicator.  Collective: every rank of the group must
        have stopped issuing forwards first (the owners free the buffers when their plans are destroyed)."""
        if self._opened:
            if dist.is_initialized():
                torch.cuda.synchronize()
                dist.barrier(group=self.group)
            for ptr in self._opened:
                self._lib.dv_ipc_close_handle(ptr)
            self._opened = []
            self._peer_cache = {}
        if self._handle:
            self._lib.dv_comm_destroy(self._handle)
            self._handle = None


def sp_ranks(world: int, n_branch: int):
    """How a rollout group of `world` ranks splits into CFG branches x sequence-parallel ranks:
    returns (n_branch_groups, sp_world) with n_branch_groups * sp_world == world, or (1, world)
    when the branches do not divide the group."""
    if n_branch > 1 and world % n_branch == 0:
        return n_branch, world // n_branch
    return 1, world


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
    sp_layouts: object = None   # n_branch -> (branch groups, sp_world, NcclExchange | None), see setup_sp

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

    # ---- replicated randomness ---------------------------------------------------------------------
    def shared_seed(self, device=None) -> int:
        """One 62-bit seed drawn by rank 0 of the group and broadcast: seeding a device generator with it on
        every rank replicates all later draws (same device type, same draw order).  Collective over the group."""
        t = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64)
        if not self.active:
            return int(t.item())
        if dist.get_backend(self.group) == "nccl":
            t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
        dist.broadcast(t, src=self._global(0), group=self.group)
        return int(t.item())

    def assert_replicated(self, x: torch.Tensor, what: str = "tensor") -> None:
        """Debug check (DV_CHECK_REPLICAS=1 in B200Rollout): `x` is bit-identical on every rank of the group."""
        if not self.active:
            return
        ref = x.detach().clone().contiguous()
        dist.broadcast(ref, src=self._global(0), group=self.group)
        if not torch.equal(ref, x):
            raise RuntimeError(f"{what} differs between the ranks of a rollout group (rank {self.rank}): "
                               "sharded rollouts need replicated latents / noise")

    # ---- CFG branches x Ulysses sequence parallelism --------------------------------------------
    def setup_sp(self, device, branch_counts=(1, 2, 3), peer_memory=None) -> None:
        """Create the sequence-parallel sub-groups (and their in-library NCCL communicators) for
        every CFG batch size the rollout uses.  A rollout group of G ranks runs `nb` branch groups
        of `sp = G / nb` ranks each (sp_ranks()); ranks [b*sp, (b+1)*sp) of the group hold branch b.
        Collective over the WHOLE world (torch.distributed.new_group)."""
        self.sp_layouts = {}
        if not self.active:
            return
        if peer_memory is None:   # DV_SP_STAGED=1: pack -> NCCL all-to-all -> unpack instead of peer stores
            import os
            peer_memory = os.environ.get("DV_SP_STAGED", "") == ""
        me, world_all = dist.get_rank(), dist.get_world_size()
        n_rollouts = world_all // self.world
        made = {}
        for nb in branch_counts:
            g_nb, sp = sp_ranks(self.world, nb)
            if sp == 1:
                self.sp_layouts[nb] = (g_nb, 1, None)
                continue
            if (g_nb, sp) not in made:
                mine = None
                for ro in range(n_rollouts):
                    for b in range(g_nb):
                        ranks = [ro * self.world + b * sp + i for i in range(sp)]
                        grp = dist.new_group(ranks)
                        if me in ranks:
                            mine = grp
                made[(g_nb, sp)] = NcclExchange(mine, device, peer_memory=peer_memory)
            self.sp_layouts[nb] = (g_nb, sp, made[(g_nb, sp)])

    def layout(self, n_branch: int):
        """(branch groups, sp_world, exchange) for a CFG batch of n_branch.  Without setup_sp():
        branch r % n_branch on rank r, no sequence parallelism (extra ranks duplicate a branch)."""
        if self.sp_layouts and n_branch in self.sp_layouts:
            return self.sp_layouts[n_branch]
        if self.active and n_branch > 1 and self.world >= n_branch:
            return (n_branch, 1, None)
        return (1, 1, None)

    def my_branch(self, n_branch: int) -> int:
        g_nb, sp, _ = self.layout(n_branch)
        if self.sp_layouts and n_branch in self.sp_layouts:
            return (self.rank // sp) if g_nb > 1 else 0
        return branch_of_rank(self.rank, n_branch)

    def gather_branches(self, pred_local: torch.Tensor, n_branch: int) -> torch.Tensor:
        """pred_local: this rank's branch prediction [1, ...] -> [n_branch, ...] ordered
        (uncond, text[, text+history]) on every rank."""
        if not self.active:
            return pred_local
        _, sp, _ = self.layout(n_branch)
        bufs = [torch.empty_like(pred_local) for _ in range(self.world)]
        dist.all_gather(bufs, pred_local.contiguous(), group=self.group)
        return torch.cat([bufs[b * sp] for b in range(n_branch)], dim=0)

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
