"""CPU, world_size 2 (gloo): the multi-GPU sharding logic of deepv_b200/parallel.py.

The kernels need a GPU, but who-computes-what and the exchange steps do not: these tests run the
real Shard helpers over gloo with stand-in 'branch predictions' and 'decoded tiles'."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepv_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = parallel.Shard.current()
        assert (sh.rank, sh.world, sh.active) == (rank, world, True)
        # ---- CFG branches: rank r computes branch r % n; gather yields (uncond, text) in order
        n_branch = 2
        mb = sh.my_branch(n_branch)
        pred = torch.full((1, 38, 1, 4, 4), float(10 + mb))
        allp = sh.gather_branches(pred, n_branch)
        assert allp.shape == (2, 38, 1, 4, 4)
        assert torch.equal(allp[0], torch.full((38, 1, 4, 4), 10.0))
        assert torch.equal(allp[1], torch.full((38, 1, 4, 4), 11.0))
        # the combine every rank then runs must give identical latents everywhere
        guided = allp[0] + 3.5 * (allp[1] - allp[0])
        chk = [torch.empty_like(guided) for _ in range(world)]
        dist.all_gather(chk, guided)
        assert all(torch.equal(c, chk[0]) for c in chk)
        # ---- VAE work items: every (modality, tile) decoded exactly once, then visible everywhere
        items = parallel.decode_items(2, 6)
        mine = sh.my_items(len(items))
        bufs = [torch.zeros(3, 5) for _ in items]
        for i in mine:
            bufs[i].fill_(100 * items[i][0] + items[i][1] + 1)
        sh.exchange_tiles(bufs)
        for i, (m, t) in enumerate(items):
            assert torch.equal(bufs[i], torch.full((3, 5), float(100 * m + t + 1))), (rank, i)
        # ---- replicated randomness: one broadcast seed, identical generator streams on every rank
        torch.manual_seed(1000 + rank)                 # ranks start from DIFFERENT global RNG states
        seed = sh.shared_seed()
        draws = torch.randn(5, generator=torch.Generator().manual_seed(seed))
        alld = [torch.empty_like(draws) for _ in range(world)]
        dist.all_gather(alld, draws)
        assert all(torch.equal(d, alld[0]) for d in alld)
        sh.assert_replicated(draws, "draws")
        try:                                          # a tensor that differs per rank is caught on every rank but the source
            sh.assert_replicated(torch.full((3,), float(rank)), "rank id")
            caught = False
        except RuntimeError:
            caught = True
        assert caught == (rank != 0)
        ret[rank] = len(mine)
    finally:
        dist.destroy_process_group()


def test_sharding_over_gloo_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert sorted(ret.values()) == [6, 6]


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8, 12, 16])
def test_item_partition_is_exact(world):
    items = parallel.decode_items(2, 6)
    seen = []
    for r in range(world):
        mine = parallel.items_of_rank(len(items), r, world)
        assert all(parallel.owner_of_item(i, world) == r for i in mine)
        seen += mine
    assert sorted(seen) == list(range(len(items)))
    # balance: no rank holds more than ceil(n / world)
    assert max(len(parallel.items_of_rank(len(items), r, world)) for r in range(world)) == -(-len(items) // world)


def test_branch_assignment():
    assert [parallel.branch_of_rank(r, 3) for r in range(8)] == [0, 1, 2, 0, 1, 2, 0, 1]
    assert parallel.branch_sources(8, 3) == [0, 1, 2]
    with pytest.raises(ValueError):
        parallel.branch_sources(2, 3)
    sh = parallel.Shard(0, 1)
    x = torch.randn(2, 3)
    assert sh.gather_branches(x, 2) is x and not sh.active


def _worker_grouped(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = parallel.Shard.grouped(2)   # world 4 -> two independent rollouts on rank pairs
        assert (sh.rank, sh.world) == (rank % 2, 2)
        rollout = rank // 2
        pred = torch.full((1, 2, 2), float(100 * rollout + sh.my_branch(2)))
        allp = sh.gather_branches(pred, 2)
        assert allp[:, 0, 0].tolist() == [100.0 * rollout, 100.0 * rollout + 1]
        items = parallel.decode_items(2, 3)
        bufs = [torch.zeros(4) for _ in items]
        for i in sh.my_items(len(items)):
            bufs[i].fill_(1000 * rollout + i + 1)
        sh.exchange_tiles(bufs)   # owners are GROUP ranks; the broadcast must map them to global ranks
        assert [b[0].item() for b in bufs] == [1000.0 * rollout + i + 1 for i in range(len(items))]
        ret[rank] = rollout
    finally:
        dist.destroy_process_group()


def test_rollout_groups_over_gloo_world4():
    world = 4
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_grouped, args=(world, port, ret), nprocs=world, join=True)
    assert [ret[r] for r in range(4)] == [0, 0, 1, 1]


def test_rollout_group_layouts():
    """A rollout group splits into CFG branch groups x Ulysses ranks (parallel.sp_ranks)."""
    assert parallel.sp_ranks(2, 2) == (2, 1)     # a GPU pair: one branch each
    assert parallel.sp_ranks(4, 2) == (2, 2)
    assert parallel.sp_ranks(8, 2) == (2, 4)
    assert parallel.sp_ranks(8, 3) == (1, 8)     # 3 branches do not divide 8: whole batch, 8-way Ulysses
    assert parallel.sp_ranks(6, 3) == (3, 2)
    assert parallel.sp_ranks(4, 1) == (1, 4)
    # without setup_sp(): branch r % n on rank r, no sequence parallelism, extra ranks duplicate
    sh = parallel.Shard(3, 4)
    assert sh.layout(2) == (2, 1, None) and sh.my_branch(2) == 1
    assert parallel.Shard(0, 1).layout(3) == (1, 1, None)
    # with layouts: branch = rank // sp_world, Ulysses rank = rank % sp_world
    sh = parallel.Shard(5, 8, sp_layouts={2: (2, 4, None), 3: (1, 8, None)})
    assert sh.my_branch(2) == 1 and sh.layout(2)[1] == 4
    assert sh.my_branch(3) == 0 and sh.layout(3)[1] == 8


def test_sharded_rollout_refuses_unseeded_noise():
    """a rollout group must draw identical noise on every rank (rollout.noise_for)."""
    from deepv_b200 import _lib
    from deepv_b200.rollout import DeviceNoise, noise_for

    class FakePipe:
        device = torch.device("cpu")

    class FakeShard:
        active = True

        def shared_seed(self, device=None):
            return 1234

    pipe = FakePipe()
    with pytest.raises(_lib.DeepVError):
        noise_for(pipe, DeviceNoise(pipe, None), FakeShard())
    seeded = noise_for(pipe, None, FakeShard())
    assert isinstance(seeded, DeviceNoise) and seeded.generator.initial_seed() == 1234
    assert isinstance(noise_for(pipe, None, None), DeviceNoise)          # un-sharded: the global generator is fine
    mine = DeviceNoise(pipe, torch.Generator().manual_seed(7))
    assert noise_for(pipe, mine, FakeShard()) is mine
