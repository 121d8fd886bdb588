"""GPU (one device): the properties the multi-GPU sharding relies on.

  * CFG-branch sharding (parallel.Shard.my_branch / gather_branches): a branch computed alone
    (batch 1) must equal the same row of the joint CFG batch — batch rows never interact inside
    MMDiT.forward (reference mmdit.py:1414-1429);
  * VAE tile sharding: decoding tile-by-tile into caller-owned buffers followed by the blend must
    reproduce the fused decode bit for bit (reference vae.py:994-1011).
The 2-rank exchange steps themselves are covered on CPU over gloo (tests/test_parallel_gloo.py)."""
import pytest
import torch

from oracle import weights
from tests.golden import cases

pytestmark = pytest.mark.gpu


def test_cfg_branch_alone_equals_row_of_joint_batch():
    from deepv_b200.mmdit import B200MMDiT
    cfg, W = weights.mmdit_weights(dict(num_layers=2), seed=1)
    model = B200MMDiT(W, cfg, out_dtype=torch.float32)
    inp = cases.mmdit_inputs(cases.MMDIT_CASES["two_block_b2"])
    dev = "cuda"

    def run(rows):
        return model(sample=[[c[rows].to(dev) for c in inp["clips"]]], timestep_ratio=inp["t"][rows].to(dev),
                     encoder_hidden_states=inp["enc"][rows].to(dev), encoder_attention_mask=inp["mask"][rows].to(dev),
                     pooled_projections=inp["pooled"][rows].to(dev))[0]

    joint = run(slice(0, 2))
    for b in range(2):
        alone = run(slice(b, b + 1))
        torch.cuda.synchronize()
        # split-K factors depend on the tile count, so sums may be re-associated: fp32-level agreement
        err = ((alone - joint[b:b + 1]).abs().max() / joint.abs().max()).item()
        assert err <= 5e-3, (b, err)


def test_tile_granular_decode_equals_fused_decode():
    from deepv_b200.parallel import Shard
    from deepv_b200.vae import B200VAE
    over = dict(decoder_block_out_channels=(128, 128, 128, 128), encoder_block_out_channels=(128, 128, 128, 128),
                decoder_layers_per_block=(1, 1, 1, 1))
    cfg, W = weights.vae_weights(over, seed=7)
    v = B200VAE(W, cfg, dtype=torch.float32)
    v.enable_tiling()
    g = torch.Generator().manual_seed(3)
    zs = [torch.randn(1, 16, 2, 40, 64, generator=g).cuda() for _ in range(2)]
    fused = [v.decode(z, temporal_chunk=True, window_size=1, tile_sample_min_size=256).sample.clone() for z in zs]
    sharded = v.decode_many(zs, Shard(0, 1), tile_sample_min_size=256)   # one rank owns every (modality, tile)
    torch.cuda.synchronize()
    for a, b in zip(fused, sharded):
        assert torch.equal(a, b)


@pytest.mark.parametrize("mode", ["staged", "peer_host_barrier", "peer_device_barrier"])
def test_ulysses_two_virtual_ranks_match_single_rank(mode):
    """Ulysses sequence parallelism (csrc/mmdit.cu + the staging kernels) with two virtual ranks on
    ONE device: two models run the same forward in two threads, each on its own stream, and the
    all-to-all callback swaps the staged blocks between them.  Both must reproduce the unsharded
    forward (same kernels on a row window; split-K factors may re-associate sums).
      staged               pack -> exchange callback -> unpack around every attention
      peer_host_barrier    epilogues store into the owner's buffers, the callback is only a barrier
      peer_device_barrier  ... and the barrier / final all-gather are kernels too (flag words + peer stores): the
                           callback is never called and the forward is replayed as a CUDA graph"""
    peer_memory = mode != "staged"
    import ctypes as C
    import threading

    from deepv_b200.mmdit import B200MMDiT
    cfg, W = weights.mmdit_weights(dict(num_layers=2), seed=1)
    inp = cases.mmdit_inputs(cases.MMDIT_CASES["two_block_b2"])
    dev = "cuda"

    def call(model):
        return model(sample=[[c.to(dev) for c in inp["clips"]]], timestep_ratio=inp["t"].to(dev),
                     encoder_hidden_states=inp["enc"].to(dev), encoder_attention_mask=inp["mask"].to(dev),
                     pooled_projections=inp["pooled"].to(dev))[0]

    ref = call(B200MMDiT(W, cfg, out_dtype=torch.float32))
    torch.cuda.synchronize()

    P = 2
    models = [B200MMDiT(W, cfg, out_dtype=torch.float32) for _ in range(P)]
    barrier = threading.Barrier(P)
    posted = [None] * P

    shared = {}
    payload_calls = [0] * P

    class PeerExchange:
        """peer_memory=True: both 'ranks' live in this process, so their buffers are plain pointers;
        the callback is then only ever called as a barrier (nbytes == 0)."""

        def __init__(self, r, fn):
            self.r, self.fn = r, fn

        def __call__(self, send, recv, nbytes, stream):
            return self.fn(send, recv, nbytes, stream)

        device_barrier = mode == "peer_device_barrier"

        def peer_pointers(self, key, ptrs):
            shared[(key, self.r)] = list(ptrs)
            barrier.wait()
            got = [shared[(key, i)] for i in range(P)]
            barrier.wait()
            return [[g[k] for g in got] for k in range(len(ptrs))]

    def make_exchange(r):
        def exchange(send, recv, nbytes, stream):
            torch.cuda.current_stream().synchronize()          # my staged blocks / peer stores are complete
            if nbytes == 0:                                    # barrier form (peer-memory variant)
                assert peer_memory
                barrier.wait()
                return 0
            payload_calls[r] += 1                              # staged path, or the final fp32 all-gather
            posted[r] = (send, recv, nbytes)
            barrier.wait()
            for i in range(P):                                 # block r of rank i's send -> block i of my recv
                src = posted[i][0] + r * nbytes
                rc = cudart.cudaMemcpy(C.c_void_p(recv + i * nbytes), C.c_void_p(src), C.c_size_t(nbytes), 3)
                assert rc == 0, rc
            barrier.wait()
            return 0
        return exchange

    cudart = C.CDLL("libcudart.so.12")
    outs, errs = [None] * P, []

    def worker(r):
        try:
            with torch.cuda.stream(torch.cuda.Stream()):
                ex = make_exchange(r)
                models[r].set_sequence_parallel(r, P, PeerExchange(r, ex) if peer_memory else ex)
                outs[r] = call(models[r])
                torch.cuda.current_stream().synchronize()
        except Exception as e:  # noqa: BLE001
            errs.append(e)
            barrier.abort()

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(P)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not errs, errs
    for r in range(P):
        err = ((outs[r] - ref).abs().max() / ref.abs().max()).item()
        assert err <= 5e-3, (r, err)
    # peer memory: q|k|v and attention rows travel as stores; with the host barrier only the all-gather of the
    # stream is staged, with the device barrier nothing is
    want = {"staged": 2 * cfg["num_layers"] + 1, "peer_host_barrier": 1, "peer_device_barrier": 0}[mode]
    assert payload_calls == [want] * P
    if mode == "peer_device_barrier":
        # a second forward replays the captured graph (epochs keep counting on the device)
        def again(r):
            try:
                with torch.cuda.stream(torch.cuda.Stream()):
                    outs[r] = call(models[r])
                    torch.cuda.current_stream().synchronize()
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        threads = [threading.Thread(target=again, args=(r,)) for r in range(P)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=120)
        assert not errs, errs
        for r in range(P):
            assert ((outs[r] - ref).abs().max() / ref.abs().max()).item() <= 5e-3
