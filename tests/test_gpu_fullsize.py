"""GPU parity at the PRODUCTION sizes the bench times.

The CPU oracle needs minutes for these shapes, so the same oracle code (oracle/*.py, plain torch, pinned
bit-for-bit to the real reference by the CPU tests) runs here in **fp32 on the CUDA device with TF32
off** as the checker; the product path under test is the sm_100a library as everywhere else.

  * VAE decoder at the App. A config (128,256,512,512) x 3 layers/block: the benched tiled decode
    [1,16,2,48,64] (6 tiles, blends), a single 32x32 tile over 8 latent frames (all 7 temporal windows
    of vae.py:903-920), both for fp32 and bf16 I/O; PSNR floor 40 dB incl. the seam bands;
  * the conv shapes only the production decoder has (512->2048 pixel-shuffle, 512->1024 frame-interleave,
    512->512 @ T3 48x64, 512->256 and 256->128 1x1x1 shortcuts) against F.conv3d;
  * VAE encode at (128,256,512,512) x 2 layers/block;
  * the 24-block denoiser on the largest layouts of a rollout: first-iteration stage 2 (B=2, L=1536+77)
    and steady-state unit 8 stage 2 (B=3, history, mixed-resolution clips, L=1968+269): <= 2e-2.
"""
import contextlib
import ctypes as C
import math

import pytest
import torch

from oracle import mmdit_ref, vae_ref, weights
from tests.golden import cases

pytestmark = pytest.mark.gpu
PSNR_FLOOR_DB = 40.0
DENOISER_TOL = 2e-2


@contextlib.contextmanager
def exact_fp32():
    """The checker's arithmetic: fp32 with TF32 off (cuDNN convs default to TF32 otherwise)."""
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            yield
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def psnr(a, b, peak=2.0):
    mse = ((a.float() - b.float()) ** 2).mean().item()
    return 10 * math.log10(peak * peak / max(mse, 1e-20))


def rel_max(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-20)).item()


def on_cuda(W):
    return {k: v.cuda() for k, v in W.items()}


# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def vae_prod():
    from deepv_b200.vae import B200VAE
    cfg, W = weights.vae_weights(None, seed=2, encoder=True)     # App. A config, the weights bench.py uses
    assert tuple(cfg["decoder_block_out_channels"]) == (128, 256, 512, 512)
    assert tuple(cfg["decoder_layers_per_block"]) == (3, 3, 3, 3)
    v32 = B200VAE(W, cfg, dtype=torch.float32)
    v32.enable_tiling()
    v16 = B200VAE(W, cfg, dtype=torch.bfloat16)
    v16.enable_tiling()
    return cfg, on_cuda(W), v32, v16


def test_vae_production_tiled_decode_benched_shape(vae_prod):
    """The shape bench.py `--workload unit` times: [1,16,2,48,64] -> [1,3,9,384,512], 6 tiles."""
    cfg, Wc, v32, v16 = vae_prod
    z = torch.randn(1, 16, 2, 48, 64, generator=torch.Generator().manual_seed(21)).cuda()
    with exact_fp32():
        ref = vae_ref.tiled_decode(Wc, cfg, z, 256, 1, True)
    assert ref.shape == (1, 3, 9, 384, 512)
    for name, v, zz in (("fp32 io", v32, z), ("bf16 io", v16, z.bfloat16())):
        y = v.decode(zz, temporal_chunk=True, window_size=1, tile_sample_min_size=256).sample
        torch.cuda.synchronize()
        assert y.shape == ref.shape and torch.isfinite(y).all()
        p_all = psnr(y, ref)
        # the blended seams: rows 192-256 (vertical blend) and columns 192-256 / 384-448 (horizontal blends)
        p_v = psnr(y[..., 192:256, :], ref[..., 192:256, :])
        p_h = min(psnr(y[..., 192:256], ref[..., 192:256]), psnr(y[..., 384:448], ref[..., 384:448]))
        print(f"production VAE tiled decode ({name}): PSNR {p_all:.1f} dB, seams v {p_v:.1f} h {p_h:.1f} dB, "
              f"max abs {(y.float() - ref).abs().max().item():.3e}, ref absmax {ref.abs().max().item():.2f}")
        assert min(p_all, p_v, p_h) >= PSNR_FLOOR_DB


def test_vae_production_single_tile_eight_latent_frames(vae_prod):
    """One full 32x32 tile, 8 latent frames = windows (2,1,1,1,1,1,1) with the 2-frame causal caches
    (vae.py:903-920,238-249) -> 57 frames; the decode of a whole rollout iteration per tile."""
    cfg, Wc, v32, v16 = vae_prod
    z = torch.randn(1, 16, 8, 32, 32, generator=torch.Generator().manual_seed(22)).cuda()
    with exact_fp32():
        ref = vae_ref.tiled_decode(Wc, cfg, z, 256, 1, True)
    assert ref.shape == (1, 3, 57, 256, 256)
    for name, v, zz in (("fp32 io", v32, z), ("bf16 io", v16, z.bfloat16())):
        y = v.decode(zz, temporal_chunk=True, window_size=1, tile_sample_min_size=256).sample
        torch.cuda.synchronize()
        assert y.shape == ref.shape and torch.isfinite(y).all()
        per_frame = [psnr(y[:, :, t], ref[:, :, t]) for t in range(57)]
        print(f"production VAE single tile x 8 latent frames ({name}): PSNR {psnr(y, ref):.1f} dB, "
              f"worst frame {min(per_frame):.1f} dB (frame {per_frame.index(min(per_frame))})")
        assert min(per_frame) >= PSNR_FLOOR_DB


@pytest.mark.parametrize("first_frame", [1, 8, 25, 56])
def test_vae_trimmed_decode_is_bit_identical_on_the_kept_frames(vae_prod, first_frame):
    """`decode(z, first_frame=k)` (what a continuation iteration of the rollout uses, k = 25: the reference decodes all
    57 frames and drops the first 25, pipeline.py:327-328) computes, layer by layer, only the frames the kept ones depend
    on — two frames of look-back per causal 3x3x3 conv — so the kept frames must equal the full decode bit for bit."""
    cfg, Wc, v32, v16 = vae_prod
    z = torch.randn(1, 16, 8, 48, 64, generator=torch.Generator().manual_seed(23)).cuda().bfloat16()   # 6 tiles, 57 frames
    full = v16.decode(z, temporal_chunk=True, window_size=1, tile_sample_min_size=256).sample
    part = v16.decode(z, temporal_chunk=True, window_size=1, tile_sample_min_size=256, first_frame=first_frame).sample
    again = v16.decode(z, temporal_chunk=True, window_size=1, tile_sample_min_size=256).sample         # the plan is reset
    torch.cuda.synchronize()
    assert part.shape == full.shape == (1, 3, 57, 384, 512)
    assert torch.equal(part[:, :, first_frame:], full[:, :, first_frame:])
    assert not part[:, :, :first_frame].any()
    assert torch.equal(again, full)


@pytest.mark.parametrize("video", [(1, 3, 1, 384, 512), (1, 3, 9, 384, 512)])
def test_vae_production_encode(vae_prod, video):
    """`vae.encode(x)` (vae.py:844-883,954-987,630-689) at the production widths: moments <= 2e-2."""
    cfg, Wc, v32, _ = vae_prod
    x = torch.randn(*video, generator=torch.Generator().manual_seed(23)).cuda()
    with exact_fp32():
        ref = vae_ref.tiled_encode(Wc, cfg, x)
    m = v32.encode(x).latent_dist.parameters
    torch.cuda.synchronize()
    assert m.shape == ref.shape
    err = rel_max(m, ref)
    print(f"production VAE encode {video}: moments max|a-b|/max|ref| = {err:.3e}")
    assert err <= 2e-2


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


@pytest.mark.parametrize("T,H,W,Cin,Cout,ks,store,drop", [
    (2, 24, 32, 512, 2048, 3, 1, 0),    # up0 spatial up-sampler on an edge tile, generic split-K path
    (3, 48, 64, 512, 2048, 3, 1, 0),    # up1 spatial up-sampler (halo, pixel shuffle)
    (2, 48, 64, 512, 1024, 3, 2, 1),    # up0 temporal up-sampler, first frame dropped
    (3, 96, 128, 512, 1024, 3, 2, 0),   # up1 temporal up-sampler
    (3, 48, 64, 512, 512, 3, 0, 0),     # up1 resnet conv @ T3
    (2, 24, 32, 512, 512, 3, 0, 0),     # mid / up0 resnet conv, split-K
    (5, 96, 128, 512, 256, 1, 0, 0),    # up2 shortcut 1x1x1
    (9, 192, 256, 256, 128, 1, 0, 0),   # up3 shortcut 1x1x1, swapped operands
    (5, 96, 128, 512, 256, 3, 0, 0),    # up2 resnet 0 conv1
    (5, 192, 256, 256, 512, 3, 2, 0),   # up2 temporal up-sampler
    (9, 192, 256, 256, 128, 3, 0, 0),   # up3 resnet 0 conv1
])
def test_production_conv_shapes(lib, T, H, W, Cin, Cout, ks, store, drop):
    from deepv_b200 import _lib
    torch.manual_seed(5)
    B = 1
    x = (torch.randn(B, T, H, W, Cin, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(Cout, ks ** 3, Cin, device="cuda") * (1.0 / math.sqrt(ks ** 3 * Cin))).bfloat16()
    bias = torch.randn(max(Cout, 32), device="cuda") * 0.1
    res = (torch.randn(B, T, H, W, Cout, device="cuda") * 0.5).bfloat16() if store == 0 else None
    oshape = {0: (B, T, H, W, Cout), 1: (B, T, 2 * H, 2 * W, Cout // 4), 2: (B, 2 * T - drop, H, W, Cout // 2)}[store]
    out = torch.zeros(oshape, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.dv_conv3d_cl(_p(x), _p(w), _p(bias), _p(res), _p(out), B, T, H, W, Cin, Cout, Cout, ks, store, drop,
                                None))
    torch.cuda.synchronize()
    xn = x.float().permute(0, 4, 1, 2, 3)
    wn = w.float().view(Cout, ks, ks, ks, Cin).permute(0, 4, 1, 2, 3)
    pad = ks // 2
    with exact_fp32():
        y = torch.nn.functional.conv3d(torch.nn.functional.pad(xn, (pad, pad, pad, pad, ks - 1, 0)), wn, bias[:Cout])
    if store == 0:
        ref = (y + res.float().permute(0, 4, 1, 2, 3)).permute(0, 2, 3, 4, 1)
    elif store == 1:
        Cq = Cout // 4
        ref = y.view(B, 2, 2, Cq, T, H, W).permute(0, 4, 5, 1, 6, 2, 3).reshape(B, T, 2 * H, 2 * W, Cq)
    else:
        Ch = Cout // 2
        ref = y.view(B, 2, Ch, T, H, W).permute(0, 3, 1, 4, 5, 2).reshape(B, 2 * T, H, W, Ch)[:, drop:]
    err = rel_max(out, ref)
    print(f"conv T{T} {H}x{W} {Cin}->{Cout} k{ks} store {store}: {err:.2e}")
    assert err <= 8e-3


# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def dit24():
    from deepv_b200.mmdit import B200MMDiT
    cfg, W = weights.mmdit_weights(None, seed=1)
    assert cfg["num_layers"] == 24
    return cfg, on_cuda(W), B200MMDiT(W, cfg, out_dtype=torch.float32)


FULL_DEPTH_CASES = {
    # first iteration, units 1-2, stage 2: B=2, L = 1536 + 77
    "stage2_b2": dict(clips=[(1, 48, 64), (1, 48, 64)], B=2, hist=None, lens=[1, 12], t=289.84625244140625, seed=51),
    # steady state, unit 8 (=7 of the iteration), stage 2: B=3, history, three resolutions, L = 1968 + 269
    "unit8_stage2_b3_hist": dict(clips=[(5, 12, 16), (1, 24, 32), (1, 48, 64), (1, 48, 64)], B=3, hist=(48, 64),
                                 lens=[1, 12, 12], t=193.6925048828125, seed=52),
    # steady state, unit 8, stage 0: B=3, history, L = 384 + 269 (the small-M path)
    "unit8_stage0_b3_hist": dict(clips=[(6, 12, 16), (1, 12, 16), (1, 12, 16)], B=3, hist=(48, 64),
                                 lens=[1, 12, 12], t=872.1279907226562, seed=53),
}


@pytest.mark.parametrize("name", list(FULL_DEPTH_CASES))
def test_mmdit_full_depth_large_layouts(dit24, name):
    cfg, Wc, model = dit24
    case = FULL_DEPTH_CASES[name]
    inp = cases.mmdit_inputs(case)
    dev = "cuda"
    clips = [c.to(dev) for c in inp["clips"]]
    hist = inp["hist"].to(dev) if inp["hist"] is not None else None
    hmask = inp["hmask"].to(dev) if inp["hmask"] is not None else None
    with exact_fp32():
        ref = mmdit_ref.mmdit_forward(Wc, cfg, clips, inp["t"].to(dev), inp["enc"].to(dev), inp["mask"].to(dev),
                                      inp["pooled"].to(dev), hist, hmask, 2 if hist is not None else None)
    for dtype in (torch.float32, torch.bfloat16):
        y = model(sample=[[c.to(dtype) for c in clips]], timestep_ratio=inp["t"].to(dev),
                  encoder_hidden_states=inp["enc"].to(dev), encoder_attention_mask=inp["mask"].to(dev),
                  pooled_projections=inp["pooled"].to(dev),
                  history=hist.to(dtype) if hist is not None else None, history_mask=hmask,
                  history_downsample_ratio=2 if hist is not None else None)[0]
        torch.cuda.synchronize()
        assert y.shape == ref.shape and torch.isfinite(y).all()
        err = rel_max(y, ref)
        print(f"24 blocks {name} ({dtype}): max|a-b|/max|ref| = {err:.3e} (ref absmax {ref.abs().max().item():.3f})")
        assert err <= DENOISER_TOL


# ---------------------------------------------------------------------------------------------
# A whole first iteration at the DEMO SHAPE with the REAL-SIZE models (BASELINE.json north_star: "decoded frames within
# a stated PSNR floor over a full rollout"): generate_i2v = VAE encode of the input frame, 8
# autoregressive units x 3 stages x 5 steps = 120 full-depth forwards with CFG, ray map -> poses, two 57-frame tiled
# decodes, at 384x512 — against the oracle's generate_i2v (oracle/rollout_ref.py, pinned to the real reference's
# generate() on CPU) run in fp32 on CUDA on the same noise tape.
# measured on B200 (profiles/r02e_pytest_gpu.log): fp32 latents 47.8 dB, per-unit latent error <= 9.1e-3; the bench's
# bf16 latents (the reference's own GPU configuration, pipeline.py:190,486) 43.0 dB, worst frame 40.0 dB, <= 2.1e-2
FULL_ROLLOUT_PSNR_FLOOR_DB = {torch.float32: 40.0, torch.bfloat16: 38.0}
FULL_ROLLOUT_LATENT_TOL = {torch.float32: 2e-2, torch.bfloat16: 3e-2}


class _CudaTape:
    """The seeded CPU tape of tests/golden/rollout_cases.py, handing its draws out on the device."""

    def __init__(self, tape):
        self.tape = tape

    def randn(self, shape):
        return self.tape.randn(shape).cuda()

    def block(self, *a):
        return self.tape.block(*a).cuda()


def test_full_size_first_iteration_vs_cuda_fp32_oracle(dit24, vae_prod):
    from deepv_b200.pipeline import B200Pipeline
    from deepv_b200.rollout import B200Rollout, PromptCache
    from deepv_b200.scheduler import B200Scheduler
    from oracle import rollout_ref, scheduler_ref
    from tests.golden import rollout_cases as rc
    cfg, Wc, model = dit24
    vcfg, VWc, _, _ = vae_prod
    case = dict(rc.ROLLOUT, height=384, width=512, seed=91)
    model_cfg = dict(case["model_cfg"], num_inference_steps=5)
    embeds = rc.text_embeds(case)
    embeds_c = {k: {n: t.cuda() for n, t in e.items()} for k, e in embeds.items()}
    m = rollout_ref.RolloutModels(cfg, Wc, vcfg, VWc, scheduler_ref.pyramid_tables(**cases.SCHEDULER_KW), embeds_c, model_cfg)
    m._pos = m.pos_table().cuda()
    motion = ["w", "a", "w", "d", "w", "s", "a", "w"]
    frame = rc.first_frame(case)
    tape = rc.RecordingTape(case["seed"] + 3)
    with exact_fp32():
        o_img, o_disp, o_t3, o_t2, o_lat = rollout_ref.generate_i2v(m, motion, frame.unsqueeze(0).cuda(), None, None, None,
                                                                    _CudaTape(tape), [5, 5, 5])
    assert o_img.shape == (1, 3, 57, 384, 512)
    from deepv_b200.vae import B200VAE
    W_cpu = {k: v.cpu() for k, v in VWc.items()}
    for dtype in (torch.float32, torch.bfloat16):
        vae = B200VAE(W_cpu, vcfg, dtype=dtype)
        vae.enable_tiling()
        model.out_dtype = dtype
        pipe = B200Pipeline(model, vae, B200Scheduler(**cases.SCHEDULER_KW), model_cfg=model_cfg, torch_dtype=dtype)
        ro = B200Rollout(pipe, PromptCache(embeds))
        replay = rc.ReplayTape(tape.draws, tape.calls, 0)
        image, disparity, t3, t2, lat = ro.generate_i2v(motion, True, ro.frames_from_uint8(frame.unsqueeze(0)), None, None, None,
                                                        temp=8, num_inference_steps=5, noise=replay, return_latents=True)
        torch.cuda.synchronize()
        assert image.shape == o_img.shape and torch.isfinite(image).all() and torch.isfinite(disparity).all()
        per_unit = [rel_max(lat[:, :, u], o_lat[:, :, u]) for u in range(lat.shape[2])]
        p_img, p_disp = psnr(image, o_img), psnr(disparity, o_disp)
        worst = min(psnr(image[:, :, t], o_img[:, :, t]) for t in range(57))
        print(f"full-size iteration ({dtype}): PSNR image {p_img:.1f} dB (worst frame {worst:.1f}), disparity {p_disp:.1f} dB; "
              f"latent max|a-b|/max|ref| per unit " + " ".join(f"{e:.1e}" for e in per_unit) +
              f"; trans3d {rel_max(t3, o_t3):.1e}")
        assert p_img >= FULL_ROLLOUT_PSNR_FLOOR_DB[dtype] and p_disp >= FULL_ROLLOUT_PSNR_FLOOR_DB[dtype]
        assert max(per_unit) <= FULL_ROLLOUT_LATENT_TOL[dtype]
        del vae, pipe, ro
    model.out_dtype = torch.float32
