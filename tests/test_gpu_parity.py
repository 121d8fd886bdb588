"""GPU parity tests: the sm_100a path (through the C ABI) against the CPU oracle and the golden
vectors of the real reference.  Tolerances follow BASELINE.json north_star:
  * scheduler / CFG / stage transition / indexing: bit-exact
  * denoiser forward: max|a-b| / max|ref| <= 2e-2 (bf16 operands vs fp32 reference)
  * decoded frames: PSNR >= 40 dB over [-1, 1]
"""
import math
from pathlib import Path

import pytest
import torch

from oracle import mmdit_ref, scheduler_ref, vae_ref, weights
from tests.golden import cases

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden"
DENOISER_TOL = 2e-2
PSNR_FLOOR_DB = 40.0


def rel_max(a, b):
    return ((a.float().cpu() - b.float().cpu()).abs().max() / b.float().abs().max()).item()


def psnr(a, b, peak=2.0):
    mse = ((a.float().cpu() - b.float().cpu()) ** 2).mean().item()
    return 10 * math.log10(peak * peak / max(mse, 1e-20))


@pytest.fixture(scope="module")
def dit2():
    from deepv_b200.mmdit import B200MMDiT
    cfg, W = weights.mmdit_weights(dict(num_layers=2), seed=1)
    return cfg, W, B200MMDiT(W, cfg, out_dtype=torch.float32)


def _run_case(model, case, dtype=torch.float32):
    inp = cases.mmdit_inputs(case)
    dev = "cuda"
    y = model(sample=[[c.to(dev, dtype) for c in inp["clips"]]], timestep_ratio=inp["t"].to(dev),
              encoder_hidden_states=inp["enc"].to(dev), encoder_attention_mask=inp["mask"].to(dev),
              pooled_projections=inp["pooled"].to(dev),
              history=inp["hist"].to(dev, dtype) if inp["hist"] is not None else None,
              history_mask=inp["hmask"].to(dev) if inp["hmask"] is not None else None,
              history_downsample_ratio=2 if inp["hist"] is not None else None)[0]
    torch.cuda.synchronize()
    return y


@pytest.mark.parametrize("name", ["two_block_b2", "two_block_b3_hist"])
def test_mmdit_vs_reference_golden(dit2, name):
    cfg, W, model = dit2
    gold = torch.load(G / "mmdit_golden.pt")[name]
    y = _run_case(model, cases.MMDIT_CASES[name])
    assert y.shape == gold.shape
    assert torch.isfinite(y).all()
    err = rel_max(y, gold)
    print(f"{name}: max|a-b|/max|ref| = {err:.3e}")
    assert err <= DENOISER_TOL


def test_mmdit_three_block_mixed_golden():
    from deepv_b200.mmdit import B200MMDiT
    case = cases.MMDIT_CASES["three_block_mixed"]
    cfg, W = weights.mmdit_weights(case["cfg"], seed=case["wseed"])
    model = B200MMDiT(W, cfg, out_dtype=torch.float32)
    gold = torch.load(G / "mmdit_golden.pt")["three_block_mixed"]
    err = rel_max(_run_case(model, case), gold)
    print(f"three_block_mixed: {err:.3e}")
    assert err <= DENOISER_TOL


def test_mmdit_bf16_io_and_padding_invariance(dit2):
    """bf16 latents in / bf16 out; padded text tokens must not influence the output
    (SURVEY.md §4 invariance: dead tokens never reach live ones)."""
    cfg, W, model = dit2
    case = cases.MMDIT_CASES["two_block_b2"]
    inp = cases.mmdit_inputs(case)
    dev = "cuda"

    def run(enc):
        return model(sample=[[c.to(dev, torch.bfloat16) for c in inp["clips"]]],
                     timestep_ratio=inp["t"].to(dev), encoder_hidden_states=enc.to(dev),
                     encoder_attention_mask=inp["mask"].to(dev),
                     pooled_projections=inp["pooled"].to(dev))[0]
    y1 = run(inp["enc"])
    enc2 = inp["enc"].clone()
    enc2[0, 1:] = 7.0   # row 0 has one live token; scribble over the padded ones
    enc2[1, 12:] = -3.0
    y2 = run(enc2)
    torch.cuda.synchronize()
    assert torch.equal(y1, y2)
    gold = torch.load(G / "mmdit_golden.pt")["two_block_b2"]
    assert rel_max(y1, gold) <= DENOISER_TOL


def test_conditioning_cache_hits_are_bit_exact(monkeypatch):
    """The adaLN modulation tables are memoised per (timestep, pooled embedding) row on the device (csrc/mmdit.cu,
    cond_lookup / cond_finish): found rows, duplicate rows inside one CFG batch, evictions and a cache-less model
    must all give bit-identical outputs."""
    from deepv_b200.mmdit import B200MMDiT
    cfg, W = weights.mmdit_weights(dict(num_layers=2), seed=1)
    case = cases.MMDIT_CASES["two_block_b3_hist"]
    inp = cases.mmdit_inputs(case)
    inp["pooled"][2] = inp["pooled"][1]            # rows 1 and 2 of a 3-branch CFG batch share prompt and timestep
    dev = "cuda"

    def run(model, t):
        y = model(sample=[[c.to(dev) for c in inp["clips"]]], timestep_ratio=torch.full((3,), t, device=dev),
                  encoder_hidden_states=inp["enc"].to(dev), encoder_attention_mask=inp["mask"].to(dev),
                  pooled_projections=inp["pooled"].to(dev), history=inp["hist"].to(dev), history_mask=inp["hmask"].to(dev),
                  history_downsample_ratio=2)[0]
        torch.cuda.synchronize()
        return y.clone()

    monkeypatch.setenv("DV_MOD_CACHE_SLOTS", "0")
    plain = B200MMDiT(W, cfg, out_dtype=torch.float32)
    monkeypatch.setenv("DV_MOD_CACHE_SLOTS", "5")  # tiny: forces slot conflicts and evictions
    cached = B200MMDiT(W, cfg, out_dtype=torch.float32)
    ts = [936.0639953613281, 872.1279907226562, 744.0, 654.5895004272461, 386.0, 289.84625244140625, 97.53875732421875]
    want = {t: run(plain, t) for t in ts}
    lib = cached.lib
    for rnd in range(3):                           # round 0: misses; later rounds: hits and re-computed evictions
        for t in (ts if rnd != 1 else reversed(ts)):
            assert torch.equal(run(cached, t), want[t]), (rnd, t)
    n0 = lib.dv_launch_count()
    run(cached, ts[0])
    assert lib.dv_launch_count() > n0              # (the cached forward still issues its launches; they return early)


def test_persistent_block_kernel_matches_the_per_launch_path(monkeypatch):
    """DV_MMDIT_PBK=1 (experimental, csrc/gemm.cu pbk_kernel): the LN / GEMM steps between two attentions as phases of one
    launch behind grid barriers.  Same arithmetic up to the split-K re-association of the small GEMMs."""
    from deepv_b200.mmdit import B200MMDiT
    case = cases.MMDIT_CASES["three_block_mixed"]
    cfg, W = weights.mmdit_weights(case["cfg"], seed=case["wseed"])
    model = B200MMDiT(W, cfg, out_dtype=torch.float32)
    gold = torch.load(G / "mmdit_golden.pt")["three_block_mixed"]
    plain = _run_case(model, case)
    monkeypatch.setenv("DV_MMDIT_PBK", "1")
    monkeypatch.setenv("DV_PBK_MAX_ROWS", "4096")
    fused = _run_case(model, case)
    again = _run_case(model, case)
    assert torch.equal(fused, again)                      # bit-reproducible
    assert rel_max(fused, gold) <= DENOISER_TOL
    assert rel_max(fused, plain) <= 2e-3
    print(f"persistent block kernel: vs golden {rel_max(fused, gold):.3e}, vs per-launch path {rel_max(fused, plain):.3e}")


def test_mmdit_full_depth_vs_oracle():
    """24 blocks (the real depth), first-unit stage-1 layout, against the fp32 oracle."""
    from deepv_b200.mmdit import B200MMDiT
    cfg, W = weights.mmdit_weights(None, seed=1)
    model = B200MMDiT(W, cfg, out_dtype=torch.float32)
    case = dict(clips=[(1, 24, 32), (1, 24, 32)], B=2, hist=None, lens=[1, 12], t=654.5895004272461,
                seed=41)
    inp = cases.mmdit_inputs(case)
    with torch.no_grad():
        ref = mmdit_ref.mmdit_forward(W, cfg, inp["clips"], inp["t"], inp["enc"], inp["mask"], inp["pooled"])
    y = _run_case(model, case)
    err = rel_max(y, ref)
    print(f"full depth: max|a-b|/max|ref| = {err:.3e} (ref absmax {ref.abs().max():.3f})")
    assert err <= DENOISER_TOL


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("nb", [1, 2, 3])
def test_cfg_euler_bit_exact(dtype, nb):
    from deepv_b200.scheduler import B200Scheduler
    s = B200Scheduler(**cases.SCHEDULER_KW)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 38, 1, 48, 64, generator=g).to(dtype)
    pred = torch.randn(nb, 38, 1, 48, 64, generator=g).to(dtype)
    tb = scheduler_ref.pyramid_tables(**cases.SCHEDULER_KW)
    _, sig = scheduler_ref.stage_schedule(tb, 5, 1)
    s.set_timesteps(5, 1)
    cur_ref, cur = x, x.cuda()
    for k in range(5):
        guided = scheduler_ref.cfg_combine(pred, 3.5, 6.0)
        cur_ref = scheduler_ref.euler_step(cur_ref, guided, float(sig[k]), float(sig[k + 1]))
        cur = s.cfg_step(pred.cuda(), cur, nb, 3.5, 6.0)
        assert torch.equal(cur.cpu(), cur_ref), f"step {k}"


def test_scheduler_step_matches_reference_golden():
    from deepv_b200.scheduler import B200Scheduler
    for dtype in (torch.bfloat16, torch.float32):
        gold = torch.load(G / "scheduler_step_golden.pt")[str(dtype)]
        x, v = cases.step_inputs(dtype)
        s = B200Scheduler(**cases.SCHEDULER_KW)
        s.set_timesteps(5, 0, device="cuda")
        cur = x.cuda()
        for k in range(5):
            cur = s.step(model_output=v.cuda(), timestep=s.timesteps[k], sample=cur).prev_sample
            assert torch.equal(cur.cpu(), gold[k])


def test_stage_renoise_bit_exact_and_block_noise_cov(lib):
    from deepv_b200 import _lib
    g = torch.Generator().manual_seed(6)
    for dtype in (torch.bfloat16, torch.float32):
        lo = torch.randn(38, 12, 16, generator=g).to(dtype)
        nz = torch.randn(38, 24, 32, generator=g).to(dtype)
        a, b = scheduler_ref.renoise_coefficients(0.6669999957084656, 0.3333)
        out = torch.empty(38, 24, 32, device="cuda", dtype=dtype)
        lo_d, nz_d = lo.cuda(), nz.cuda()
        _lib.check(lib.dv_stage_renoise(lo_d.data_ptr(), nz_d.data_ptr(), out.data_ptr(), 38, 12, 16, a, b,
                                        _lib.dtype_code(dtype), None))
        torch.cuda.synchronize()
        up = torch.nn.functional.interpolate(lo[None], size=(24, 32), mode="nearest")[0]
        assert torch.equal(out.cpu(), a * up + b * nz)
    # distribution-level parity of the block noise (pipeline.py:431-437)
    z = torch.randn(400000, 4, device="cuda")
    out = torch.empty(100, 80, 200, device="cuda")
    _lib.check(lib.dv_block_noise(z.data_ptr(), out.data_ptr(), 100, 80, 200, 0.3333, 0, None))
    blk = out.view(100, 40, 2, 100, 2).permute(0, 1, 3, 2, 4).reshape(-1, 4).double()
    cov = (blk.t() @ blk / blk.shape[0]).cpu()
    assert (cov - scheduler_ref.block_noise_cov(0.3333).double()).abs().max() < 0.02


# ---------------------------------------------------------------------------------------------
def _vae(cfg_over, seed):
    from deepv_b200.vae import B200VAE
    cfg, W = weights.vae_weights(cfg_over, seed=seed)
    v = B200VAE(W, cfg, dtype=torch.float32)
    v.enable_tiling()
    return cfg, W, v


def test_vae_tiled_decode_vs_oracle():
    over = dict(decoder_block_out_channels=(128, 128, 128, 128), encoder_block_out_channels=(128, 128, 128, 128),
                decoder_layers_per_block=(1, 1, 1, 1))
    cfg, W, v = _vae(over, 7)
    z = torch.randn(1, 16, 2, 40, 64, generator=torch.Generator().manual_seed(8))
    with torch.no_grad():
        ref = vae_ref.tiled_decode(W, cfg, z, 256, 1, True)
    y = v.decode(z.cuda(), temporal_chunk=True, window_size=1, tile_sample_min_size=256).sample
    torch.cuda.synchronize()
    assert y.shape == ref.shape == (1, 3, 9, 320, 512)
    p = psnr(y, ref)
    print(f"vae tiled: PSNR {p:.1f} dB, max abs {(y.cpu() - ref).abs().max():.3e}, ref absmax {ref.abs().max():.2f}")
    assert p >= PSNR_FLOOR_DB
    # seams: the blended bands must match as well as the interior
    band = psnr(y[..., 180:260, :], ref[..., 180:260, :])
    assert band >= PSNR_FLOOR_DB


def test_vae_untiled_three_frames_vs_oracle():
    over = dict(decoder_block_out_channels=(128, 128, 128, 128), encoder_block_out_channels=(128, 128, 128, 128),
                decoder_layers_per_block=(1, 1, 1, 1))
    cfg, W, v = _vae(over, 9)
    z = torch.randn(1, 16, 3, 16, 32, generator=torch.Generator().manual_seed(10))
    with torch.no_grad():
        ref = vae_ref.tiled_decode(W, cfg, z, 256, 1, True)
    y = v.decode(z.cuda(), temporal_chunk=True, window_size=1, tile_sample_min_size=256).sample
    assert y.shape == ref.shape == (1, 3, 17, 128, 256)
    p = psnr(y, ref)
    print(f"vae untiled: PSNR {p:.1f} dB")
    assert p >= PSNR_FLOOR_DB


# ---------------------------------------------------------------------------------------------
def test_generate_one_unit_vs_oracle(dit2):
    """3 stages x 2 steps through the public pipeline API with injected block noise."""
    from deepv_b200.pipeline import B200Pipeline
    from deepv_b200.scheduler import B200Scheduler
    cfg, W, model = dit2
    pipe = B200Pipeline(model, None, B200Scheduler(**cases.SCHEDULER_KW), torch_dtype=torch.float32)
    g = torch.Generator().manual_seed(12)
    lat = torch.randn(1, 38, 1, 8, 8, generator=g)
    conds = [[torch.randn(2, 38, 1, 8 * 2 ** i, 8 * 2 ** i, generator=g)] for i in range(3)]
    noise = [torch.randn(1, 38, 1, 16, 16, generator=g), torch.randn(1, 38, 1, 32, 32, generator=g)]
    enc = torch.randn(2, 77, 4096, generator=g)
    pooled = torch.randn(2, 2048, generator=g)
    mask = torch.zeros(2, 77, dtype=torch.long)
    mask[0, :1] = 1
    mask[1, :12] = 1
    tb = scheduler_ref.pyramid_tables(**cases.SCHEDULER_KW)

    def model_fn(clips, tt):
        return mmdit_ref.mmdit_forward(W, cfg, clips, tt.float(), enc, mask, pooled)

    with torch.no_grad():
        ref = scheduler_ref.generate_one_unit(model_fn, tb, lat, conds, noise, 2, [2, 2, 2], 3.5, 6.0,
                                              timestep_dtype=torch.float32)
    out = pipe.generate_one_unit(lat.cuda(), None, [[c.cuda() for c in cl] for cl in conds], enc, mask, pooled,
                                 [2, 2, 2], block_noise=noise, timestep_dtype=torch.float32)
    torch.cuda.synchronize()
    for i in range(3):
        e = rel_max(out[i], ref[i])
        print(f"unit stage {i}: latent max|a-b|/max|ref| = {e:.3e}")
        assert out[i].shape == ref[i].shape
        assert e <= 3e-2  # ~2% latent drift per unit for bf16 vs fp32


def test_generate_one_unit_no_need_depth_vs_oracle(dit2):
    """`model_cfg['no_need_depth']` (pipeline.py:476-478): channels 16.. of every clip are cleared before the
    denoiser sees them; the oracle's version of the flag is pinned to the live reference on the CPU side."""
    from deepv_b200.pipeline import B200Pipeline
    from deepv_b200.scheduler import B200Scheduler
    cfg, W, model = dit2
    pipe = B200Pipeline(model, None, B200Scheduler(**cases.SCHEDULER_KW), model_cfg=dict(no_need_depth=True),
                        torch_dtype=torch.float32)
    g = torch.Generator().manual_seed(14)
    lat = torch.randn(1, 38, 1, 8, 8, generator=g)
    conds = [[torch.randn(2, 38, 1, 8 * 2 ** i, 8 * 2 ** i, generator=g)] for i in range(3)]
    noise = [torch.randn(1, 38, 1, 16, 16, generator=g), torch.randn(1, 38, 1, 32, 32, generator=g)]
    enc = torch.randn(2, 77, 4096, generator=g)
    pooled = torch.randn(2, 2048, generator=g)
    mask = torch.zeros(2, 77, dtype=torch.long)
    mask[0, :1] = 1
    mask[1, :9] = 1
    tb = scheduler_ref.pyramid_tables(**cases.SCHEDULER_KW)

    def model_fn(clips, tt):
        return mmdit_ref.mmdit_forward(W, cfg, clips, tt.float(), enc, mask, pooled)

    with torch.no_grad():
        ref = scheduler_ref.generate_one_unit(model_fn, tb, lat, conds, noise, 2, [1, 2, 1], 3.5, 6.0,
                                              timestep_dtype=torch.float32, no_need_depth=True)
        plain = scheduler_ref.generate_one_unit(model_fn, tb, lat, conds, noise, 2, [1, 2, 1], 3.5, 6.0,
                                                timestep_dtype=torch.float32)
    lat_dev = lat.cuda()
    out = pipe.generate_one_unit(lat_dev, None, [[c.cuda() for c in cl] for cl in conds], enc, mask, pooled,
                                 [1, 2, 1], block_noise=noise, timestep_dtype=torch.float32)
    torch.cuda.synchronize()
    assert torch.equal(lat_dev.cpu(), lat)           # with CFG the caller's latents are left alone, as in the reference
    for i in range(3):
        e = rel_max(out[i], ref[i])
        print(f"no_need_depth stage {i}: {e:.3e} (flag changes the result by {rel_max(plain[i], ref[i]):.2e})")
        assert e <= 3e-2
    assert rel_max(plain[2], ref[2]) > 5 * rel_max(out[2], ref[2])


# ---------------------------------------------------------------------------------------------
# VAE encode
@pytest.mark.parametrize("video", [(1, 3, 1, 64, 128), (1, 3, 9, 320, 320), (1, 3, 17, 384, 512)])
def test_vae_tiled_encode_vs_oracle(video):
    """`vae.encode(x)` with tiling on (vae.py:844-883,954-987): moments against the fp32 oracle
    restatement (itself pinned to the real reference, tests/test_oracle_golden.py); tolerance as for
    the denoiser: max|a-b| / max|ref| <= 2e-2 with bf16 activations."""
    from deepv_b200.vae import B200VAE
    over = dict(decoder_block_out_channels=(128, 128, 128, 128), encoder_block_out_channels=(128, 128, 256, 256),
                decoder_layers_per_block=(1, 1, 1, 1), encoder_layers_per_block=(1, 2, 1, 1))
    cfg, W = weights.vae_weights(over, seed=11, encoder=True)
    v = B200VAE(W, cfg, dtype=torch.float32)
    v.enable_tiling()
    x = torch.randn(*video, generator=torch.Generator().manual_seed(12))
    with torch.no_grad():
        ref = vae_ref.tiled_encode(W, cfg, x)
    dist = v.encode(x.cuda()).latent_dist
    torch.cuda.synchronize()
    m = dist.parameters
    assert m.shape == ref.shape
    err = rel_max(m, ref)
    print(f"vae encode {video}: moments max|a-b|/max|ref| = {err:.3e} (ref absmax {ref.abs().max():.2f})")
    assert err <= 2e-2
    # the sample with an injected draw: mean + exp(0.5 clamp(logvar)) * noise, bit-for-bit in fp32
    noise = torch.randn(dist.mean.shape, generator=torch.Generator().manual_seed(13))
    z = dist.sample_with_noise(noise.cuda())
    want = vae_ref.gaussian_sample(m.cpu(), noise)
    assert (z.cpu() - want).abs().max().item() <= 1e-5 * max(1.0, want.abs().max().item())


# ---------------------------------------------------------------------------------------------
# pyramid resampling
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_resize_half_bit_exact_vs_aten(lib, dtype):
    """get_pyramid_latent (pipeline.py:226-240) and the initial noise pyramid (:554-557): bilinear
    halving (and `* 2`) bit-equal to F.interpolate on the same device."""
    import ctypes as C

    import torch.nn.functional as F

    from deepv_b200 import _lib
    g = torch.Generator().manual_seed(21)
    x = (torch.randn(2, 38, 3, 48, 64, generator=g) * 3).to(dtype).cuda()
    b, c, t, h, w = x.shape
    for scale in (1.0, 2.0):
        out = torch.empty(b, c, t, h // 2, w // 2, device="cuda", dtype=dtype)
        _lib.check(lib.dv_resize_half(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), b * c * t, h, w, scale,
                                      _lib.dtype_code(dtype), None))
        torch.cuda.synchronize()
        ref = F.interpolate(x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w), size=(h // 2, w // 2), mode="bilinear")
        ref = ref.view(b, t, c, h // 2, w // 2).permute(0, 2, 1, 3, 4)
        if scale != 1.0:
            ref = ref * scale
        assert torch.equal(out, ref.contiguous()), (dtype, scale)
