"""GPU parity of the rollout around the hot path against the oracle restatement
of `InferencePipeline.generate` / `generate_i2v` (oracle/rollout_ref.py, pinned to the real reference by
tests/test_oracle_rollout.py), on the same seeded noise tape.

  * feedback kernels (uint8 round trip, disparity post/renorm, ray map <-> pose): same inputs on both
    sides; integer work bit-exact, fp32 work to float round-off;
  * one full `generate_i2v` iteration (8 units x 3 stages, both decodes): decoded frames against the
    fp32 oracle with the PSNR floor below; the second iteration teacher-forced with the oracle's inputs;
  * the whole two-iteration `generate`: draws consumed in the reference's order, result layout, and
    every feedback step re-checked against the oracle applied to the GPU's own iteration-0 output.
"""
import math

import pytest
import torch

from oracle import rollout_ref
from tests.golden import cases, rollout_cases as rc

pytestmark = pytest.mark.gpu

# bf16 tensor-core operands vs the fp32 oracle through 8 autoregressive units and both VAE decodes; measured on
# B200: 49 dB / 6e-3 per-unit latent error (profiles/r01_rollout_parity.log).  The floor is the decode's own.
ROLLOUT_PSNR_FLOOR_DB = 40.0
ROLLOUT_LATENT_TOL = 3e-2


def psnr(a, b, peak=2.0):
    mse = ((a.float().cpu() - b.float().cpu()) ** 2).mean().item()
    return 10 * math.log10(peak * peak / max(mse, 1e-20))


def rel_max(a, b):
    return ((a.float().cpu() - b.float().cpu()).abs().max() / b.float().abs().max().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def gpu_rollout():
    from deepv_b200.mmdit import B200MMDiT
    from deepv_b200.pipeline import B200Pipeline
    from deepv_b200.rollout import B200Rollout, PromptCache
    from deepv_b200.scheduler import B200Scheduler
    from deepv_b200.vae import B200VAE
    case = rc.ROLLOUT_GPU
    m = rc.build_oracle_models(case)
    dit = B200MMDiT(m.dit_W, m.dit_cfg, out_dtype=torch.float32)
    vae = B200VAE(m.vae_W, m.vae_cfg, dtype=torch.float32)
    vae.enable_tiling()
    pipe = B200Pipeline(dit, vae, B200Scheduler(**cases.SCHEDULER_KW), model_cfg=case["model_cfg"],
                        torch_dtype=torch.float32)
    return case, m, B200Rollout(pipe, PromptCache(rc.text_embeds(case)))


@pytest.fixture(scope="module")
def oracle_run(gpu_rollout):
    case, m, _ = gpu_rollout
    tape = rc.RecordingTape(case["seed"] + 3)
    trace = []
    with torch.no_grad():
        res = rollout_ref.generate(m, rc.first_frame(case), rc.prompts(case), tape, [1, 1, 1], trace)
    return res, trace, tape


# ---------------------------------------------------------------------------------------------
# feedback kernels on identical inputs
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_requantise_bit_exact(gpu_rollout, dtype):
    """pipeline.py:339-344 + 564-568: the uint8 frames are exact (truncation included), and so is the
    normalised tensor the VAE encoder then sees."""
    _, _, ro = gpu_rollout
    g = torch.Generator().manual_seed(5)
    v = (torch.rand(1, 3, 30, 64, 96, generator=g) * 2.6 - 1.3).to(dtype)
    v[0, 0, 7, 0, :8] = torch.tensor([-1.0, 1.0, 0.0, 1 / 255, 2 / 255 - 1, 0.999, -0.999, 0.5]).to(dtype)
    out, u8 = ro.requantise(v.cuda(), 5, 25, want_u8=True)
    torch.cuda.synchronize()
    # the reference's op order (pipeline.py:341): `* 0.5 + 0.5` and the clamp in the decode dtype, THEN fp32
    last = v[0, :, 5:30].permute(1, 2, 3, 0)
    want_u8 = (torch.clamp(last * 0.5 + 0.5, 0, 1).to(torch.float32) * 255).to(torch.uint8)
    assert torch.equal(u8.cpu(), want_u8)
    assert torch.equal(out.cpu(), rollout_ref.frames_to_input(want_u8))
    assert torch.equal(ro.frames_from_uint8(want_u8).cpu(), rollout_ref.frames_to_input(want_u8))


def test_disparity_feedback_vs_oracle(gpu_rollout):
    """pipeline.py:311-313 and :346-350 / :399-401 with the scale kept as a device scalar."""
    _, m, ro = gpu_rollout
    g = torch.Generator().manual_seed(6)
    raw = torch.randn(1, 3, 33, 64, 96, generator=g) * 0.7
    fb = rollout_ref.Feedback(m)
    scale_prev = torch.tensor([1.7])
    for prev in (None, scale_prev):
        fb.scale = 1.0 if prev is None else prev
        want = torch.clamp(raw.mean(dim=1, keepdim=True).repeat(1, 3, 1, 1, 1) * 0.5 + 0.5, 0, 1) ** 2 / fb.scale / 0.95
        got = ro.disparity_post(raw.cuda(), None if prev is None else prev.cuda())
        assert (got.cpu() - want).abs().max().item() <= 2e-7 * want.abs().max().item()
    d = want[:, :, -25:]
    want_scale = 1 / d[:, :, 0].max()
    want_in = torch.sqrt(d * want_scale * 0.95) * 2 - 1
    scale = torch.empty(1, device="cuda")
    got_in = ro.disparity_renorm(got, 33 - 25, 25, scale, True, False)
    torch.cuda.synchronize()
    assert abs(scale.item() - want_scale.item()) <= 1e-6 * want_scale.item()
    assert (got_in.cpu() - want_in).abs().max().item() <= 1e-5
    one = want[:, :, 3:4]
    want_h = torch.clamp(torch.sqrt(one * want_scale * 0.95) * 2 - 1, -1, 1)
    got_h = ro.disparity_renorm(got[:, :, 3:4].contiguous(), 0, 1, scale, False, True)
    assert (got_h.cpu() - want_h).abs().max().item() <= 1e-5


def _random_cameras(n, g):
    """Plausible camera-to-world matrices (small rotations, translations) and pinhole intrinsics."""
    ang = torch.randn(n, 3, generator=g) * 0.2
    K = torch.zeros(n, 3, 3)
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 2] = -ang[:, 2], ang[:, 1], -ang[:, 0]
    R = torch.linalg.matrix_exp(K - K.transpose(1, 2))
    c2w = torch.eye(4).repeat(n, 1, 1)
    c2w[:, :3, :3] = R
    c2w[:, :3, 3] = torch.randn(n, 3, generator=g) * 0.5
    intr = torch.zeros(n, 4, 4)
    intr[:, 0, 0] = intr[:, 1, 1] = 200 + 30 * torch.rand(n, generator=g)
    intr[:, 0, 2], intr[:, 1, 2] = 128.0, 96.0
    intr[:, 2, 2] = intr[:, 3, 3] = 1.0
    return c2w.unsqueeze(0), intr.unsqueeze(0)


def test_camera_raymap_and_pose_vs_oracle(gpu_rollout):
    """pipeline.py:29-75 and :77-163 on the same cameras / ray maps: fp32 on both sides."""
    _, _, ro = gpu_rollout
    g = torch.Generator().manual_seed(7)
    c2w, intr = _random_cameras(5, g)
    want = rollout_ref.camera_raymap(intr, c2w, (192, 256), 8).permute(0, 2, 1, 3, 4)          # [1,6,5,24,32]
    got = ro.camera_raymap(intr.cuda(), c2w.cuda(), 192, 256, normalise=False)
    assert got.shape == want.shape
    assert rel_max(got, want) <= 1e-5
    mean = torch.tensor(rollout_ref.RAYMAP_MEAN).view(1, 6, 1, 1, 1)
    std = torch.tensor(rollout_ref.RAYMAP_STD).view(1, 6, 1, 1, 1)
    got_n = ro.camera_raymap(intr.cuda(), c2w.cuda(), 192, 256, normalise=True)
    assert rel_max(got_n, (want - mean) / std) <= 1e-5

    # ray map -> pose: a camera-made ray map (well conditioned) and latent noise (what a random denoiser emits)
    for name, ray_n in (("cameras", (want - mean) / std), ("noise", torch.randn(1, 6, 8, 24, 32, generator=g))):
        lat = torch.cat([torch.randn(1, 32, ray_n.shape[2], 24, 32, generator=g), ray_n], dim=1)
        raymap = lat[:, -6:] * std + mean
        p3, p2 = rollout_ref.raymap_to_pose(raymap[:, :, 1:].clone())
        g3, g2 = ro.raymap_to_pose(lat.cuda())
        torch.cuda.synchronize()
        e3, e2 = rel_max(g3, p3), rel_max(g2, p2)
        print(f"raymap_to_pose[{name}]: trans3d {e3:.2e} trans2d {e2:.2e}")
        assert g3.shape == p3.shape and g2.shape == p2.shape
        assert e3 <= 1e-4 and e2 <= 1e-4


def test_prompt_cache_encodes_each_text_once(gpu_rollout):
    """Row f4: text mode re-encodes the same prompt every unit in the reference (pipeline.py:602-603)."""
    from deepv_b200.rollout import PromptCache
    case, _, _ = gpu_rollout
    table = rc.text_embeds(case)
    calls = []

    def text_encoder(prompt, device):
        calls.append(prompt)
        e = table["w"]
        return e["prompt_embeds"], e["prompt_attention_mask"], e["pooled_prompt_embeds"]

    pc = PromptCache({"empty": table["empty"]}, text_encoder, "cuda")
    for _ in range(8):
        enc, mask, pooled = pc.branches("a corridor, moving forward", 3, use_table=False)
    assert calls == ["a corridor, moving forward"] and pc.encoder_calls == 1
    assert enc.shape == (3, 77, 4096) and mask.shape == (3, 77) and pooled.shape == (3, 2048) and enc.is_cuda
    assert torch.equal(enc[0].cpu(), table["empty"]["prompt_embeds"][0]) and torch.equal(enc[1], enc[2])


# ---------------------------------------------------------------------------------------------
# whole iterations
def test_generate_i2v_first_iteration_vs_oracle(gpu_rollout, oracle_run):
    """pipeline.py:526-700 from one input frame: 8 units, both decodes, poses."""
    case, _, ro = gpu_rollout
    _, trace, tape = oracle_run
    t = trace[0]
    replay = rc.ReplayTape(tape.draws, tape.calls, 0)
    frames = ro.frames_from_uint8(t["frames"])
    image, disparity, t3, t2, lat = ro.generate_i2v(t["motion_prompt"], True, frames, None, None, None, temp=8,
                                                    num_inference_steps=1, noise=replay, return_latents=True)
    torch.cuda.synchronize()
    assert image.shape == t["images"].shape == (1, 3, 57, case["height"], case["width"])
    assert torch.isfinite(image).all() and torch.isfinite(disparity).all()
    for u in range(lat.shape[2]):
        print(f"unit {u + 1}: latent max|a-b|/max|ref| = {rel_max(lat[:, :, u], t['latents'][:, :, u]):.3e}")
    p_img, p_disp = psnr(image, t["images"]), psnr(disparity, t["disparity"])
    print(f"iteration 0: PSNR image {p_img:.1f} dB, disparity {p_disp:.1f} dB; "
          f"trans3d {rel_max(t3, t['trans3d']):.2e}, trans2d {rel_max(t2, t['trans2d']):.2e}")
    assert p_img >= ROLLOUT_PSNR_FLOOR_DB and p_disp >= ROLLOUT_PSNR_FLOOR_DB
    assert rel_max(lat, t["latents"]) <= ROLLOUT_LATENT_TOL
    assert t3.shape == t["trans3d"].shape and t2.shape == t["trans2d"].shape


def test_generate_i2v_second_iteration_teacher_forced(gpu_rollout, oracle_run):
    """The continuation call (25 input frames + disparity + ray map + history, 3 CFG branches, units 4..7)
    with the oracle's own inputs, so the comparison is of this iteration alone."""
    case, _, ro = gpu_rollout
    _, trace, tape = oracle_run
    t = trace[1]
    start = next(i for i, c in enumerate(tape.calls) if c == ("randn", (1, 38, 8, case["height"] // 8, case["width"] // 8)))
    replay = rc.ReplayTape(tape.draws, tape.calls, start)
    args = (t["motion_prompt"], True, ro.frames_from_uint8(t["frames"]), t["input_disparity"].cuda(),
            t["input_raymap"].cuda(), t["input_history"].cuda())
    image_k, disparity_k, *_ = ro.generate_i2v(*args, temp=8, num_inference_steps=1, noise=replay, return_latents=True)
    # the rollout's default decodes only the 32 frames a continuation iteration keeps; the full decode (what the
    # reference computes before it drops 25 frames) must agree with it bit for bit there
    ro.trim_continuation_decode = False
    try:
        image, disparity, t3, t2, lat = ro.generate_i2v(*args, temp=8, num_inference_steps=1,
                                                        noise=rc.ReplayTape(tape.draws, tape.calls, start), return_latents=True)
    finally:
        ro.trim_continuation_decode = True
    torch.cuda.synchronize()
    assert image.shape == t["images"].shape
    assert torch.equal(image_k[:, :, 25:], image[:, :, 25:]) and torch.equal(disparity_k[:, :, 25:], disparity[:, :, 25:])
    assert not image_k[:, :, :25].any()
    e_in = rel_max(lat[:, :, :4], t["latents"][:, :, :4])
    print(f"iteration 1 (teacher-forced): input latents {e_in:.2e}, generated latents "
          f"{rel_max(lat[:, :, 4:], t['latents'][:, :, 4:]):.2e}")
    p_img, p_disp = psnr(image, t["images"]), psnr(disparity, t["disparity"])
    print(f"iteration 1 (teacher-forced): PSNR image {p_img:.1f} dB, disparity {p_disp:.1f} dB")
    assert e_in <= 2e-2                      # VAE encode with bf16 activations
    assert rel_max(lat[:, :, 4:], t["latents"][:, :, 4:]) <= ROLLOUT_LATENT_TOL
    assert p_img >= ROLLOUT_PSNR_FLOOR_DB and p_disp >= ROLLOUT_PSNR_FLOOR_DB


def test_generate_two_iterations(gpu_rollout, oracle_run):
    """`generate` end to end on the device: same draws in the same order as the reference, the reference's
    result layout, and each feedback step equal to the oracle's applied to the GPU's own iteration 0."""
    case, m, ro = gpu_rollout
    res_o, trace_o, tape_o = oracle_run
    tape = rc.NoiseTape(case["seed"] + 3)
    trace = []
    res = ro.generate(dict(img=rc.first_frame(case), prompt=rc.prompts(case), prompt_type="action"), noise=tape,
                      trace=trace)
    torch.cuda.synchronize()
    assert tape.calls == tape_o.calls
    assert res["motion_prompt_list"] == res_o["motion_prompt_list"]
    for k in ("pred_img", "pred_disparity", "trans3d", "trans2d"):
        assert res[k].shape == res_o[k].shape, k
        assert torch.isfinite(res[k]).all(), k
    p = psnr(res["pred_img"][:, :, :57], res_o["pred_img"][:, :, :57])
    print(f"generate: iteration-0 frames PSNR {p:.1f} dB")
    assert p >= ROLLOUT_PSNR_FLOOR_DB

    # feedback: oracle arithmetic on the GPU's iteration-0 output vs what the GPU fed its iteration 1
    t0, t1 = trace
    fb = rollout_ref.Feedback(m)
    img, disp = fb.absorb(0, t0["images"].float().cpu(), t0["disparity"].float().cpu(), t0["trans3d"].cpu(),
                          t0["trans2d"].cpu(), t0["motion_prompt"])
    want_frames, want_disp, want_ray, _ = fb.next_inputs(img, disp, rc.NoiseTape(0))
    assert torch.equal(t1["frames"].cpu(), rollout_ref.frames_to_input(want_frames))
    assert (t1["input_disparity"].cpu() - want_disp).abs().max().item() <= 1e-5
    mean = torch.tensor(rollout_ref.RAYMAP_MEAN).view(1, 6, 1, 1, 1)
    std = torch.tensor(rollout_ref.RAYMAP_STD).view(1, 6, 1, 1, 1)
    want_ray = (want_ray - mean) / std
    got_ray = t1["input_raymap"].cpu()
    assert rel_max(got_ray[:, :3], want_ray[:, :3]) <= 1e-4
    assert rel_max(got_ray[:, 3:, 1:], want_ray[:, 3:, 1:]) <= 1e-3
    # input frame 0's origin is sqrt(round-off) in the reference too (tests/test_oracle_rollout.py)
    assert (got_ray[:, 3:, :1] - want_ray[:, 3:, :1]).abs().max().item() <= 2e-2
    assert t1["input_history"].shape == (1, 38, 1, case["height"] // 8, case["width"] // 8)
