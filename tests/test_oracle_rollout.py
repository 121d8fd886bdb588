"""The rollout restatement (oracle/rollout_ref.py) against the golden rollout of the REAL reference
(`InferencePipeline.generate`, two iterations; tests/golden/make_rollout_golden.py), on the same
seeded noise tape.  CPU only, fp32 on both sides.

Iteration 0 agrees to float round-off (<= 1e-5).  The feedback into iteration 1 contains one
ill-conditioned step of the reference itself: the relative pose of the first input unit is
inv(P) @ P (pipeline.py:354-356), i.e. a translation of ~1e-7 of pure round-off, which
sign(x)*sqrt|x| (pipeline.py:361) turns into +-6e-4 before the division by the ray-map std.  That
entry (ray origin of input frame 0) is therefore compared with an absolute tolerance, and everything
downstream of it (iteration 1) with 1e-3 instead of 1e-5."""
from pathlib import Path

import pytest
import torch

from oracle import rollout_ref
from tests.golden import rollout_cases as rc

G = Path(__file__).resolve().parent / "golden"


build_models = rc.build_oracle_models


def close(a, b, tol, what):
    err = (a.float() - b.float()).abs().max().item() / max(b.float().abs().max().item(), 1e-12)
    print(f"{what}: max|a-b|/max|ref| = {err:.2e}")
    assert a.shape == b.shape, what
    assert err <= tol, what


def digest_close(v, gold, tol, what):
    d = rc.digest_frames(v)
    assert d["shape"] == gold["shape"], what
    close(d["sub"], gold["sub"], tol, what + ".sub")
    close(d["frame_mean"], gold["frame_mean"], tol, what + ".frame_mean")
    close(d["frame_std"], gold["frame_std"], tol, what + ".frame_std")


@pytest.fixture(scope="module")
def rollout():
    torch.set_grad_enabled(False)
    case = rc.ROLLOUT
    tape = rc.NoiseTape(case["seed"] + 3)
    trace = []
    steps = [case["model_cfg"]["num_inference_steps"]] * 3
    res = rollout_ref.generate(build_models(case), rc.first_frame(case), rc.prompts(case), tape, steps, trace)
    return res, trace, tape, torch.load(G / "rollout_golden.pt")


def test_noise_draws_in_reference_order(rollout):
    _, _, tape, gold = rollout
    assert tape.calls == gold["tape_calls"]


def test_feedback_inputs_match_reference(rollout):
    """What `generate` hands to the second `generate_i2v`: uint8 frames (exact), renormalised
    disparity, ray map of the relative poses, history latent (pipeline.py:339-414)."""
    _, trace, _, gold = rollout
    assert len(trace) == len(gold["calls"]) == 2
    for it, (t, g) in enumerate(zip(trace, gold["calls"])):
        assert list(t["motion_prompt"]) == g["motion_prompt"]
        assert t["frames"].shape[0] == g["n_images"]
        # truncation to uint8 amplifies a 1-ulp difference to one grey level; allow a handful of pixels
        diff = (t["frames"][:, ::8, ::8].int() - g["input_frames_sub"].int()).abs()
        assert diff.max() <= 1 and (diff != 0).float().mean() < 1e-3, f"iteration {it} uint8 frames"
        if it == 0:
            assert t["input_disparity"] is None and g["input_disparity"] is None
            continue
        digest_close(t["input_disparity"], g["input_disparity"], 1e-4, "input_disparity")
        a, b = t["input_raymap"], g["input_raymap"]
        close(a[:, :3], b[:, :3], 1e-5, "input_raymap.directions")
        close(a[:, 3:, 1:], b[:, 3:, 1:], 1e-4, "input_raymap.origins[1:]")
        assert (a[:, 3:, :1] - b[:, 3:, :1]).abs().max() < 5e-3        # sqrt of round-off, see the module docstring
        close(t["input_history"], g["input_history"], 1e-3, "input_history")


def test_iteration_outputs_match_reference(rollout):
    _, trace, _, gold = rollout
    for it, (t, g) in enumerate(zip(trace, gold["calls"])):
        tol = 1e-5 if it == 0 else 1e-3
        digest_close(t["images"], g["images"], tol, f"it{it}.images")
        digest_close(t["disparity"], g["disparity"], tol, f"it{it}.disparity")
        close(t["trans3d"], g["trans3d"], 3 * tol, f"it{it}.trans3d")
        close(t["trans2d"], g["trans2d"], tol, f"it{it}.trans2d")


def test_rollout_result_matches_reference(rollout):
    res, _, _, gold = rollout
    assert res["motion_prompt_list"] == gold["motion_prompt_list"]
    assert rollout[1] is not None
    digest_close(res["pred_img"], gold["pred_img"], 1e-3, "pred_img")
    digest_close(res["pred_disparity"], gold["pred_disparity"], 1e-3, "pred_disparity")
    close(res["trans3d"], gold["trans3d"], 3e-3, "trans3d")
    close(res["trans2d"], gold["trans2d"], 1e-3, "trans2d")


def test_no_need_depth_unit_matches_live_reference():
    """`model_cfg['no_need_depth']` (pipeline.py:476-478) through the real `generate_one_unit`, when the
    reference tree is present (build container); the restatement must zero the same channels."""
    from oracle import _shim, mmdit_ref, scheduler_ref
    if not _shim.reference_available():
        pytest.skip("reference tree not present")
    from tests.golden.make_rollout_golden import build_reference_pipeline
    torch.set_grad_enabled(False)
    case = rc.ROLLOUT
    tape = rc.NoiseTape(5)
    _, pipe = build_reference_pipeline(case, tape)
    pipe.model_cfg["no_need_depth"] = True
    pipe._guidance_scale, pipe._video_guidance_scale = 4.0, 3.5
    g = torch.Generator().manual_seed(9)
    lat = torch.randn(1, 38, 1, 6, 8, generator=g)
    conds = [[torch.randn(2, 38, 1, 6 * 2 ** i, 8 * 2 ** i, generator=g)] for i in range(3)]
    te = rc.text_embeds(case)
    enc = torch.cat([te["empty"]["prompt_embeds"], te["w"]["prompt_embeds"]])
    pooled = torch.cat([te["empty"]["pooled_prompt_embeds"], te["w"]["pooled_prompt_embeds"]])
    mask = torch.cat([te["empty"]["prompt_attention_mask"], te["w"]["prompt_attention_mask"]])
    want = pipe.generate_one_unit(lat.clone(), None, [[c.clone() for c in st] for st in conds], enc, mask, pooled,
                                  [1, 1, 1], 6, 8, 1, torch.device("cpu"), torch.float32)
    m = build_models(case)
    replay = rc.NoiseTape(5)

    def model_fn(clips, tt):
        return mmdit_ref.mmdit_forward(m.dit_W, m.dit_cfg, clips, tt.float(), enc, mask, pooled, pos_table=m.pos_table())

    block = [replay.block(1, 38, 1, 12, 16, m.tables["gamma"]), replay.block(1, 38, 1, 24, 32, m.tables["gamma"])]
    got = scheduler_ref.generate_one_unit(model_fn, m.tables, lat, conds, block, 2, [1, 1, 1], 3.5, 6.0,
                                          no_need_depth=True)
    for i in range(3):
        close(got[i], want[i], 1e-5, f"no_need_depth stage {i}")
    plain = scheduler_ref.generate_one_unit(model_fn, m.tables, lat, conds, block, 2, [1, 1, 1], 3.5, 6.0)
    assert (plain[2] - got[2]).abs().max() > 1e-3          # the flag does change the result
