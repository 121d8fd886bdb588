"""CPU: the algorithmic-work model (deepv_b200/work.py) against the traced call shapes of SURVEY.md App. B and the
worked FLOP values of SURVEY.md §8d / App. C (measured on the real reference during the survey)."""
import pytest

from deepv_b200 import work as w

APP_B = {1: (96, 384, 1536), 2: (96, 384, 1536), 3: (144, 432, 1728), 4: (192, 480, 1776), 5: (240, 528, 1824),
         6: (288, 576, 1872), 7: (336, 624, 1920), 8: (384, 672, 1968)}


def test_first_iteration_layouts_match_traced_shapes():
    fw = w.rollout_forwards(1)
    assert len(fw) == 24 and sum(f["count"] for f in fw) == 120
    for f in fw:
        lv, lc = w.mmdit_tokens(f["clips"], f["hist"])
        assert (lv, lc, f["B"]) == (APP_B[f["unit"]][f["stage"]], 77, 2), f
    u4s2 = next(f for f in fw if f["unit"] == 4 and f["stage"] == 2)
    assert u4s2["clips"] == [(1, 12, 16), (1, 24, 32), (1, 48, 64), (1, 48, 64)]
    u8s0 = next(f for f in fw if f["unit"] == 8 and f["stage"] == 0)
    assert u8s0["clips"] == [(6, 12, 16), (1, 12, 16), (1, 12, 16)]


def test_steady_iteration_layouts():
    fw = [f for f in w.rollout_forwards(2) if f["iteration"] == 1]
    assert [f["unit"] for f in fw[::3]] == [4, 5, 6, 7] and sum(f["count"] for f in fw) == 60
    for f in fw:   # "units 4...7 have exactly the unit-5...8 layouts" with B=3 and 192 history tokens
        lv, lc = w.mmdit_tokens(f["clips"], f["hist"])
        assert (lv, lc, f["B"]) == (APP_B[f["unit"] + 1][f["stage"]], 269, 3), f


def test_flops_match_survey_worked_values():
    big = w.mmdit_flops(3, [(5, 12, 16), (1, 24, 32), (1, 48, 64), (1, 48, 64)], True)
    assert big["linear"] / 1e12 == pytest.approx(9.09, abs=0.01)
    assert big["attention"] / 1e12 == pytest.approx(1.40, abs=0.005)
    dense = w.mmdit_flops(3, [(5, 12, 16), (1, 24, 32), (1, 48, 64), (1, 48, 64)], True, masked=False)
    assert dense["attention"] / 1e12 == pytest.approx(2.21, abs=0.005)
    assert w.mmdit_flops(2, [(1, 12, 16), (1, 12, 16)], False)["total"] / 1e12 == pytest.approx(0.47, abs=0.01)
    assert w.rollout_work(1)["mmdit"] / 1e12 == pytest.approx(328, abs=0.5)
    assert (w.rollout_work(2)["mmdit"] - w.rollout_work(1)["mmdit"]) / 1e12 == pytest.approx(326, abs=0.5)
    assert w.vae_decode_flops(8) / 1e12 == pytest.approx(322.7, abs=0.1)             # tiled as the reference does
    assert w.vae_decode_tile_flops(8, 48, 64) / 1e12 == pytest.approx(221.3, abs=0.2)  # untiled
    assert w.rollout_work(2)["frames"] == 89


def test_tile_grid_is_the_reference_geometry():
    assert w.tile_grid(48, 64) == [(32, 32), (32, 32), (32, 16), (24, 32), (24, 32), (24, 16)]


def test_trimmed_decode_schedule_and_executed_flops():
    """The frame schedule of the trimmed decode (csrc/vae.cu:run_tile, mirrored by work.decode_frame_schedule) and the work it
    saves, pinned to what the B200 profiler counted: a rollout executed 2070.5 TFLOP before the trimmed decode
    (profiles/r02s_bench_rollout.json, classes' achieved x ms) and 1974.7 TFLOP with it (profiles/r02v_bench_rollout.json)."""
    res_t, sp_t0, tp_t0, taps_t0 = w.decode_frame_schedule(25)
    assert taps_t0 == 23                                   # conv_out reads two frames back
    assert res_t[3] == [15, 19, 23]                        # 128-channel level (57 frames): conv2 of resnet j starts here,
    assert tp_t0[2] == 6 and sp_t0[2] == 4                 # conv1 two frames earlier; upsamplers in front of it (29 frames)
    assert res_t[2] == [0, 0, 2] and res_t[1] == [0, 0, 0] and res_t[0] == [0, 0, 0]
    assert w.decode_frame_schedule(0) == ([[0] * 3 for _ in range(4)], [0] * 4, [0] * 4, 0)
    assert w.vae_decode_flops(8, first_frame=0) == w.vae_decode_flops(8)
    full, part = w.vae_decode_flops(8), w.vae_decode_flops(8, first_frame=25)
    assert 0.84 < part / full < 0.86                       # a trimmed decode is ~15 % cheaper
    r = w.rollout_work(2)
    assert r["executed"] < r["total"]
    assert (r["total"] - r["executed"]) / 1e12 == pytest.approx(2070.5 - 1974.7, abs=1.0)
    assert r["executed"] / 1e12 == pytest.approx(1974.7, rel=5e-3)
    assert w.rollout_work(1)["executed"] == w.rollout_work(1)["total"]     # the first iteration keeps all 57 frames
