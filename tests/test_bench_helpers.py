"""bench.py's post-processing (no GPU): the per-class roofline from a per-launch profile of the final build
(profiles/r02v_prof.csv.gz, written by `bench.py --profile-dump` on a B200) and the whole-step summary."""
import gzip
import importlib.util
import json
import shutil
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("dv_bench", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)       # (main() only runs under __main__)
    return mod


def test_roofline_and_whole_step_from_a_recorded_profile(bench, tmp_path):
    src = ROOT / "profiles" / "r02v_prof.csv.gz"
    if not src.exists():
        pytest.skip("recorded profile not in the tree")
    csv_path = tmp_path / "prof.csv"
    with gzip.open(src, "rb") as fi, open(csv_path, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    peaks = {"bf16_tflops_sustained": 1384.6, "bf16_tflops": 1644.1, "hbm_gbs": 6464.9}
    r = bench.roofline_from_profile(str(csv_path), 2335.0, peaks)
    assert set(r["classes"]) == {"conv", "dense_large_M", "dense_small_M", "attention", "hbm_elementwise"}
    assert r["kernel_class"] == "conv" and r["kernel"].startswith("conv T57 H256 W256 Ci128 N128")
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and 0.9 < r["frac"] < 1.2
    assert r["traffic"] is not None and r["traffic"] > r["algorithmic_bytes_per_launch"]      # ncu DRAM bytes >= algorithmic
    assert r["executed_tflop"] == pytest.approx(1974.7, abs=1.0)
    assert r["classes"]["dense_small_M"]["bound"] == "hbm" and r["classes"]["attention"]["frac"] < 0.5
    ws = bench.whole_step_summary(r, 2087.1, 1, peaks, True, 24)
    assert ws["algorithmic_tflop"] == pytest.approx(2063.4, abs=0.1)
    assert ws["executed_tflop_work_model"] == pytest.approx(1967.7, abs=0.1)
    assert ws["frac_of_sustained_peak"] == pytest.approx(0.714, abs=2e-3)
    assert ws["executed_frac_of_sustained_peak"] == pytest.approx(0.683, abs=2e-3)
    json.dumps({"roofline": r, "whole_step": ws})                                          # the line must serialise
    unit = bench.whole_step_summary(r, 153.6, 1, peaks, False, 24)
    assert unit["executed_tflop_work_model"] == unit["algorithmic_tflop"]


def test_classify_is_stable_under_launch_tag_suffixes(bench):
    assert bench.classify(0, "gemm B2 M384+77 N6144 K1536 s1 2cta e1 flat") == bench.classify(0, "gemm B2 M384+77 N6144 K1536 s1 2cta e1")
    assert bench.classify(0, "gemm B3 M1920+269 N6144 K1536 s1 2cta e1 flat") == "dense_large_M"
    assert bench.classify(0, "gemm B3 M269 N1536 K6144 s2 e2 flat") == "dense_small_M"
    assert bench.classify(2, "attn B3 L2237 H24 k64") == "attention" and bench.classify(1, "conv T44 H256 W256 Ci256 N128 k3 halo e0") == "conv"
