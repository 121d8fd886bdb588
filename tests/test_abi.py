"""CPU: the C-ABI shared library builds, loads, and exports every symbol the header declares."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    src = (ROOT / "include" / "deepv_b200.h").read_text()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dv_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    from deepv_b200 import build, _lib
    path = build.build()
    assert path.exists()
    lib = ctypes.CDLL(str(path))
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/deepv_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_load_and_error_channel():
    from deepv_b200 import _lib
    lib = _lib.load()
    assert lib.dv_version() >= 100
    # invalid call: must fail loudly with a message, never silently fall back
    rc = lib.dv_cfg_euler_step(None, 2, None, None, 0, 0.0, 0.0, 0.0, 0.0, 1, None)
    assert rc != 0
    assert b"null" in lib.dv_last_error()


def test_no_cpu_fallback():
    import pytest
    import torch
    from deepv_b200 import _lib
    from deepv_b200.scheduler import B200Scheduler
    s = B200Scheduler()
    s.set_timesteps(5, 0)
    with pytest.raises(_lib.DeepVError):
        s.step(model_output=torch.zeros(4), timestep=s.timesteps[0], sample=torch.zeros(4))


def test_product_does_not_import_oracle():
    for p in (ROOT / "deepv_b200").rglob("*.py"):
        txt = p.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt, p
