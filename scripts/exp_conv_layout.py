"""Bring-up experiment: does the 5-D TMA box of the implicit-GEMM conv run faster when the 16 pixels
of a box row are contiguous (Cin = 64: pixel stride 128 B) than when they are Cin*2 bytes apart?"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from deepv_b200 import _lib  # noqa: E402

lib = _lib.load()
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
for (T, H, Wd, Ci, Co) in [(5, 128, 128, 64, 256), (5, 128, 128, 128, 256), (5, 128, 128, 256, 256), (5, 128, 128, 512, 256),
                           (9, 256, 256, 64, 128), (9, 256, 256, 128, 128)]:
    x = (torch.randn(1, T, H, Wd, Ci, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(Co, 27, Ci, device="cuda") * 0.05).bfloat16()
    out = torch.empty(1, T, H, Wd, Co, device="cuda", dtype=torch.bfloat16)
    run = lambda: _lib.check(lib.dv_conv3d_cl(p(x), p(w), None, None, p(out), 1, T, H, Wd, Ci, Co, Co, 3, 0, 0, None))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * T * H * Wd * Co * 27 * Ci
    kb = 27 * Ci // 64
    print(f"conv T{T} H{H} W{Wd} Ci{Ci} Co{Co}: {ms * 1e3:8.1f} us {fl / ms / 1e9:8.1f} TFLOP/s  ({kb} k-blocks/tile)", flush=True)
