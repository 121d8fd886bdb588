"""Bring-up experiment: time dv_gemm_bf16 / dv_conv3d_cl at a few shapes (CUDA events, L2-flushing
rotation of buffers).  Env knobs DV_GEMM_BN / DV_GEMM_DBG are read by the library."""
import ctypes as C
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from deepv_b200 import _lib  # noqa: E402

lib = _lib.load()


def p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


tag = f"BN={os.environ.get('DV_GEMM_BN', '-')} DBG={os.environ.get('DV_GEMM_DBG', '0')}"
for (B, M, N, K) in [(1, 8192, 8192, 8192), (1, 4096, 4096, 4096), (2, 1536, 6144, 1536), (2, 1536, 1536, 6144),
                     (2, 77, 4608, 1536), (2, 77, 6144, 1536)]:
    A = (torch.randn(B, M, K, device="cuda") * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    out = torch.empty(B, M, N, device="cuda", dtype=torch.bfloat16)
    ms = timeit(lambda: _lib.check(lib.dv_gemm_bf16(p(A), p(W), None, p(out), B, M, N, K, 0, None)))
    print(f"{tag} gemm B{B} M{M} N{N} K{K}: {ms * 1e3:9.1f} us  {2.0 * B * M * N * K / ms / 1e9:8.1f} TFLOP/s", flush=True)
for (T, H, Wd, Ci, Co) in [(9, 256, 256, 128, 128), (5, 128, 128, 256, 256), (3, 64, 64, 512, 512)]:
    x = (torch.randn(1, T, H, Wd, Ci, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(Co, 27, Ci, device="cuda") * 0.05).bfloat16()
    out = torch.empty(1, T, H, Wd, Co, device="cuda", dtype=torch.bfloat16)
    ms = timeit(lambda: _lib.check(lib.dv_conv3d_cl(p(x), p(w), None, None, p(out), 1, T, H, Wd, Ci, Co, Co, 3, 0, 0, None)), n=10)
    fl = 2.0 * T * H * Wd * Co * 27 * Ci
    print(f"{tag} conv T{T} H{H} W{Wd} Ci{Ci} Co{Co}: {ms * 1e3:9.1f} us  {fl / ms / 1e9:8.1f} TFLOP/s", flush=True)
