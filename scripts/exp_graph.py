"""Bring-up experiment: is a chain of small GEMMs GPU-bound or launch-bound?  Same launches eager vs
replayed from a CUDA graph, with and without an elementwise kernel (different smem carve-out) between."""
import ctypes as C
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from deepv_b200 import _lib  # noqa: E402

lib = _lib.load()
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None


def bench(name, body, n_inner):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        sp = C.c_void_p(s.cuda_stream)
        for _ in range(3):
            body(sp)
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(s)
        for _ in range(5):
            body(sp)
        e1.record(s)
        t_cpu = (time.perf_counter() - t0) / 5
        s.synchronize()
        eager = e0.elapsed_time(e1) / 5
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            body(C.c_void_p(torch.cuda.current_stream().cuda_stream))
        for _ in range(3):
            g.replay()
        s.synchronize()
        e0.record(s)
        for _ in range(5):
            g.replay()
        e1.record(s)
        s.synchronize()
        graph = e0.elapsed_time(e1) / 5
    print(f"{name:50s} eager {eager * 1e3 / n_inner:7.1f} us/launch (cpu issue {t_cpu * 1e6 / n_inner:6.1f} us)   graph {graph * 1e3 / n_inner:7.1f} us/launch", flush=True)


for (B, M, N, K) in [(2, 77, 4608, 1536), (2, 77, 1536, 6144), (2, 77, 1536, 1536), (2, 384, 6144, 1536), (2, 1536, 6144, 1536)]:
    A = (torch.randn(B, M, K, device="cuda") * 0.5).bfloat16()
    Ws = [(torch.randn(N, K, device="cuda") * 0.05).bfloat16() for _ in range(10)]
    out = torch.empty(B, M, N, device="cuda", dtype=torch.bfloat16)
    x = torch.randn(B, M, K, device="cuda")

    def gemms(sp):
        for i in range(10):
            _lib.check(lib.dv_gemm_bf16(p(A), p(Ws[i]), None, p(out), B, M, N, K, 0, sp))

    def mixed(sp):
        for i in range(10):
            _lib.check(lib.dv_gemm_bf16(p(A), p(Ws[i]), None, p(out), B, M, N, K, 0, sp))
            x.mul_(1.0001)

    bench(f"10x gemm B{B} M{M} N{N} K{K}", gemms, 10)
    bench(f"10x (gemm + elementwise) B{B} M{M} N{N} K{K}", mixed, 20)
