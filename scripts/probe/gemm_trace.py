"""Probe (GPU): where a dense GEMM launch spends its time — per-CTA globaltimer stamps of gemm_tc_kernel / gemm_pair_kernel.

Builds a copy of the library with -DDV_GEMM_TRACE (gemm.cu only), launches the video + context pair of a joint-block GEMM
through dv_gemm_bf16 (one problem per call) at shapes of the rollout and prints the phase durations over all CTAs.
    python scripts/probe/gemm_trace.py
"""
import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
here = Path(__file__).resolve().parent
so = here / "libdeepv_gemm_trace.so"
csrc = ROOT / "deepv_b200" / "csrc"
flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]
if not so.exists() or "--rebuild" in sys.argv:
    from deepv_b200 import build
    build.build(force=not list((ROOT / "deepv_b200" / "build").glob("*.o")))   # (the object files do not travel with gpurun)
    obj = here / "gemm_trace.o"
    subprocess.check_call(["nvcc", *flags, "-DDV_GEMM_TRACE", "-c", str(csrc / "gemm.cu"), "-o", str(obj)])
    others = [str(o) for o in (ROOT / "deepv_b200" / "build").glob("*.o") if o.name != "gemm.o"]
    subprocess.check_call(["nvcc", "-shared", "-o", str(so), str(obj), *others, "-gencode", "arch=compute_100a,code=sm_100a",
                           "-cudart", "static"])
if "--build-only" in sys.argv:
    sys.exit(0)

lib = C.CDLL(str(so))
lib.dv_gemm_trace_buffer.restype = C.c_void_p
lib.dv_gemm_bf16.restype = C.c_int
lib.dv_gemm_bf16.argtypes = [C.c_void_p] * 4 + [C.c_int] * 5 + [C.c_void_p]
cudart = C.CDLL("libcudart.so.12")
buf = lib.dv_gemm_trace_buffer()
names = ["entry->setup", "setup->pdl", "pdl->first operands", "main loop (MMA issue)", "issue->accumulator done",
         "accumulator->all splits parked", "reduce / epilogue", "epilogue->exit"]
torch.manual_seed(0)
shapes = [(2, 384, 1536, 1536, 0), (2, 384, 1536, 6144, 0), (2, 384, 6144, 1536, 1), (3, 269, 1536, 6144, 0),
          (2, 1536, 6144, 1536, 1), (3, 2189, 6144, 1536, 1), (3, 1920, 1536, 6144, 0)]
for (Bt, M, N, K, epi) in shapes:
    A = (torch.randn(Bt, M, K, device="cuda") * 0.5).bfloat16()
    Wt = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda")
    Cc = torch.empty(Bt, M, N, device="cuda", dtype=torch.bfloat16)
    args = (A.data_ptr(), Wt.data_ptr(), bias.data_ptr(), Cc.data_ptr(), Bt, M, N, K, epi, None)
    for _ in range(3):
        assert lib.dv_gemm_bf16(*args) == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        lib.dv_gemm_bf16(*args)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / 20
    cudart.cudaMemset(C.c_void_p(buf), 0, C.c_size_t(4096 * 12 * 8))
    lib.dv_gemm_bf16(*args)
    torch.cuda.synchronize()
    tr = np.zeros(4096 * 12, dtype=np.int64)
    assert cudart.cudaMemcpy(tr.ctypes.data_as(C.c_void_p), C.c_void_p(buf), C.c_size_t(tr.nbytes), 2) == 0
    tr = tr.reshape(4096, 12)
    tr = tr[tr[:, 1] > 0]
    t0 = tr[:, 1].min()
    span = tr[:, 9].max() - t0
    fl = 2.0 * Bt * M * N * K
    print(f"== B{Bt} M{M} N{N} K{K} epi{epi}: {us:.1f} us back-to-back ({fl / us / 1e6:.0f} TFLOP/s), {len(tr)} CTAs, "
          f"first entry -> last exit {span / 1000:.1f} us")
    st = tr[:, 1:10].astype(np.float64)
    st[st == 0] = np.nan
    # forward-fill stamps a CTA did not record (no split-K: slot 7; idle CTAs)
    for c in range(1, st.shape[1]):
        m = np.isnan(st[:, c])
        st[m, c] = st[m, c - 1]
    d = np.diff(st, axis=1)
    for i, nm in enumerate(names):
        print(f"   {nm:34s} mean {np.nanmean(d[:, i]) / 1000:7.2f}  median {np.nanmedian(d[:, i]) / 1000:7.2f}  max {np.nanmax(d[:, i]) / 1000:7.2f} us")
    print(f"   entry spread {(tr[:, 1].max() - t0) / 1000:.2f} us; whole CTA mean {np.nanmean(st[:, 8] - st[:, 0]) / 1000:.2f} us")
