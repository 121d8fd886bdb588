"""Probe (GPU): cost of a grid-wide barrier over 148 co-resident CTAs, per variant (see grid_barrier.cu)."""
import ctypes as C
import os
import subprocess
import sys

import torch

here = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(here, "grid_barrier.so")
if not os.path.exists(so):
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-shared", "-Xcompiler", "-fPIC",
                           "-o", so, os.path.join(here, "grid_barrier.cu")])
lib = C.CDLL(so)
lib.run_variant.restype = C.c_float
lib.run_variant.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
bar = torch.zeros(16, dtype=torch.int32, device="cuda")
sink = torch.zeros(16, device="cuda")
names = ["full (fence.proxy.async x2, threadfence, red.release + ld.acquire poll)", "no trailing proxy fence",
         "threadfence only", "atomicAdd + volatile poll, no fences", "as 2 with nanosleep(100) back-off"]
iters = 2000
for v, n in enumerate(names):
    lib.run_variant(v, 100, bar.data_ptr(), sink.data_ptr())
    ms = lib.run_variant(v, iters, bar.data_ptr(), sink.data_ptr())
    print(f"variant {v}: {ms * 1e3 / iters:7.2f} us per barrier   {n}", flush=True)
