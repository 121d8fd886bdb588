"""Driver of scripts/probe/desc_shift.cu (build: see the nvcc line below; run on a B200)."""
import ctypes as C
import sys
from pathlib import Path

import torch

here = Path(__file__).resolve().parent
C.CDLL(str(here.parents[1] / "deepv_b200" / "libdeepv_b200.so"), mode=C.RTLD_GLOBAL)
lib = C.CDLL(str(here / "desc_shift.so"))
g = torch.Generator().manual_seed(1)
A = torch.randn(128, 64, generator=g).bfloat16().cuda()
B = torch.randn(320, 64, generator=g).bfloat16().cuda()
out = torch.empty(128, 256, device="cuda")
for mode in (0, 1):
    for shift in (0, 1, 2, 3, 8, 9):
        if mode == 0:
            rows = torch.arange(256) + shift
        else:
            rows = (torch.arange(32).view(32, 1) * 10 + torch.arange(8).view(1, 8)).reshape(-1) + shift
        want = A.float() @ B.float()[rows.cuda()].T
        res = []
        for base_off in (0, 1):
            out.zero_()
            rc = lib.probe_desc_shift(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(out.data_ptr()),
                                      shift, mode, base_off)
            err = (out - want).abs().max().item() / want.abs().max().item()
            res.append(f"base_off={base_off}: rc={rc} err={err:.2e}")
        print(f"mode {mode} shift {shift}: " + "  ".join(res), flush=True)
