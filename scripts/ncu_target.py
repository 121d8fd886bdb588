"""Short ncu target: one launch of each hot kernel at its benchmark shape.

    ncu --set full --clock-control none --import-source on -k regex:"gemm_|attn_" -o prof python scripts/ncu_target.py
"""
import ctypes as C
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from deepv_b200 import _lib  # noqa: E402

lib = _lib.load()
which = sys.argv[1] if len(sys.argv) > 1 else "all"
st = None


def p(t):
    return t.data_ptr()


torch.manual_seed(0)
if which in ("all", "conv"):
    # up3 resnet conv of the VAE decoder: 128 -> 128 channels at 256x256, 9 frames (swapped-operand path)
    B, T, H, W, Ci, Co = 1, 9, 256, 256, 128, 128
    x = (torch.randn(B, T, H, W, Ci, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(Co, 27, Ci, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(Co, device="cuda")
    out = torch.empty(B, T, H, W, Co, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.dv_conv3d_cl(p(x), p(w), p(bias), None, p(out), B, T, H, W, Ci, Co, Co, 3, 0, 0, st))
    # up2 resnet conv: 256 -> 256 at 128x128, 5 frames
    B, T, H, W, Ci, Co = 1, 5, 128, 128, 256, 256
    x = (torch.randn(B, T, H, W, Ci, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(Co, 27, Ci, device="cuda") * 0.05).bfloat16()
    out = torch.empty(B, T, H, W, Co, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.dv_conv3d_cl(p(x), p(w), None, None, p(out), B, T, H, W, Ci, Co, Co, 3, 0, 0, st))
if which in ("all", "gemm"):
    # FF1 of a stage-2 forward: [2][1536, 1536] x [6144, 1536]^T, GELU epilogue
    Bt, M, N, K = 2, 1536, 6144, 1536
    A = (torch.randn(Bt, M, K, device="cuda") * 0.5).bfloat16()
    Wt = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda")
    Cc = torch.empty(Bt, M, N, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.dv_gemm_bf16(p(A), p(Wt), p(bias), p(Cc), Bt, M, N, K, 1, st))
    # a large square problem
    M = N = K = 4096
    A = (torch.randn(1, M, K, device="cuda") * 0.5).bfloat16()
    Wt = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    Cc = torch.empty(1, M, N, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.dv_gemm_bf16(p(A), p(Wt), None, p(Cc), 1, M, N, K, 0, st))
    # FF2 of the context stream (small M, split-K cluster): [2][77, 6144] x [1536, 6144]^T
    A = (torch.randn(2, 77, 6144, device="cuda") * 0.5).bfloat16()
    Wt = (torch.randn(1536, 6144, device="cuda") * 0.05).bfloat16()
    Cc = torch.empty(2, 77, 1536, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.dv_gemm_bf16(p(A), p(Wt), None, p(Cc), 2, 77, 1536, 6144, 0, st))
if which in ("all", "small"):
    # stage-0 shapes (fewer tiles than SMs: cluster split-K): FF2 / out-proj / QKV-sized problems at M = 96 rows x B 2
    for (Bt, M, N, K) in ((2, 96, 1536, 6144), (2, 96, 1536, 1536), (2, 96, 4608, 1536), (2, 384, 6144, 1536)):
        A = (torch.randn(Bt, M, K, device="cuda") * 0.5).bfloat16()
        Wt = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        Cc = torch.empty(Bt, M, N, device="cuda", dtype=torch.bfloat16)
        _lib.check(lib.dv_gemm_bf16(p(A), p(Wt), None, p(Cc), Bt, M, N, K, 0, st))
if which in ("all", "conv57"):
    # the dominant launch of a rollout: up3 resnet conv 128 -> 128 at 256x256 over the 57 frames of an iteration
    B, T, H, W, Ci, Co = 1, 57, 256, 256, 128, 128
    x = (torch.randn(B, T, H, W, Ci, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(Co, 27, Ci, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(Co, device="cuda")
    out = torch.empty(B, T, H, W, Co, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.dv_conv3d_cl(p(x), p(w), p(bias), None, p(out), B, T, H, W, Ci, Co, Co, 3, 0, 0, st))
if which in ("all", "attn"):
    B, L, H = 2, 1613, 24
    qkv = torch.randn(B, L, 3 * H * 64, device="cuda").bfloat16()
    sizes = [77, 768, 768]
    bounds, acc = [], 0
    for s in sizes:
        acc += s
        bounds.append(acc)
    kv = torch.empty(L, dtype=torch.int32)
    pos = 0
    for i, s in enumerate(sizes):
        kv[pos:pos + s] = bounds[max(i, 1)]
        pos += s
    Lpad = (L + 127) // 128 * 128
    kb = torch.zeros(B, Lpad, device="cuda")
    kb[:, L:] = float("-inf")
    kb[0, 1:77] = float("-inf")
    kb[1, 12:77] = float("-inf")
    out = torch.empty(B, L, H * 64, device="cuda", dtype=torch.bfloat16)
    kvd = kv.cuda()
    _lib.check(lib.dv_attention(p(qkv), p(out), p(kvd), p(kb), B, L, Lpad, H, st))
torch.cuda.synchronize()
print("ncu target done")
