"""GPU: the exported building-block kernels (tcgen05 GEMM, implicit-GEMM conv3d, attention)
through the C ABI against plain torch fp32 references of the same op on the same bf16 inputs."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 8e-3  # bf16 output rounding (2^-9) + fp32 accumulation-order differences


def p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-20)).item()


@pytest.mark.parametrize("B,M,N,K,epi", [
    (1, 128, 64, 64, 0),        # one tile, one k-block
    (2, 300, 384, 512, 0),      # ragged M (TMA zero fill), batch
    (1, 1000, 1536, 1536, 1),   # GELU-tanh epilogue
    (3, 2000, 4608, 1536, 0),   # BN = 256 path, many waves
    (2, 96, 1536, 6144, 0),     # small-M, long K
    (1, 1, 32, 64, 0),          # a single row
])
def test_gemm_bf16(lib, B, M, N, K, epi):
    from deepv_b200 import _lib
    torch.manual_seed(0)
    A = (torch.randn(B, M, K, device="cuda") * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda")
    out = torch.zeros(B, M, N, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.dv_gemm_bf16(p(A), p(W), p(bias), p(out), B, M, N, K, epi, None))
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t() + bias
    if epi == 1:
        ref = torch.nn.functional.gelu(ref, approximate="tanh")
    assert rel(out, ref) <= TOL


def test_gemm_rejects_bad_shapes(lib):
    A = torch.zeros(1, 8, 48, device="cuda", dtype=torch.bfloat16)
    W = torch.zeros(32, 48, device="cuda", dtype=torch.bfloat16)
    out = torch.zeros(1, 8, 32, device="cuda", dtype=torch.bfloat16)
    assert lib.dv_gemm_bf16(p(A), p(W), None, p(out), 1, 8, 32, 48, 0, None) != 0  # K % 64
    assert b"K=48" in lib.dv_last_error()


@pytest.mark.parametrize("B,T,H,W,Cin,Cout,ks,store,drop", [
    (1, 2, 8, 16, 64, 64, 3, 0, 0),
    (1, 3, 16, 32, 128, 256, 3, 0, 0),
    (2, 2, 24, 16, 64, 128, 1, 0, 0),     # 1x1x1 (shortcut), batch 2
    (1, 2, 16, 16, 64, 256, 3, 1, 0),     # pixel-shuffle store
    (1, 3, 8, 16, 64, 128, 3, 2, 1),      # frame-interleave store, first frame dropped
    (1, 3, 8, 16, 64, 128, 3, 2, 0),
    (1, 1, 8, 16, 64, 64, 3, 0, 0),       # single frame: two zero history frames
    # >= 148 tiles of 8x32 pixels with <= 128 output channels: the halo kernel (conv_halo.cu)
    (1, 3, 128, 128, 128, 128, 3, 0, 0),
    (1, 4, 88, 96, 256, 128, 3, 0, 0),    # H not a multiple of the 32-row tile, two channel blocks per tap
    (2, 2, 72, 80, 64, 64, 3, 0, 0),      # batch 2, 64 of the 128 accumulator rows live
    (1, 1, 256, 256, 128, 128, 3, 0, 0),  # single frame (all history out of bounds), one full image row per tile row
    (1, 3, 96, 128, 128, 256, 3, 0, 0),   # halo kernel, two 128-channel chunks per pixel tile
    (1, 2, 64, 128, 64, 512, 3, 1, 0),    # ... pixel-shuffle store
    (1, 3, 128, 64, 128, 256, 3, 2, 1),   # ... frame-interleave store, first frame dropped
    (1, 3, 128, 64, 128, 256, 3, 2, 0),
])
def test_conv3d_channels_last(lib, B, T, H, W, Cin, Cout, ks, store, drop):
    from deepv_b200 import _lib
    torch.manual_seed(2)
    x = (torch.randn(B, T, H, W, Cin, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(Cout, ks ** 3, Cin, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(max(Cout, 32), device="cuda")
    res = (torch.randn(B, T, H, W, Cout, device="cuda") * 0.5).bfloat16() if store == 0 else None
    oshape = {0: (B, T, H, W, Cout), 1: (B, T, 2 * H, 2 * W, Cout // 4),
              2: (B, 2 * T - drop, H, W, Cout // 2)}[store]
    out = torch.zeros(oshape, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.dv_conv3d_cl(p(x), p(w), p(bias), p(res), p(out), B, T, H, W, Cin, Cout, Cout, ks,
                                store, drop, None))
    torch.cuda.synchronize()
    xn = x.float().permute(0, 4, 1, 2, 3)
    wn = w.float().view(Cout, ks, ks, ks, Cin).permute(0, 4, 1, 2, 3)
    pad = ks // 2
    y = torch.nn.functional.conv3d(torch.nn.functional.pad(xn, (pad, pad, pad, pad, ks - 1, 0)), wn,
                                   bias[:Cout])
    if store == 0:
        ref = (y + res.float().permute(0, 4, 1, 2, 3)).permute(0, 2, 3, 4, 1)
    elif store == 1:  # packed rows (p1, p2, c) -> [b, t, 2h+p1, 2w+p2, c]
        Cq = Cout // 4
        ref = y.view(B, 2, 2, Cq, T, H, W).permute(0, 4, 5, 1, 6, 2, 3).reshape(B, T, 2 * H, 2 * W, Cq)
    else:             # packed rows (p, c) -> [b, 2t+p, h, w, c]
        Ch = Cout // 2
        ref = y.view(B, 2, Ch, T, H, W).permute(0, 3, 1, 4, 5, 2).reshape(B, 2 * T, H, W, Ch)[:, drop:]
    assert rel(out, ref) <= TOL


def _attn_ref(qkv, kv_end, key_bias, H):
    B, L, _ = qkv.shape
    D = H * 64
    q, k, v = (t.view(B, L, H, 64).transpose(1, 2) for t in qkv.float().split(D, dim=-1))
    s = q @ k.transpose(-1, -2) / 8.0 + key_bias[:, None, None, :L]
    vis = torch.arange(L, device=qkv.device)[None, :] < kv_end.to(qkv.device)[:, None]
    s = s.masked_fill(~vis[None, None], float("-inf"))
    return (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, L, D)


@pytest.mark.parametrize("B,L,H,groups,dead", [
    (1, 128, 1, [128], False),                        # one tile
    (1, 100, 2, [40, 60], False),                     # ragged single tile, context + 1 frame
    (2, 300, 2, [77, 96, 127], True),                 # frame boundary inside a tile + dead keys
    (2, 461, 3, [77, 192, 192], True),                # stage-1-like
    (3, 1000, 4, [269, 240, 192, 299], True),         # history-style dead prefix, 4 query tile pairs
])
def test_joint_attention(lib, B, L, H, groups, dead):
    from deepv_b200 import _lib
    assert sum(groups) == L
    torch.manual_seed(1)
    qkv = torch.randn(B, L, 3 * H * 64, device="cuda").bfloat16()
    bounds, acc = [], 0
    for g in groups:
        acc += g
        bounds.append(acc)
    kv_end = torch.empty(L, dtype=torch.int32)
    pos = 0
    for i, g in enumerate(groups):   # group 0 = context, sees frame 0 (group 1) as well
        kv_end[pos:pos + g] = bounds[max(i, 1)] if len(bounds) > 1 else bounds[0]
        pos += g
    Lpad = (L + 127) // 128 * 128
    kb = torch.zeros(B, Lpad, device="cuda")
    kb[:, L:] = float("-inf")
    if dead:
        kb[0, 5:40] = float("-inf")
        kb[-1, 0:min(192, groups[0] - 1)] = float("-inf")
    out = torch.zeros(B, L, H * 64, device="cuda", dtype=torch.bfloat16)
    kvd = kv_end.cuda()
    _lib.check(lib.dv_attention(p(qkv), p(out), p(kvd), p(kb), B, L, Lpad, H, None))
    torch.cuda.synchronize()
    ref = _attn_ref(qkv, kv_end, kb, H)
    assert torch.isfinite(out).all()
    assert rel(out, ref) <= TOL


def test_attention_large_scores_rescale_path(lib):
    """Scores spread over > 2^8 so the lazy running-max rescale of O in TMEM is exercised."""
    from deepv_b200 import _lib
    torch.manual_seed(4)
    B, L, H = 1, 512, 1
    qkv = torch.randn(B, L, 192, device="cuda")
    qkv[:, :, :128] *= 6.0   # |q.k|/8 up to ~100s, increasing along keys
    qkv[:, :, 64:128] *= torch.linspace(0.2, 2.0, L, device="cuda")[None, :, None]
    qkv = qkv.bfloat16()
    kv_end = torch.full((L,), L, dtype=torch.int32)
    kb = torch.zeros(B, 512, device="cuda")
    out = torch.zeros(B, L, 64, device="cuda", dtype=torch.bfloat16)
    kvd = kv_end.cuda()
    _lib.check(lib.dv_attention(p(qkv), p(out), p(kvd), p(kb), B, L, 512, H, None))
    torch.cuda.synchronize()
    assert rel(out, _attn_ref(qkv, kv_end, kb, H)) <= 2e-2


@pytest.mark.parametrize("B,T,H,W,Cin,Cout,stride", [
    (1, 3, 32, 64, 64, 256, (1, 2, 2)),     # spatial down-sample, wide output (regular tiles)
    (1, 2, 64, 64, 128, 128, (1, 2, 2)),    # ... Cout <= 128: swapped-operand tiles
    (1, 9, 16, 32, 64, 256, (2, 1, 1)),     # temporal down-sample: 9 -> 5 frames
    (1, 1, 16, 32, 128, 128, (2, 1, 1)),    # a single image: 1 -> 1 frame (history frames all zero)
    (2, 5, 32, 32, 64, 64, (2, 1, 1)),      # batch 2, 5 -> 3 frames
])
def test_strided_causal_conv3d(lib, B, T, H, W, Cin, Cout, stride):
    """The encoder's down-sampling convs (vae.py:322,346 through CausalConv3d :229-231,251)."""
    from deepv_b200 import _lib
    torch.manual_seed(3)
    x = (torch.randn(B, T, H, W, Cin, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(Cout, 27, Cin, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(max(Cout, 32), device="cuda")
    sT, sH, sW = stride
    oT, oH, oW = (T - 1) // sT + 1, H // sH, W // sW
    out = torch.zeros(B, oT, oH, oW, Cout, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.dv_conv3d_strided_cl(p(x), p(w), p(bias), p(out), B, T, H, W, Cin, Cout, Cout, sT, sH, sW, None))
    torch.cuda.synchronize()
    xn = x.float().permute(0, 4, 1, 2, 3)
    wn = w.float().view(Cout, 3, 3, 3, Cin).permute(0, 4, 1, 2, 3)
    y = torch.nn.functional.conv3d(torch.nn.functional.pad(xn, (1, 1, 1, 1, 2, 0)), wn, bias[:Cout], stride=stride)
    assert tuple(y.shape[2:]) == (oT, oH, oW)
    assert rel(out, y.permute(0, 2, 3, 4, 1)) <= TOL
