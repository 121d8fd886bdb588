"""GPU bring-up probe: each building-block kernel against a plain torch reference.

Usage (on a B200 box):  python scripts/probe_kernels.py <gemm|attn|conv|sampler|all>
Each case runs in this process; run cases as separate processes so that a faulting kernel
cannot poison the CUDA context of the others.
"""
import ctypes as C
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
LIB = C.CDLL(str(Path(__file__).resolve().parents[1] / "deepv_b200" / "libdeepv_b200.so"))
LIB.dv_last_error.restype = C.c_char_p


def ck(rc, what):
    torch.cuda.synchronize()
    if rc != 0:
        raise RuntimeError(f"{what}: rc={rc} {LIB.dv_last_error().decode()}")


def p(t):
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-20)).item()


def case_gemm():
    torch.manual_seed(0)
    for (B, M, N, K, epi) in [(1, 128, 64, 64, 0), (2, 300, 384, 512, 0), (1, 1000, 1536, 1536, 1),
                              (3, 2000, 4608, 1536, 0), (2, 96, 1536, 6144, 0), (1, 4096, 2048, 1024, 0)]:
        A = (torch.randn(B, M, K, device="cuda") * 0.5).bfloat16()
        W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        bias = torch.randn(N, device="cuda")
        Cc = torch.zeros(B, M, N, device="cuda", dtype=torch.bfloat16)
        rc = LIB.dv_gemm_bf16(p(A), p(W), p(bias), p(Cc), B, M, N, K, epi, stream())
        ck(rc, "gemm")
        ref = A.float() @ W.float().t() + bias
        if epi == 1:
            ref = torch.nn.functional.gelu(ref, approximate="tanh")
        print(f"gemm B{B} M{M} N{N} K{K} epi{epi}: rel_err={rel(Cc, ref):.3e}", flush=True)
    # timing of a large one
    B, M, N, K = 1, 8192, 8192, 8192
    A = (torch.randn(B, M, K, device="cuda") * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    Cc = torch.zeros(B, M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        LIB.dv_gemm_bf16(p(A), p(W), None, p(Cc), B, M, N, K, 0, stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        LIB.dv_gemm_bf16(p(A), p(W), None, p(Cc), B, M, N, K, 0, stream())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"gemm 8192^3: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    ref = torch.matmul(A[0], W.t())
    print(f"  vs torch.matmul rel_err={rel(Cc[0], ref):.3e}")
    e0.record()
    for _ in range(10):
        torch.matmul(A[0], W.t())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"cublas 8192^3: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    for (M, N, K) in [(3072, 4608, 1536), (3072, 1536, 1536), (3072, 6144, 1536), (3072, 1536, 6144), (5904, 6144, 1536)]:
        A = (torch.randn(1, M, K, device="cuda") * 0.5).bfloat16()
        W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        Cc = torch.zeros(1, M, N, device="cuda", dtype=torch.bfloat16)
        for _ in range(3):
            LIB.dv_gemm_bf16(p(A), p(W), None, p(Cc), 1, M, N, K, 0, stream())
        e0.record()
        for _ in range(20):
            LIB.dv_gemm_bf16(p(A), p(W), None, p(Cc), 1, M, N, K, 0, stream())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        e0.record()
        for _ in range(20):
            torch.matmul(A[0], W.t())
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / 20
        print(f"gemm M{M} N{N} K{K}: ours {ms*1e3:.1f} us {2*M*N*K/ms/1e9:.0f} TF/s | cublas {ms2*1e3:.1f} us {2*M*N*K/ms2/1e9:.0f} TF/s", flush=True)


def attn_ref(qkv, kv_end, key_bias, H):
    B, L, _ = qkv.shape
    D = H * 64
    q, k, v = qkv.float().split(D, dim=-1)
    q = q.view(B, L, H, 64).transpose(1, 2)
    k = k.view(B, L, H, 64).transpose(1, 2)
    v = v.view(B, L, H, 64).transpose(1, 2)
    s = q @ k.transpose(-1, -2) / 8.0
    ar = torch.arange(L, device=qkv.device)
    mask = ar[None, :] < kv_end[:, None].to(qkv.device)          # [Lq, Lk]
    s = s + key_bias[:, None, None, :L]
    s = s.masked_fill(~mask[None, None], float("-inf"))
    o = torch.softmax(s, dim=-1) @ v
    return o.transpose(1, 2).reshape(B, L, D)


def case_attn():
    torch.manual_seed(1)
    for (B, L, H, frames) in [(1, 128, 1, [128]), (2, 300, 2, [77, 96, 127]), (3, 2237, 24, [269, 240, 192, 768, 768])]:
        D = H * 64
        qkv = (torch.randn(B, L, 3 * D, device="cuda")).bfloat16()
        # kv_end: prefix ends by "frame" groups
        ends, acc = [], 0
        sizes = frames
        assert sum(sizes) == L, (sum(sizes), L)
        # first group = context + frame 0 share visibility
        bounds = []
        for sidx, sz in enumerate(sizes):
            acc += sz
            bounds.append(acc)
        kv_end = torch.empty(L, dtype=torch.int32)
        pos = 0
        for sidx, sz in enumerate(sizes):
            vis = bounds[max(sidx, 1)] if len(bounds) > 1 else bounds[0]
            kv_end[pos:pos + sz] = vis
            pos += sz
        Lpad = (L + 127) // 128 * 128
        key_bias = torch.zeros(B, Lpad, device="cuda")
        key_bias[:, L:] = float("-inf")
        if L > 100:
            key_bias[0, 5:40] = float("-inf")
            if B > 1:
                key_bias[1, 0:70] = float("-inf")
        out = torch.zeros(B, L, D, device="cuda", dtype=torch.bfloat16)
        kvd = kv_end.cuda()
        rc = LIB.dv_attention(p(qkv), p(out), p(kvd), p(key_bias), B, L, Lpad, H, stream())
        ck(rc, "attention")
        ref = attn_ref(qkv, kv_end, key_bias, H)
        print(f"attn B{B} L{L} H{H}: rel_err={rel(out, ref):.3e}", flush=True)
        if L > 2000:
            for _ in range(3):
                LIB.dv_attention(p(qkv), p(out), p(kvd), p(key_bias), B, L, Lpad, H, stream())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                LIB.dv_attention(p(qkv), p(out), p(kvd), p(key_bias), B, L, Lpad, H, stream())
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            pairs = kv_end.double().sum().item()
            fl = 4 * 64 * H * B * pairs
            print(f"  attn time {ms*1e3:.1f} us, {fl/ms/1e9:.1f} TFLOP/s (masked pairs)", flush=True)


def case_conv():
    torch.manual_seed(2)
    for (B, T, H, W, Cin, Cout, ks, store, drop) in [
            (1, 2, 8, 16, 64, 64, 3, 0, 0), (1, 3, 16, 32, 128, 256, 3, 0, 0), (2, 2, 24, 16, 64, 128, 1, 0, 0),
            (1, 2, 16, 16, 64, 256, 3, 1, 0), (1, 3, 8, 16, 64, 128, 3, 2, 1), (1, 3, 8, 16, 64, 128, 3, 2, 0),
            (1, 9, 64, 64, 128, 128, 3, 0, 0)]:
        x = (torch.randn(B, T, H, W, Cin, device="cuda") * 0.5).bfloat16()
        taps = ks ** 3
        w = (torch.randn(Cout, taps, Cin, device="cuda") * 0.05).bfloat16()
        bias = torch.randn(max(Cout, 32), device="cuda")
        res = None
        if store == 0:
            res = (torch.randn(B, T, H, W, Cout, device="cuda") * 0.5).bfloat16()
        if store == 1:
            oshape = (B, T, 2 * H, 2 * W, Cout // 4)
        elif store == 2:
            oshape = (B, 2 * T - drop, H, W, Cout // 2)
        else:
            oshape = (B, T, H, W, Cout)
        out = torch.zeros(oshape, device="cuda", dtype=torch.bfloat16)
        rc = LIB.dv_conv3d_cl(p(x), p(w), p(bias), p(res) if res is not None else None, p(out), B, T, H, W,
                              Cin, Cout, Cout, ks, store, drop, stream())
        ck(rc, "conv")
        # reference: NCDHW conv3d with causal zero padding
        xn = x.float().permute(0, 4, 1, 2, 3)
        wn = w.float().view(Cout, ks, ks, ks, Cin).permute(0, 4, 1, 2, 3)
        pad = ks // 2
        xp = torch.nn.functional.pad(xn, (pad, pad, pad, pad, ks - 1, 0))
        y = torch.nn.functional.conv3d(xp, wn, bias[:Cout])          # [B, Cout, T, H, W]
        if store == 0:
            y = y + res.float().permute(0, 4, 1, 2, 3)
            ref = y.permute(0, 2, 3, 4, 1)
        elif store == 1:
            Cq = Cout // 4   # our packed row order: (p1, p2, c)
            y = y.view(B, 2, 2, Cq, T, H, W).permute(0, 4, 5, 1, 6, 2, 3).reshape(B, T, 2 * H, 2 * W, Cq)
            ref = y
        else:
            Ch = Cout // 2   # packed row order: (p, c)
            y = y.view(B, 2, Ch, T, H, W).permute(0, 3, 1, 4, 5, 2).reshape(B, 2 * T, H, W, Ch)
            ref = y[:, drop:]
        print(f"conv B{B} T{T} H{H} W{W} Cin{Cin} Cout{Cout} k{ks} store{store} drop{drop}: rel_err={rel(out, ref):.3e}", flush=True)
    # timing: C=128 @ 256x256 x 8 frames
    B, T, H, W, Cin, Cout = 1, 8, 256, 256, 128, 128
    x = (torch.randn(B, T, H, W, Cin, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(Cout, 27, Cin, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(Cout, device="cuda")
    out = torch.zeros(B, T, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        LIB.dv_conv3d_cl(p(x), p(w), p(bias), None, p(out), B, T, H, W, Cin, Cout, Cout, 3, 0, 0, stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        LIB.dv_conv3d_cl(p(x), p(w), p(bias), None, p(out), B, T, H, W, Cin, Cout, Cout, 3, 0, 0, stream())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 2.0 * B * T * H * W * Cout * 27 * Cin
    print(f"conv 128->128 @8x256x256: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
    B, T, H, W, Cin, Cout = 1, 2, 64, 64, 512, 512
    x = (torch.randn(B, T, H, W, Cin, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(Cout, 27, Cin, device="cuda") * 0.05).bfloat16()
    out = torch.zeros(B, T, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        LIB.dv_conv3d_cl(p(x), p(w), None, None, p(out), B, T, H, W, Cin, Cout, Cout, 3, 0, 0, stream())
    e0.record()
    for _ in range(5):
        LIB.dv_conv3d_cl(p(x), p(w), None, None, p(out), B, T, H, W, Cin, Cout, Cout, 3, 0, 0, stream())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 2.0 * B * T * H * W * Cout * 27 * Cin
    print(f"conv 512->512 @2x64x64: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s", flush=True)


def case_sampler():
    torch.manual_seed(3)
    n = 38 * 48 * 64
    for nb in (1, 2, 3):
        pred = torch.randn(nb, n, device="cuda").bfloat16()
        x = torch.randn(n, device="cuda").bfloat16()
        out = torch.empty_like(x)
        sig = torch.tensor([1.0, 0.75025, 0.5005, 0.25075000000000003, 0.0010000000000000009, 0.0], dtype=torch.float64, device="cuda")
        for i in (0, 4):
            rc = LIB.dv_cfg_euler_step(p(pred), nb, p(x), p(out), C.c_longlong(n), C.c_float(3.5), C.c_float(6.0),
                                       C.c_double(sig[i].item()), C.c_double(sig[i + 1].item()), 1, stream())
            ck(rc, "cfg_euler")
            if nb == 1:
                g = pred[0]
            elif nb == 2:
                u, t = pred[0], pred[1]
                g = u + 3.5 * (t - u)
            else:
                u, t, h = pred[0], pred[1], pred[2]
                g = u + 3.5 * (t - u) + 6.0 * (h - t)
            ref = (x.to(torch.float32) + (sig[i + 1] - sig[i]) * g).to(torch.bfloat16)
            print(f"cfg_euler nb{nb} step{i}: bit_exact={torch.equal(ref, out)} maxdiff={(ref.float()-out.float()).abs().max().item():.3e}", flush=True)
    lo = torch.randn(38, 12, 16, device="cuda").bfloat16()
    nz = torch.randn(38, 24, 32, device="cuda").bfloat16()
    out = torch.empty_like(nz)
    a, b = 0.5998620228185144, 0.6930419797086197
    rc = LIB.dv_stage_renoise(p(lo), p(nz), p(out), 38, 12, 16, C.c_double(a), C.c_double(b), 1, stream())
    ck(rc, "renoise")
    up = torch.nn.functional.interpolate(lo[None], size=(24, 32), mode="nearest")[0]
    ref = a * up + b * nz
    print(f"stage_renoise: bit_exact={torch.equal(ref, out)}", flush=True)
    z = torch.randn(200000, 4, device="cuda")
    outn = torch.empty(100, 40, 200, device="cuda")
    rc = LIB.dv_block_noise(p(z), p(outn), 100, 40, 200, C.c_float(1 / 3), 0, stream())
    ck(rc, "block_noise")
    blk = outn.view(100, 20, 2, 100, 2).permute(0, 1, 3, 2, 4).reshape(-1, 4)
    cov = (blk.t() @ blk) / blk.shape[0]
    print("block_noise cov:\n", cov.cpu().numpy().round(3), flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    print(torch.cuda.get_device_name(0), flush=True)
    cases = {"gemm": case_gemm, "attn": case_attn, "conv": case_conv, "sampler": case_sampler}
    for name, fn in cases.items():
        if which in (name, "all"):
            fn()
