"""CPU emulation of the data movement of `conv_halo_kernel` (deepv_b200/csrc/conv_halo.cu): one zero-filled
(32+2) x (8+2) pixel box per (frame tap, 64-channel block), rows r = y*10 + x, and every spatial tap read as
the row-shifted view  rows[(dy*10 + dx) + y*10 + x]  — against a plain causal conv3d.  It pins the index
arithmetic (tile decode, halo origin, weight K order, the three store maps) without a GPU; the kernel itself is
checked on the B200 by tests/test_gpu_kernels.py::test_conv3d_channels_last."""
import pytest
import torch
import torch.nn.functional as F

TILE_W, TILE_H, BOX_W, BOX_H = 8, 32, 10, 34


def decode_tile(tile, n_chunks, tiles_w, tiles_h, T, t0):
    """`decode()` of conv_halo.cu: flat tile index -> (batch, frame, h0, w0, weight chunk); frames [t0, T) only."""
    chunk = tile % n_chunks
    tile //= n_chunks
    per_frame = tiles_w * tiles_h
    r, f = tile % per_frame, tile // per_frame
    nt = T - t0
    return f // nt, t0 + f % nt, (r // tiles_w) * TILE_H, (r % tiles_w) * TILE_W, chunk


def halo_conv_emulated(x, w, bias, store, drop, t0=0):
    """x [T,H,W,C] (C % 64 == 0), w [Cout, 27, C] (tap-major K, tap = (dt*3+dy)*3+dx) -> stored tensor.
    t0 > 0 (GemmDesc::conv_t0, the trimmed decode): only conv frames [t0, T) are computed, indices stay absolute."""
    T, H, W, C = x.shape
    Cout = w.shape[0]
    c_blocks = C // 64
    n_chunks = (Cout + 127) // 128
    out_C = {0: Cout, 1: Cout // 4, 2: Cout // 2}[store]
    oT = 2 * T - drop if store == 2 else T
    oH, oW = (2 * H, 2 * W) if store == 1 else (H, W)
    out = torch.zeros(oT, oH, oW, out_C)
    wk = w.reshape(Cout, 27 * C)
    tiles_w, tiles_h = W // TILE_W, (H + TILE_H - 1) // TILE_H
    seen = set()
    for tile in range((T - t0) * tiles_w * tiles_h * n_chunks):
        if True:
            if True:
                if True:
                    b, t, h0, w0, chunk = decode_tile(tile, n_chunks, tiles_w, tiles_h, T, t0)
                    assert b == 0 and t0 <= t < T and (t, h0, w0, chunk) not in seen
                    seen.add((t, h0, w0, chunk))
                    rows_w = wk[chunk * 128:(chunk + 1) * 128]                     # TMA box of 128 weight rows
                    acc = torch.zeros(rows_w.shape[0], TILE_H * TILE_W)            # D[channel][pixel n = y*8 + x]
                    for dt in range(3):
                        for cb in range(c_blocks):
                            box = torch.zeros(BOX_H * BOX_W, 64)                   # out-of-bounds = zero fill
                            tt = t + dt - 2
                            for y in range(BOX_H):
                                for xx in range(BOX_W):
                                    yy, xw = h0 - 1 + y, w0 - 1 + xx
                                    if 0 <= tt < T and 0 <= yy < H and 0 <= xw < W:
                                        box[y * BOX_W + xx] = x[tt, yy, xw, cb * 64:(cb + 1) * 64]
                            for s in range(9):
                                dy, dx = divmod(s, 3)
                                kb = (dt * 9 + s) * c_blocks + cb
                                a = rows_w[:, kb * 64:(kb + 1) * 64]
                                start = dy * BOX_W + dx
                                idx = torch.tensor([start + y * BOX_W + xx for y in range(TILE_H) for xx in range(TILE_W)])
                                acc += a @ box[idx].T
                    for r in range(rows_w.shape[0]):
                        n = chunk * 128 + r
                        if n >= Cout:
                            continue
                        oc, ot, sh, p1, p2 = n, t, 1, 0, 0
                        if store == 1:
                            q, oc = divmod(n, out_C)
                            p1, p2, sh = q >> 1, q & 1, 2
                        elif store == 2:
                            p, oc = divmod(n, out_C)
                            ot = 2 * t + p - drop
                        if ot < 0:
                            continue
                        for y in range(TILE_H):
                            if h0 + y >= H:
                                continue
                            for xx in range(TILE_W):
                                out[ot, sh * (h0 + y) + p1, sh * (w0 + xx) + p2, oc] = acc[r, y * TILE_W + xx] + bias[n]
    return out


@pytest.mark.parametrize("T,H,W,C,Cout,store,drop", [
    (2, 40, 16, 64, 64, 0, 0),      # H not a multiple of the 32-row tile
    (2, 32, 8, 64, 256, 1, 0),      # two 128-channel chunks, pixel-shuffle store
    (2, 32, 8, 128, 128, 2, 1),     # two channel blocks per tap, frame-interleave store with the first frame dropped
])
def test_halo_tile_arithmetic_matches_conv3d(T, H, W, C, Cout, store, drop):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(T, H, W, C, generator=g)
    w = torch.randn(Cout, 27, C, generator=g) * 0.05
    bias = torch.randn(Cout, generator=g)
    got = halo_conv_emulated(x, w, bias, store, drop)
    xn = x.permute(3, 0, 1, 2).unsqueeze(0)
    wn = w.view(Cout, 3, 3, 3, C).permute(0, 4, 1, 2, 3)
    y = F.conv3d(F.pad(xn, (1, 1, 1, 1, 2, 0)), wn, bias)[0]                       # [Cout, T, H, W]
    if store == 0:
        ref = y.permute(1, 2, 3, 0)
    elif store == 1:
        Cq = Cout // 4
        ref = y.view(2, 2, Cq, T, H, W).permute(3, 4, 0, 5, 1, 2).reshape(T, 2 * H, 2 * W, Cq)
    else:
        Ch = Cout // 2
        ref = y.view(2, Ch, T, H, W).permute(2, 0, 3, 4, 1).reshape(2 * T, H, W, Ch)[drop:]
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() <= 1e-3 * ref.abs().max().item()


@pytest.mark.parametrize("store,drop,t0", [(0, 0, 1), (0, 0, 3), (2, 1, 2), (1, 0, 2)])
def test_halo_tiles_of_a_trimmed_conv(store, drop, t0):
    """Frames [t0, T) of the conv output (after the store map: 2*t0 - drop for the frame-interleave store) equal the full
    conv; nothing in front is written; every (frame, tile, chunk) is visited exactly once."""
    T, H, W, C, Cout = 4, 32, 8, 64, 128
    g = torch.Generator().manual_seed(4)
    x = torch.randn(T, H, W, C, generator=g)
    w = torch.randn(Cout, 27, C, generator=g) * 0.05
    bias = torch.randn(Cout, generator=g)
    full = halo_conv_emulated(x, w, bias, store, drop)
    part = halo_conv_emulated(x, w, bias, store, drop, t0)
    first = max(0, 2 * t0 - drop) if store == 2 else t0
    assert torch.equal(part[first:], full[first:])
    assert not part[:first].any()
