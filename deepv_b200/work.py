"""Algorithmic work of the hot path: which denoiser forwards a rollout issues, with which
token layouts, and the FLOPs of every forward / VAE decode / VAE encode.

Pure host arithmetic (no torch, no device): `bench.py` uses it to scale a bounded CPU sample of the
reference to the whole workload and to state `roofline.achieved` numerators; tests pin it against the
traced shapes of SURVEY.md App. B and the worked FLOP values of §8d.  The condition-clip logic mirrors
`generate_i2v` (pipeline.py:621-658) on shapes only.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

D_MODEL, N_LAYERS, C_LAT, JOINT_DIM, POOLED_DIM, TEXT_LEN = 1536, 24, 38, 4096, 2048, 77
NUM_INPUT_UNIT = 4   # pipeline.py:270


def condition_layout(unit_index: int, firstframe_mask: int, stage: int, frame_per_unit: int = 1) -> List[Tuple[int, int]]:
    """Condition clips of one stage as (frames, pyramid level), oldest first (pipeline.py:626-658):
    the newest clean frame at the stage's level, each older frame one level coarser, everything
    older than that (minus the masked first frame) as one level-0 clip."""
    fpu = frame_per_unit
    clips = [(fpu, stage)]
    cur, ptx = stage, 1
    while ptx < unit_index - firstframe_mask:
        cur = max(cur - 1, 0)
        if cur == 0:
            break
        ptx += 1
        clips.append((fpu, cur))
    if cur == 0 and ptx < unit_index - firstframe_mask:
        clips.append(((unit_index - firstframe_mask - ptx) * fpu, 0))
    return list(reversed(clips))


def clip_dims(layout: Sequence[Tuple[int, int]], stage: int, lat_h: int, lat_w: int, n_stages: int = 3):
    """(t, h, w) latent dims of the condition clips + the noisy clip of `stage` (the last entry)."""
    def hw(level):
        s = 2 ** (n_stages - 1 - level)
        return lat_h // s, lat_w // s
    return [(t, *hw(lv)) for t, lv in layout] + [(1, *hw(stage))]


def rollout_forwards(n_iterations: int, lat_h: int = 48, lat_w: int = 64, steps: Sequence[int] = (5, 5, 5),
                     units: int = 8) -> List[Dict]:
    """Every distinct (layout, count) of denoiser forwards in `generate()` with `n_iterations` iterations:
    iteration 0 = units 1..7+1 from one frame (CFG batch 2, no history, first frame masked), later ones =
    units 4..7 from 25 frames (CFG batch 3 with a history frame)."""
    out = []
    for it in range(n_iterations):
        first = it == 0
        unit_range = range(1, units + 1) if first else range(NUM_INPUT_UNIT, units)
        for u in unit_range:
            for st in range(len(steps)):
                lay = condition_layout(u, 1 if first else 0, st)
                out.append(dict(iteration=it, unit=u, stage=st, B=2 if first else 3, hist=not first,
                                clips=clip_dims(lay, st, lat_h, lat_w, len(steps)), count=steps[st]))
    return out


def mmdit_tokens(clips: Sequence[Tuple[int, int, int]], hist: bool, hist_hw=(48, 64)):
    lv = sum(t * (h // 2) * (w // 2) for t, h, w in clips)
    lc = TEXT_LEN + ((hist_hw[0] // 4) * (hist_hw[1] // 4) if hist else 0)
    return lv, lc


def mmdit_flops(B: int, clips: Sequence[Tuple[int, int, int]], hist: bool, hist_hw=(48, 64),
                n_layers: int = N_LAYERS, masked: bool = True) -> Dict[str, float]:
    """linear + attention FLOPs of one forward (all context rows counted, as
    dv_mmdit_plan_flops does; attention over the frame-causal visible pairs when `masked`)."""
    d, NL = float(D_MODEL), float(n_layers)
    lv, lc = mmdit_tokens(clips, hist, hist_hw)
    n_last = clips[-1][0] * (clips[-1][1] // 2) * (clips[-1][2] // 2)
    lin = 2.0 * B * (lv * 12 * d * d * NL + lc * (12 * d * d * (NL - 1) + 3 * d * d))
    lin += 2.0 * B * lv * (C_LAT * 4) * d + 2.0 * B * n_last * (C_LAT * 4) * d
    lin += 2.0 * B * TEXT_LEN * JOINT_DIM * d
    if hist:
        lin += 2.0 * B * (lc - TEXT_LEN) * (C_LAT * 4) * d
    mod_rows = (NL - 1) * 12 * d + 6 * d + 2 * d + 2 * d
    lin += 2.0 * B * d * mod_rows + 2.0 * B * (256 * d + d * d + POOLED_DIM * d + d * d)
    # visible keys per query: context + every frame up to the query's own (context queries see frame 0 too)
    frame_sizes = []
    for t, h, w in clips:
        frame_sizes += [(h // 2) * (w // 2)] * t
    if masked:
        pairs, seen = 0.0, float(lc)
        for i, n in enumerate(frame_sizes):
            seen += n
            pairs += n * seen
            if i == 0:
                pairs += lc * seen
    else:
        pairs = float(lv + lc) ** 2
    attn = 4.0 * d * NL * B * pairs
    return dict(linear=lin, attention=attn, total=lin + attn, Lv=lv, Lc=lc)


# ---- VAE (SURVEY.md App. A config / App. C shapes) -------------------------------------------------------
VAE_BLOCK_CHANNELS = (128, 256, 512, 512)


def _conv(t, h, w, cin, cout, k=3):
    return 2.0 * t * h * w * cout * cin * k ** 3


def decode_frame_schedule(first_frame: int, layers=(3, 3, 3, 3), spatial_up=(1, 1, 1, 0), temporal_up=(1, 1, 1, 0)):
    """First frame each conv of the decoder's up blocks has to compute when only output frames >= first_frame are
    wanted (the trimmed decode of a continuation iteration; csrc/vae.cu:run_tile walks the decoder backwards the same
    way): a causal 3x3x3 conv looks two frames back at its own temporal resolution, GroupNorm is per frame, the
    frame-interleave upsampler maps conv frame t to output frames 2t-1 and 2t.  Returns (res_t[i][j] = first output
    frame of resnet j of up block i, sp_t0[i], tp_t0[i] = first conv frame of the spatial / temporal upsampler,
    taps_t0 = first frame conv_out reads)."""
    n = max(0, first_frame - 2)
    taps_t0 = n
    res_t = [[0] * layers[i] for i in range(4)]
    sp_t0, tp_t0 = [0] * 4, [0] * 4
    for i in range(3, -1, -1):
        if temporal_up[i]:
            tp_t0[i] = (n + 1) // 2
            n = max(0, tp_t0[i] - 2)
        if spatial_up[i]:
            sp_t0[i] = n
            n = max(0, n - 2)
        for j in range(layers[i] - 1, -1, -1):
            res_t[i][j] = n
            n = max(0, n - 4)
    return res_t, sp_t0, tp_t0, taps_t0


def vae_decode_tile_flops(t_lat: int, h: int, w: int, chans=VAE_BLOCK_CHANNELS, layers=(3, 3, 3, 3), zc: int = 16,
                          first_frame: int = 0) -> float:
    """One tile [zc, t_lat, h, w] through post_quant_conv + decoder (vae.py:731-751); the temporal windows with
    their caches are the same contraction as one causal pass.  first_frame > 0: the FLOPs the
    trimmed decode executes (decode_frame_schedule); 0: the reference's work."""
    res_t, sp_t0, tp_t0, taps_t0 = decode_frame_schedule(first_frame, layers)
    c = list(reversed(chans))            # 512, 512, 256, 128
    f = _conv(t_lat, h, w, zc, zc, 1) + _conv(t_lat, h, w, zc, c[0])
    f += 4 * _conv(t_lat, h, w, c[0], c[0])                              # two mid resnets
    tok = h * w
    f += t_lat * (4 * 2.0 * tok * c[0] * c[0] + 2 * 2.0 * tok * tok * c[0])  # mid attention: q,k,v,out + QK^T, PV
    t, prev = t_lat, c[0]
    for i, co in enumerate(c):
        for j in range(layers[i]):
            ci = prev if j == 0 else co
            n = res_t[i][j]
            f += _conv(t - max(0, n - 2), h, w, ci, co) + _conv(t - n, h, w, co, co)
            if ci != co:
                f += _conv(t - n, h, w, ci, co, 1)
        prev = co
        if i < 3:                                                       # spatial then temporal up-sampling
            f += _conv(t - sp_t0[i], h, w, co, 4 * co)
            h, w = 2 * h, 2 * w
            f += _conv(t - tp_t0[i], h, w, co, 2 * co)
            t = 2 * t - 1                                               # first frame dropped (vae.py:408-409)
    f += _conv(t - taps_t0, h, w, c[-1], 3)
    return f


def tile_grid(lat_h: int, lat_w: int, tile: int = 32, stride: int = 24):
    """vae.py:990-997: latent tile origins every 24, tiles of <= 32."""
    return [(min(tile, lat_h - i), min(tile, lat_w - j)) for i in range(0, lat_h, stride) for j in range(0, lat_w, stride)]


def vae_decode_flops(t_lat: int, lat_h: int = 48, lat_w: int = 64, first_frame: int = 0) -> float:
    return sum(vae_decode_tile_flops(t_lat, th, tw, first_frame=first_frame) for th, tw in tile_grid(lat_h, lat_w))


def vae_encode_tile_flops(t_px: int, h: int, w: int, chans=VAE_BLOCK_CHANNELS, layers=(2, 2, 2, 2), zc: int = 16) -> float:
    """One pixel tile [3, t_px, h, w] through the encoder + quant_conv (vae.py:630-689), un-chunked."""
    f = _conv(t_px, h, w, 3, chans[0])
    t, prev = t_px, chans[0]
    for i, co in enumerate(chans):
        for j in range(layers[i]):
            ci = prev if j == 0 else co
            f += _conv(t, h, w, ci, co) + _conv(t, h, w, co, co)
            if ci != co:
                f += _conv(t, h, w, ci, co, 1)
        prev = co
        if i < 3:
            h, w = h // 2, w // 2
            f += _conv(t, h, w, co, co)                                 # stride (1,2,2)
            t = (t - 1) // 2 + 1
            f += _conv(t, h, w, co, co)                                 # stride (2,1,1)
    f += 4 * _conv(t, h, w, prev, prev)
    tok = h * w
    f += t * (4 * 2.0 * tok * prev * prev + 2 * 2.0 * tok * tok * prev)
    f += _conv(t, h, w, prev, 2 * zc) + _conv(t, h, w, 2 * zc, 2 * zc, 1)
    return f


def vae_encode_flops(t_px: int, H: int = 384, W: int = 512) -> float:
    tiles = [(min(256, H - i), min(256, W - j)) for i in range(0, H, 192) for j in range(0, W, 192)] \
        if (H > 256 or W > 256) else [(H, W)]
    return sum(vae_encode_tile_flops(t_px, th, tw) for th, tw in tiles)


def rollout_work(n_iterations: int = 2, lat_h: int = 48, lat_w: int = 64, steps=(5, 5, 5)) -> Dict[str, float]:
    """FLOPs and emitted frames of `generate()` (pipeline.py:264-424) with n_iterations iterations."""
    fw = rollout_forwards(n_iterations, lat_h, lat_w, steps)
    mm = sum(f["count"] * mmdit_flops(f["B"], f["clips"], f["hist"], (lat_h, lat_w))["total"] for f in fw)
    dec = n_iterations * 2 * vae_decode_flops(8, lat_h, lat_w)
    H, W = lat_h * 8, lat_w * 8
    enc = vae_encode_flops(1, H, W) + (n_iterations - 1) * (2 * vae_encode_flops(25, H, W))   # image (+ disparity)
    enc += n_iterations * 2 * vae_encode_flops(1, H, W)                                       # history frame, rgb + disparity
    frames = 57 + 32 * (n_iterations - 1)
    # what B200Rollout executes: continuation iterations decode only the 32 frames they keep (pipeline.py:327-328)
    dec_exec = 2 * vae_decode_flops(8, lat_h, lat_w) + (n_iterations - 1) * 2 * vae_decode_flops(8, lat_h, lat_w, first_frame=25)
    return dict(mmdit=mm, vae_decode=dec, vae_encode=enc, total=mm + dec + enc, frames=frames,
                forwards=sum(f["count"] for f in fw), vae_decode_executed=dec_exec, executed=mm + dec_exec + enc)
