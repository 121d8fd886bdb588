"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of the reference VAE decode path.

Restates /root/reference/model/vae.py: CausalConv3d with the temporal-chunk feature cache
(:225-252), per-frame GroupNorm (:161-167), resnet / mid / up blocks (:293-310, :459-469,
:551-563), pixel-shuffle and frame-interleave up-samplers (:376-383, :401-410), the decoder
(:731-751), chunk_decode (:903-920), tiled_decode with its sequential in-place blends
(:942-952, :989-1014), and InferencePipeline.decode_latent's un-normalisation
(pipeline.py:703-713).  The mid-block attention restates diffusers 0.31.0 `Attention` as used at
vae.py:439-445 (see oracle/_shim.py).

Pinned against the real reference by tests/test_oracle_vs_reference.py and tests/golden/.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


class DecoderState:
    """Feature caches of every causal conv between temporal windows (vae.py:202,238,249)."""

    def __init__(self):
        self.cache: Dict[str, Tensor] = {}


def causal_conv(W, name: str, x: Tensor, st: Optional[DecoderState], is_init: bool) -> Tensor:
    w, b = W[name + ".conv.weight"], W[name + ".conv.bias"]
    k = w.shape[2]
    sp = w.shape[3] // 2
    if st is None or is_init:
        x = F.pad(x, (sp, sp, sp, sp, k - 1, 0))            # zero causal + spatial pad (:229-236)
    else:
        x = F.pad(x, (sp, sp, sp, sp, 0, 0))
        if k == 3:
            x = torch.cat([st.cache[name], x], dim=2)        # last 2 padded frames (:244-245)
    if st is not None:
        st.cache[name] = x[:, :, -2:].clone()
    return F.conv3d(x, w, b)


def frame_group_norm(W, name: str, x: Tensor, groups: int) -> Tensor:
    b, c, t, h, w = x.shape
    y = F.group_norm(x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w), groups, W[name + ".weight"],
                     W[name + ".bias"], 1e-6)
    return y.view(b, t, c, h, w).permute(0, 2, 1, 3, 4)


def resnet(W, name: str, x: Tensor, st, is_init: bool, groups: int) -> Tensor:
    h = causal_conv(W, name + ".conv1", F.silu(frame_group_norm(W, name + ".norm1", x, groups)), st, is_init)
    h = causal_conv(W, name + ".conv2", F.silu(frame_group_norm(W, name + ".norm2", h, groups)), st, is_init)
    if name + ".conv_shortcut.conv.weight" in W:
        x = causal_conv(W, name + ".conv_shortcut", x, st, is_init)
    return x + h


def mid_attention(W, name: str, x: Tensor, groups: int) -> Tensor:
    """Per frame: GN -> q,k,v -> single-head softmax(QK^T / sqrt(C)) V -> out -> + residual."""
    b, c, t, h, w = x.shape
    xf = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    y = F.group_norm(xf, groups, W[name + ".group_norm.weight"], W[name + ".group_norm.bias"], 1e-6)
    y = y.view(b * t, c, h * w).transpose(1, 2)
    q = F.linear(y, W[name + ".to_q.weight"], W[name + ".to_q.bias"])
    k = F.linear(y, W[name + ".to_k.weight"], W[name + ".to_k.bias"])
    v = F.linear(y, W[name + ".to_v.weight"], W[name + ".to_v.bias"])
    o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
    o = F.linear(o, W[name + ".to_out.0.weight"], W[name + ".to_out.0.bias"])
    o = o.transpose(1, 2).reshape(b * t, c, h, w) + xf
    return o.view(b, t, c, h, w).permute(0, 2, 1, 3, 4)


def decoder_forward(W, cfg: dict, z: Tensor, st: Optional[DecoderState], is_init: bool) -> Tensor:
    """post_quant_conv + CausalVaeDecoder.forward for one temporal window."""
    g = cfg.get("decoder_norm_num_groups", 32)
    chans = list(reversed(cfg["decoder_block_out_channels"]))
    x = causal_conv(W, "post_quant_conv", z, st, is_init)
    x = causal_conv(W, "decoder.conv_in", x, st, is_init)
    x = resnet(W, "decoder.mid_block.resnets.0", x, st, is_init, g)
    x = mid_attention(W, "decoder.mid_block.attentions.0", x, g)
    x = resnet(W, "decoder.mid_block.resnets.1", x, st, is_init, g)
    for i in range(len(chans)):
        for j in range(cfg["decoder_layers_per_block"][i]):
            x = resnet(W, f"decoder.up_blocks.{i}.resnets.{j}", x, st, is_init, g)
        if cfg["decoder_spatial_up_sample"][i]:
            x = causal_conv(W, f"decoder.up_blocks.{i}.upsamplers.0.conv", x, st, is_init)
            b, c4, t, h, w = x.shape  # 'b (c p1 p2) t h w -> b c t (h p1) (w p2)'
            x = x.view(b, c4 // 4, 2, 2, t, h, w).permute(0, 1, 4, 5, 2, 6, 3).reshape(b, c4 // 4, t, 2 * h, 2 * w)
        if cfg["decoder_temporal_up_sample"][i]:
            x = causal_conv(W, f"decoder.up_blocks.{i}.temporal_upsamplers.0.conv", x, st, is_init)
            b, c2, t, h, w = x.shape  # 'b (c p) t h w -> b c (t p) h w'
            x = x.view(b, c2 // 2, 2, t, h, w).permute(0, 1, 3, 2, 4, 5).reshape(b, c2 // 2, 2 * t, h, w)
            if is_init:
                x = x[:, :, 1:]
    x = F.silu(frame_group_norm(W, "decoder.conv_norm_out", x, g))
    return causal_conv(W, "decoder.conv_out", x, st, is_init)


def chunk_decode(W, cfg, z: Tensor, window_size: int = 1) -> Tensor:
    """vae.py:903-920: first window = window_size + 1 latent frames, then window_size each."""
    n = z.shape[2]
    cuts = [(0, window_size + 1)]
    fid = window_size + 1
    for _ in range((n - fid) // window_size):
        cuts.append((fid, fid + window_size))
        fid += window_size
    if fid < n:
        cuts.append((fid, n))
    st = DecoderState()
    outs = [decoder_forward(W, cfg, z[:, :, a:b], st, i == 0) for i, (a, b) in enumerate(cuts)]
    return torch.cat(outs, dim=2)


def full_decode(W, cfg, z: Tensor) -> Tensor:
    """temporal_chunk=False path (vae.py:896-897): one pass over all latent frames."""
    return decoder_forward(W, cfg, z, None, True)


def _blend_v(a: Tensor, b: Tensor, extent: int) -> Tensor:
    extent = min(a.shape[3], b.shape[3], extent)
    for y in range(extent):
        b[:, :, :, y, :] = a[:, :, :, -extent + y, :] * (1 - y / extent) + b[:, :, :, y, :] * (y / extent)
    return b


def _blend_h(a: Tensor, b: Tensor, extent: int) -> Tensor:
    extent = min(a.shape[4], b.shape[4], extent)
    for x in range(extent):
        b[:, :, :, :, x] = a[:, :, :, :, -extent + x] * (1 - x / extent) + b[:, :, :, :, x] * (x / extent)
    return b


def tiled_decode(W, cfg, z: Tensor, tile_sample_min_size: int = 256, window_size: int = 1,
                 temporal_chunk: bool = True, scale: int = 8) -> Tensor:
    """vae.py:885-900, 989-1014: tiles of tile/8 latent pixels every 3/4 tile, blends of 1/4 tile
    written IN PLACE in row-major tile order (later tiles read already-blended neighbours)."""
    tl = int(tile_sample_min_size / scale)

    def one(zt):
        return chunk_decode(W, cfg, zt, window_size) if temporal_chunk else full_decode(W, cfg, zt)

    if not (z.shape[-1] > tl or z.shape[-2] > tl):
        return one(z)
    overlap = int(tl * 0.75)
    extent = int(tile_sample_min_size * 0.25)
    limit = tile_sample_min_size - extent
    rows: List[List[Tensor]] = []
    for i in range(0, z.shape[3], overlap):
        rows.append([one(z[:, :, :, i:i + tl, j:j + tl]) for j in range(0, z.shape[4], overlap)])
    out_rows = []
    for i, row in enumerate(rows):
        res = []
        for j, tile in enumerate(row):
            if i > 0:
                tile = _blend_v(rows[i - 1][j], tile, extent)
            if j > 0:
                tile = _blend_h(row[j - 1], tile, extent)
            res.append(tile[:, :, :, :limit, :limit])
        out_rows.append(torch.cat(res, dim=4))
    return torch.cat(out_rows, dim=3)


VAE_SHIFT, VAE_SCALE = 0.1490, 1 / 1.8415              # pipeline.py:194-195
VAE_VIDEO_SHIFT, VAE_VIDEO_SCALE = -0.2343, 1 / 3.0986  # pipeline.py:196-197


def unnormalise_latents(lat: Tensor) -> Tensor:
    """pipeline.py:705-709 (frame 0 uses the image statistics, the rest the video ones)."""
    lat = lat.clone()
    if lat.shape[2] == 1:
        return lat / VAE_SCALE + VAE_SHIFT
    lat[:, :, :1] = lat[:, :, :1] / VAE_SCALE + VAE_SHIFT
    lat[:, :, 1:] = lat[:, :, 1:] / VAE_VIDEO_SCALE + VAE_VIDEO_SHIFT
    return lat


def decode_latent(W, cfg, lat: Tensor) -> Tensor:
    """pipeline.py:703-713 with save_memory=True."""
    return tiled_decode(W, cfg, unnormalise_latents(lat), 256, 1, True)


# ---------------------------------------------------------------------------------------------
# Encode side: vae.py:630-689 (encoder), :844-883 (encode / tiled_encode as
# the rollout calls it: `vae.encode(x)` with tiling on -> temporal_chunk=False, 256-px tiles every
# 192 px, 8-latent blends, 24-latent crops), :599-615 (DiagonalGaussianDistribution).
# ---------------------------------------------------------------------------------------------
def strided_causal_conv(W, name: str, x: Tensor, stride=(1, 1, 1)) -> Tensor:
    """CausalConv3d without the chunk cache (vae.py:229-231,251): zero pad (k-1) frames in front
    and k//2 pixels around, then a VALID conv with the given stride."""
    w, b = W[name + ".conv.weight"], W[name + ".conv.bias"]
    k = w.shape[2]
    sp = w.shape[3] // 2
    return F.conv3d(F.pad(x, (sp, sp, sp, sp, k - 1, 0)), w, b, stride=stride)


def encoder_forward(W, cfg: dict, x: Tensor) -> Tensor:
    """CausalVaeEncoder.forward + quant_conv for one tile, temporal_chunk=False -> moments [b,2z,t,h,w]."""
    g = cfg.get("encoder_norm_num_groups", 32)
    chans = list(cfg["encoder_block_out_channels"])

    def res(name, h):
        return resnet(W, name, h, None, True, g)

    h = strided_causal_conv(W, "encoder.conv_in", x)
    for i in range(len(chans)):
        for j in range(cfg["encoder_layers_per_block"][i]):
            h = res(f"encoder.down_blocks.{i}.resnets.{j}", h)
        if cfg["encoder_spatial_down_sample"][i]:
            h = strided_causal_conv(W, f"encoder.down_blocks.{i}.downsamplers.0.conv", h, (1, 2, 2))
        if cfg["encoder_temporal_down_sample"][i]:
            h = strided_causal_conv(W, f"encoder.down_blocks.{i}.temporal_downsamplers.0.conv", h, (2, 1, 1))
    h = res("encoder.mid_block.resnets.0", h)
    h = mid_attention(W, "encoder.mid_block.attentions.0", h, g)
    h = res("encoder.mid_block.resnets.1", h)
    h = F.silu(frame_group_norm(W, "encoder.conv_norm_out", h, g))
    h = strided_causal_conv(W, "encoder.conv_out", h)
    return strided_causal_conv(W, "quant_conv", h)


def tiled_encode(W, cfg, x: Tensor, tile_sample_min_size: int = 256, scale: int = 8) -> Tensor:
    """vae.py:844-851,954-987 with temporal_chunk=False: moments of the whole frame."""
    ts = tile_sample_min_size
    tl = int(ts / scale)
    if not (x.shape[-1] > ts or x.shape[-2] > ts):
        return encoder_forward(W, cfg, x)
    overlap = int(ts * 0.75)
    extent = int(tl * 0.25)
    limit = tl - extent
    rows: List[List[Tensor]] = []
    for i in range(0, x.shape[3], overlap):
        rows.append([encoder_forward(W, cfg, x[:, :, :, i:i + ts, j:j + ts]) for j in range(0, x.shape[4], overlap)])
    out_rows = []
    for i, row in enumerate(rows):
        resl = []
        for j, tile in enumerate(row):
            if i > 0:
                tile = _blend_v(rows[i - 1][j], tile, extent)
            if j > 0:
                tile = _blend_h(row[j - 1], tile, extent)
            resl.append(tile[:, :, :, :limit, :limit])
        out_rows.append(torch.cat(resl, dim=4))
    return torch.cat(out_rows, dim=3)


def gaussian_sample(moments: Tensor, noise: Tensor) -> Tensor:
    """DiagonalGaussianDistribution(moments).sample() with the normal draw injected (vae.py:602-615)."""
    mean, logvar = torch.chunk(moments, 2, dim=1)
    return mean + torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0)) * noise
