"""Deterministic synthetic ("random-init") weights for the parity tests, smoke() and the bench.

Checkpoints are unavailable offline, and the reference's own random init
zeroes every adaLN / output layer so that the denoiser output is identically 0 (SURVEY.md §0.5,
model/mmdit.py:1275-1286).  This module therefore enumerates the reference's parameter names
and shapes (state_dict keys of MMDiT / CausalVideoVAE.decoder for a given config) and fills them
with seeded, non-degenerate values.  The same dict feeds the real reference (load_state_dict),
the oracle restatement and the CUDA path.

Names follow model/mmdit.py:1194-1238 and model/vae.py:701-728,823-824.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from typing import Dict, Tuple

import torch

MMDIT_DEFAULT = dict(
    num_layers=24, num_attention_heads=24, attention_head_dim=64, in_channels=38, patch_size=2,
    sample_size=128, pos_embed_max_size=192, joint_attention_dim=4096, pooled_projection_dim=2048,
    caption_projection_dim=1536, max_num_frames=200, qk_norm="rms_norm", pos_embed_type="sincos",
    temp_pos_embed_type="rope", add_temp_pos_embed=True, use_temporal_causal=True,
    interp_condition_pos=True,
)

VAE_DEFAULT = dict(
    encoder_out_channels=16, decoder_in_channels=16,
    encoder_block_out_channels=(128, 256, 512, 512), decoder_block_out_channels=(128, 256, 512, 512),
    encoder_layers_per_block=(2, 2, 2, 2), decoder_layers_per_block=(3, 3, 3, 3),
    encoder_spatial_down_sample=(True, True, True, False), decoder_spatial_up_sample=(True, True, True, False),
    encoder_temporal_down_sample=(True, True, True, False), decoder_temporal_up_sample=(True, True, True, False),
    interpolate=False,
)

Shapes = "OrderedDict[str, Tuple[int, ...]]"


def mmdit_shapes(cfg: dict) -> Shapes:
    d = cfg["num_attention_heads"] * cfg["attention_head_dim"]
    hd = cfg["attention_head_dim"]
    c, p = cfg["in_channels"], cfg["patch_size"]
    s: Shapes = OrderedDict()

    def lin(name, out_f, in_f):
        s[name + ".weight"] = (out_f, in_f)
        s[name + ".bias"] = (out_f,)

    s["pos_embed.proj.weight"] = (d, c, p, p)
    s["pos_embed.proj.bias"] = (d,)
    s["pos_embed.proj_history.weight"] = (d, c, p, p)
    s["pos_embed.proj_history.bias"] = (d,)
    lin("time_text_embed.timestep_embedder.linear_1", d, 256)
    lin("time_text_embed.timestep_embedder.linear_2", d, d)
    lin("time_text_embed.text_embedder.linear_1", d, cfg["pooled_projection_dim"])
    lin("time_text_embed.text_embedder.linear_2", d, d)
    lin("context_embedder", cfg["caption_projection_dim"], cfg["joint_attention_dim"])
    nl = cfg["num_layers"]
    for i in range(nl):
        b = f"transformer_blocks.{i}."
        last = i == nl - 1
        lin(b + "norm1.linear", 6 * d, d)
        lin(b + "norm1_context.linear", (2 if last else 6) * d, d)
        s[b + "attn.norm_q.weight"] = (hd,)
        s[b + "attn.norm_k.weight"] = (hd,)
        for n in ("to_q", "to_k", "to_v", "add_k_proj", "add_v_proj", "add_q_proj"):
            lin(b + "attn." + n, d, d)
        s[b + "attn.norm_add_q.weight"] = (hd,)
        s[b + "attn.norm_add_k.weight"] = (hd,)
        lin(b + "attn.to_out.0", d, d)
        if not last:
            lin(b + "attn.to_add_out", d, d)
        lin(b + "ff.net.0.proj", 4 * d, d)
        lin(b + "ff.net.2", d, 4 * d)
        if not last:
            lin(b + "ff_context.net.0.proj", 4 * d, d)
            lin(b + "ff_context.net.2", d, 4 * d)
    lin("norm_out.linear", 2 * d, d)
    lin("proj_out", p * p * c, d)
    return s


def vae_decoder_shapes(cfg: dict) -> Shapes:
    """post_quant_conv + decoder.* (model/vae.py:691-728,824)."""
    chans = list(reversed(cfg["decoder_block_out_channels"]))
    layers = cfg["decoder_layers_per_block"]
    sp, tp = cfg["decoder_spatial_up_sample"], cfg["decoder_temporal_up_sample"]
    zc = cfg["decoder_in_channels"]
    s: Shapes = OrderedDict()

    def conv(name, co, ci, k):
        s[name + ".conv.weight"] = (co, ci, k, k, k)
        s[name + ".conv.bias"] = (co,)

    def norm(name, ch):
        s[name + ".weight"] = (ch,)
        s[name + ".bias"] = (ch,)

    def resnet(name, ci, co):
        norm(name + ".norm1", ci)
        conv(name + ".conv1", co, ci, 3)
        norm(name + ".norm2", co)
        conv(name + ".conv2", co, co, 3)
        if ci != co:
            conv(name + ".conv_shortcut", co, ci, 1)

    conv("decoder.conv_in", chans[0], zc, 3)
    c0 = chans[0]
    # diffusers Attention registers to_q/k/v/out after group_norm; resnets after attentions
    norm("decoder.mid_block.attentions.0.group_norm", c0)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        s[f"decoder.mid_block.attentions.0.{n}.weight"] = (c0, c0)
        s[f"decoder.mid_block.attentions.0.{n}.bias"] = (c0,)
    resnet("decoder.mid_block.resnets.0", c0, c0)
    resnet("decoder.mid_block.resnets.1", c0, c0)
    prev = c0
    for i, co in enumerate(chans):
        for j in range(layers[i]):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else co, co)
        if sp[i]:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", co * 4, co, 3)
        if tp[i]:
            conv(f"decoder.up_blocks.{i}.temporal_upsamplers.0.conv", co * 2, co, 3)
        prev = co
    norm("decoder.conv_norm_out", chans[-1])
    conv("decoder.conv_out", 3, chans[-1], 3)
    conv("post_quant_conv", zc, zc, 1)
    return s


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def make_weights(shapes: Shapes, seed: int = 1, kind: str = "mmdit") -> Dict[str, torch.Tensor]:
    """Seeded fp32 CPU tensors; scale rules keep activations O(1) through the stack."""
    out: Dict[str, torch.Tensor] = OrderedDict()
    for name, shape in shapes.items():
        g = _gen(name, seed)
        if name.endswith(".bias"):
            if ".norm" in name or "group_norm" in name or "conv_norm_out" in name:
                t = 0.05 * torch.randn(shape, generator=g)
            else:
                t = 0.02 * torch.randn(shape, generator=g)
        elif len(shape) == 1:  # norm scales
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = math.prod(shape[1:])
            fan_out = shape[0]
            if kind == "mmdit":
                if "norm1" in name or "norm_out" in name or name.startswith("proj_out") or \
                        "time_text_embed" in name or name.startswith("context_embedder"):
                    std = 0.02  # conditioning / zero-init layers (mmdit.py:1269-1286) re-drawn
                else:
                    std = math.sqrt(2.0 / (fan_in + fan_out))  # xavier, mmdit.py:1258
            else:
                std = 1.0 / math.sqrt(fan_in)
            t = std * torch.randn(shape, generator=g)
        out[name] = t.float()
    return out


def mmdit_weights(cfg: dict | None = None, seed: int = 1):
    cfg = dict(MMDIT_DEFAULT, **(cfg or {}))
    return cfg, make_weights(mmdit_shapes(cfg), seed, "mmdit")


def vae_encoder_shapes(cfg: dict) -> Shapes:
    """encoder.* + quant_conv (model/vae.py:630-665,823)."""
    chans = list(cfg["encoder_block_out_channels"])
    layers = cfg["encoder_layers_per_block"]
    sp, tp = cfg["encoder_spatial_down_sample"], cfg["encoder_temporal_down_sample"]
    zc = cfg["encoder_out_channels"]
    s: Shapes = OrderedDict()

    def conv(name, co, ci, k):
        s[name + ".conv.weight"] = (co, ci, k, k, k)
        s[name + ".conv.bias"] = (co,)

    def norm(name, ch):
        s[name + ".weight"] = (ch,)
        s[name + ".bias"] = (ch,)

    def resnet(name, ci, co):
        norm(name + ".norm1", ci)
        conv(name + ".conv1", co, ci, 3)
        norm(name + ".norm2", co)
        conv(name + ".conv2", co, co, 3)
        if ci != co:
            conv(name + ".conv_shortcut", co, ci, 1)

    conv("encoder.conv_in", chans[0], cfg.get("encoder_in_channels", 3), 3)
    prev = chans[0]
    for i, co in enumerate(chans):
        for j in range(layers[i]):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", prev if j == 0 else co, co)
        if sp[i]:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", co, co, 3)
        if tp[i]:
            conv(f"encoder.down_blocks.{i}.temporal_downsamplers.0.conv", co, co, 3)
        prev = co
    c0 = chans[-1]
    norm("encoder.mid_block.attentions.0.group_norm", c0)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        s[f"encoder.mid_block.attentions.0.{n}.weight"] = (c0, c0)
        s[f"encoder.mid_block.attentions.0.{n}.bias"] = (c0,)
    resnet("encoder.mid_block.resnets.0", c0, c0)
    resnet("encoder.mid_block.resnets.1", c0, c0)
    norm("encoder.conv_norm_out", c0)
    conv("encoder.conv_out", 2 * zc, c0, 3)
    conv("quant_conv", 2 * zc, 2 * zc, 1)
    return s


def vae_weights(cfg: dict | None = None, seed: int = 2, encoder: bool = False):
    """Decoder weights (+ the encoder's, drawn from an independent stream so that adding them does
    not change the decoder values existing fixtures were made with)."""
    cfg = dict(VAE_DEFAULT, **(cfg or {}))
    W = make_weights(vae_decoder_shapes(cfg), seed, "vae")
    if encoder:
        W.update(make_weights(vae_encoder_shapes(cfg), seed + 1000, "vae"))
    return cfg, W
