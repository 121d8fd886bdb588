"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of the reference denoiser forward.

Restates /root/reference/model/mmdit.py MMDiT.forward (:1467-1530) for the only runnable
configuration (sincos spatial pos-emb, temporal RoPE, temporal-causal joint
attention, one stage, caption_projection_dim == inner_dim).  Functional style over a flat
state-dict (names from oracle/weights.py == the reference's state_dict keys).

Parity pinning: tests/test_oracle_vs_reference.py checks this file against the real reference
(imported through oracle/_shim.py) in this container; tests/golden/*.pt hold outputs of the
real reference that this file (and the CUDA path) must reproduce on the GPU box, where
/root/reference does not exist.  The reference has no golden vectors of its own.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference arm may
import this module.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# positional tables
# --------------------------------------------------------------------------------------
def sincos_1d(dim: int, pos: np.ndarray) -> np.ndarray:
    """mmdit.py:624-642: [sin(pos w) | cos(pos w)], w_d = 10000^(-d/(dim/2)), fp64."""
    omega = 1.0 / 10000 ** (np.arange(dim // 2, dtype=np.float64) / (dim / 2.0))
    ang = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(ang), np.cos(ang)], axis=1)


def sincos_2d_table(dim: int, size: int, base_size: int) -> Tensor:
    """mmdit.py:590-621 with interpolation_scale 1: rows indexed y*size + x;
    first half of the channels encodes x ("w goes first"), second half y."""
    coord = np.arange(size, dtype=np.float32) / (size / base_size)
    gx, gy = np.meshgrid(coord, coord)  # gx[y, x] = coord[x]
    emb = np.concatenate([sincos_1d(dim // 2, gx), sincos_1d(dim // 2, gy)], axis=1)
    return torch.from_numpy(emb).float()  # [size*size, dim]


def cropped_pos(table: Tensor, size: int, h: int, w: int, ori_h: int, ori_w: int) -> Tensor:
    """mmdit.py:841-880 with interp_condition_pos=True (token-grid units)."""
    assert ori_h >= h and ori_w >= w  # mmdit.py:851-852
    top, left = (size - ori_h) // 2, (size - ori_w) // 2
    p = table.view(size, size, -1)[top:top + ori_h, left:left + ori_w]
    if (ori_h, ori_w) != (h, w):
        p = F.interpolate(p.permute(2, 0, 1)[None], size=(h, w), mode="bilinear")[0].permute(1, 2, 0)
    return p.reshape(h * w, -1)


def timestep_features(t: Tensor, dim: int = 256) -> Tensor:
    """mmdit.py:645-683 with flip_sin_to_cos=True, downscale_freq_shift=0: [cos | sin]."""
    half = dim // 2
    expo = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    ang = t[:, None].float() * torch.exp(expo)[None]
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)


def rope_table(frame_ids: Tensor, head_dim: int) -> Tensor:
    """mmdit.py:999-1012: (cos, sin) of frame * 10000^(-2i/hd), fp64 -> fp32.  [L, hd/2, 2]"""
    omega = 1.0 / (10000 ** (torch.arange(0, head_dim, 2, dtype=torch.float64, device=frame_ids.device) / head_dim))
    ang = frame_ids.double()[:, None] * omega[None]
    return torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).float()


def apply_rope(x: Tensor, cs: Tensor) -> Tensor:
    """mmdit.py:131-136: interleaved pairs (2i, 2i+1) rotated by the angle of pair i.
    x: [B, L, H, hd]; cs: [L, hd/2, 2]"""
    xr = x.float().reshape(*x.shape[:-1], -1, 2)
    a, b = xr[..., 0], xr[..., 1]
    c, s = cs[None, :, None, :, 0], cs[None, :, None, :, 1]
    return torch.stack([c * a - s * b, s * a + c * b], dim=-1).reshape(x.shape)


# --------------------------------------------------------------------------------------
# small layers
# --------------------------------------------------------------------------------------
def linear(W: Dict[str, Tensor], name: str, x: Tensor) -> Tensor:
    return F.linear(x, W[name + ".weight"], W.get(name + ".bias"))


def layer_norm(x: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), None, None, 1e-6)  # mmdit.py:375,532,1237


def rms_norm(x: Tensor, w: Tensor, eps: float = 1e-5) -> Tensor:
    """mmdit.py:451-464 (eps 1e-5 from JointAttention, :195)."""
    var = x.float().pow(2).mean(-1, keepdim=True)
    return x * torch.rsqrt(var + eps) * w


def patch_tokens(W, prefix: str, lat: Tensor, patch: int) -> Tensor:
    """Conv2d(k=stride=patch) patchify, mmdit.py:892-899.  lat [B,C,t,h,w] -> [B, t*gh*gw, D]"""
    B, C, t, h, w = lat.shape
    y = F.conv2d(lat.permute(0, 2, 1, 3, 4).reshape(B * t, C, h, w), W[prefix + ".weight"],
                 W[prefix + ".bias"], stride=patch)
    return y.flatten(2).transpose(1, 2).reshape(B, t * (h // patch) * (w // patch), -1)


# --------------------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------------------
def mmdit_forward(W: Dict[str, Tensor], cfg: dict, clips: Sequence[Tensor], timestep: Tensor,
                  enc: Tensor, enc_mask: Tensor, pooled: Tensor,
                  history: Optional[Tensor] = None, history_mask: Optional[Tensor] = None,
                  history_downsample_ratio: Optional[int] = None,
                  pos_table: Optional[Tensor] = None) -> Tensor:
    """clips: oldest first, the last one is the noisy clip; returns [B, C, t_last, h, w]."""
    H, hd = cfg["num_attention_heads"], cfg["attention_head_dim"]
    D, P, NL = H * hd, cfg["patch_size"], cfg["num_layers"]
    S = cfg["pos_embed_max_size"]
    B = clips[-1].shape[0]
    dev = clips[-1].device   # the checker may run on any device (CPU here, CUDA fp32 for full-size cases)
    if pos_table is None:
        pos_table = sincos_2d_table(D, S, cfg["sample_size"] // P)
    pos_table = pos_table.to(dev)

    # conditioning vector (mmdit.py:747-753)
    temb = linear(W, "time_text_embed.timestep_embedder.linear_2",
                  F.silu(linear(W, "time_text_embed.timestep_embedder.linear_1",
                                timestep_features(timestep))))
    temb = temb + linear(W, "time_text_embed.text_embedder.linear_2",
                         F.silu(linear(W, "time_text_embed.text_embedder.linear_1", pooled.float())))
    semb = F.silu(temb)

    # context stream = [history tokens | projected text] (mmdit.py:1480-1485, :977-996)
    ctx = linear(W, "context_embedder", enc.float())
    ctx_mask = enc_mask
    if history is not None:
        r = history_downsample_ratio
        hh, hw = history.shape[-2] // r, history.shape[-1] // r
        hist = F.interpolate(history[:, :, 0].float(), size=(hh, hw), mode="bilinear")[:, :, None]
        htok = patch_tokens(W, "pos_embed.proj_history", hist, P)
        htok = htok + cropped_pos(pos_table, S, hh // P, hw // P, hh // P, hw // P)[None]
        ctx = torch.cat([htok, ctx], dim=1)
        ctx_mask = torch.cat([history_mask.to(enc_mask.dtype), enc_mask], dim=1)
    Lc = ctx.shape[1]

    # video stream: patchify every clip, add its positional crop, frame ids (mmdit.py:944-975,1336-1356)
    gh_n, gw_n = clips[-1].shape[-2] // P, clips[-1].shape[-1] // P
    toks, frames, f0 = [], [], 0
    for clip in clips:
        _, _, t, h, w = clip.shape
        gh, gw = h // P, w // P
        pos = cropped_pos(pos_table, S, gh, gw, gh_n, gw_n)
        tk = patch_tokens(W, "pos_embed.proj", clip.float(), P).view(B, t, gh * gw, D) + pos[None, None]
        toks.append(tk.reshape(B, t * gh * gw, D))
        frames.append(torch.arange(f0, f0 + t, device=dev).repeat_interleave(gh * gw))
        f0 += t
    x = torch.cat(toks, dim=1)
    Lv = x.shape[1]
    frame_ids = torch.cat([torch.zeros(Lc, dtype=torch.long, device=dev), torch.cat(frames)])
    cs = rope_table(frame_ids, hd)

    # attention mask: same sample id (0 = padded) AND frame(q) >= frame(k) (mmdit.py:1414-1434)
    sid = torch.arange(1, B + 1, device=dev)[:, None].expand(B, Lc + Lv).clone()
    sid[:, :Lc][ctx_mask == 0] = 0
    mask = (sid[:, :, None] == sid[:, None, :]) & (frame_ids[None, :, None] >= frame_ids[None, None, :])
    mask = mask[:, None]

    def heads(t: Tensor) -> Tensor:
        return t.view(t.shape[0], t.shape[1], H, hd)

    for i in range(NL):
        b = f"transformer_blocks.{i}."
        last = i == NL - 1
        mx = linear(W, b + "norm1.linear", semb).chunk(6, dim=1)  # shift, scale, gate | shift, scale, gate
        xn = layer_norm(x) * (1 + mx[1][:, None]) + mx[0][:, None]
        mc = linear(W, b + "norm1_context.linear", semb)
        if last:  # AdaLayerNormContinuous: (scale, shift), mmdit.py:513-514
            sc, sh = mc.chunk(2, dim=1)
            cn = layer_norm(ctx) * (1 + sc[:, None]) + sh[:, None]
        else:
            mc = mc.chunk(6, dim=1)
            cn = layer_norm(ctx) * (1 + mc[1][:, None]) + mc[0][:, None]
        a = b + "attn."
        q = rms_norm(heads(linear(W, a + "to_q", xn)), W[a + "norm_q.weight"])
        k = rms_norm(heads(linear(W, a + "to_k", xn)), W[a + "norm_k.weight"])
        v = heads(linear(W, a + "to_v", xn))
        qc = rms_norm(heads(linear(W, a + "add_q_proj", cn)), W[a + "norm_add_q.weight"])
        kc = rms_norm(heads(linear(W, a + "add_k_proj", cn)), W[a + "norm_add_k.weight"])
        vc = heads(linear(W, a + "add_v_proj", cn))
        qj = apply_rope(torch.cat([qc, q], dim=1), cs).transpose(1, 2)
        kj = apply_rope(torch.cat([kc, k], dim=1), cs).transpose(1, 2)
        vj = torch.cat([vc, v], dim=1).transpose(1, 2)
        o = F.scaled_dot_product_attention(qj, kj, vj, attn_mask=mask).transpose(1, 2).flatten(2)
        oc, ox = o[:, :Lc], o[:, Lc:]
        x = x + mx[2][:, None] * linear(W, a + "to_out.0", ox)
        xn = layer_norm(x) * (1 + mx[4][:, None]) + mx[3][:, None]
        ff = linear(W, b + "ff.net.2", F.gelu(linear(W, b + "ff.net.0.proj", xn), approximate="tanh"))
        x = x + mx[5][:, None] * ff
        if not last:
            ctx = ctx + mc[2][:, None] * linear(W, a + "to_add_out", oc)
            cn = layer_norm(ctx) * (1 + mc[4][:, None]) + mc[3][:, None]
            ffc = linear(W, b + "ff_context.net.2",
                         F.gelu(linear(W, b + "ff_context.net.0.proj", cn), approximate="tanh"))
            ctx = ctx + mc[5][:, None] * ffc

    sc, sh = linear(W, "norm_out.linear", semb).chunk(2, dim=1)  # (scale, shift), mmdit.py:504
    y = linear(W, "proj_out", layer_norm(x) * (1 + sc[:, None]) + sh[:, None])
    # keep the noisy clip, unpatchify (mmdit.py:1450-1457)
    t = clips[-1].shape[2]
    C = cfg["in_channels"]
    y = y[:, -t * gh_n * gw_n:].reshape(B, t, gh_n, gw_n, P, P, C)
    return y.permute(0, 6, 1, 2, 4, 3, 5).reshape(B, C, t, gh_n * P, gw_n * P)
