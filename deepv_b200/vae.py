"""Host-side mirror of the reference VAE decode (`model/vae.py:CausalVideoVAE.decode`).

`B200VAE` is what `InferencePipeline._create_models` returns in place of `CausalVideoVAE`
for the decode half of the hot path: `.decode(z, temporal_chunk=True, window_size=1,
tile_sample_min_size=256).sample` (pipeline.py:713), `.enable_tiling()`, `.eval()`, `.to()`,
`.device`, `.dtype`.  `encode` is outside the hot path and is delegated to an
optional reference encoder object supplied by the caller.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import TensorRef, VAEConfig, check


def pack_decoder_weights(sd: Dict[str, torch.Tensor], cfg: dict, device) -> Dict[str, torch.Tensor]:
    """Re-layout the decoder state dict (names of oracle/weights.py == reference state_dict):
    conv weights [Cout,Cin,kt,kh,kw] -> bf16 [Cout_rows][taps*Cin_pad] (tap-major, Cin padded to a
    multiple of 64); pixel-shuffle convs get their rows permuted (c p1 p2)->(p1 p2 c), frame-
    interleave convs (c p)->(p c) so that the GEMM epilogue stores contiguous channel runs
    (vae.py:382,407); biases fp32 in the same order, padded to a multiple of 32."""
    out: Dict[str, torch.Tensor] = {}

    def pad_to(t, n, dim=0):
        if t.shape[dim] >= n:
            return t
        pad = [0, 0] * (t.dim() - dim - 1) + [0, n - t.shape[dim]]
        return F.pad(t, pad)

    def conv(name, perm=None, rows=None):
        w = sd[name + ".weight"].float()
        b = sd[name + ".bias"].float()
        co, ci = w.shape[0], w.shape[1]
        if perm == "hw":      # rows (c, p1, p2) -> (p1, p2, c)
            idx = torch.arange(co).view(co // 4, 4).t().reshape(-1)
            w, b = w[idx], b[idx]
        elif perm == "t":     # rows (c, p) -> (p, c)
            idx = torch.arange(co).view(co // 2, 2).t().reshape(-1)
            w, b = w[idx], b[idx]
        cip = (ci + 63) // 64 * 64
        w = pad_to(w.permute(0, 2, 3, 4, 1), cip, dim=4).reshape(co, -1)
        if rows is not None:
            w = pad_to(w, rows, dim=0)
        b = pad_to(b, max(32, (b.shape[0] + 31) // 32 * 32, rows or 0))
        out[name + ".weight"] = w.to(device=device, dtype=torch.bfloat16).contiguous()
        out[name + ".bias"] = b.to(device=device, dtype=torch.float32).contiguous()

    def vec(name):
        out[name] = sd[name].to(device=device, dtype=torch.float32).contiguous()

    chans = list(reversed(cfg["decoder_block_out_channels"]))
    conv("post_quant_conv.conv", rows=64)
    conv("decoder.conv_in.conv")
    a = "decoder.mid_block.attentions.0."
    vec(a + "group_norm.weight")
    vec(a + "group_norm.bias")
    out[a + "to_qkv.weight"] = torch.cat([sd[a + "to_q.weight"], sd[a + "to_k.weight"], sd[a + "to_v.weight"]]) \
        .to(device=device, dtype=torch.bfloat16).contiguous()
    out[a + "to_qkv.bias"] = torch.cat([sd[a + "to_q.bias"], sd[a + "to_k.bias"], sd[a + "to_v.bias"]]) \
        .to(device=device, dtype=torch.float32).contiguous()
    out[a + "to_out.0.weight"] = sd[a + "to_out.0.weight"].to(device=device, dtype=torch.bfloat16).contiguous()
    vec(a + "to_out.0.bias")

    def resnet(name):
        for n in ("norm1", "norm2"):
            vec(f"{name}.{n}.weight")
            vec(f"{name}.{n}.bias")
        conv(name + ".conv1.conv")
        conv(name + ".conv2.conv")
        if name + ".conv_shortcut.conv.weight" in sd:
            conv(name + ".conv_shortcut.conv")

    resnet("decoder.mid_block.resnets.0")
    resnet("decoder.mid_block.resnets.1")
    for i in range(len(chans)):
        for j in range(cfg["decoder_layers_per_block"][i]):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}")
        if cfg["decoder_spatial_up_sample"][i]:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv.conv", perm="hw")
        if cfg["decoder_temporal_up_sample"][i]:
            conv(f"decoder.up_blocks.{i}.temporal_upsamplers.0.conv.conv", perm="t")
    vec("decoder.conv_norm_out.weight")
    vec("decoder.conv_norm_out.bias")
    conv("decoder.conv_out.conv", rows=16)
    # conv_out as per-tap partial products (csrc/vae.cu): rows (tap, c padded to 4), K = Cin
    w = sd["decoder.conv_out.conv.weight"].float()            # [co, ci, kt, kh, kw]
    co, ci = w.shape[0], w.shape[1]
    if co > 4 or tuple(w.shape[2:]) != (3, 3, 3):
        raise _lib.DeepVError(f"conv_out {tuple(w.shape)}: the tap-gather path needs <= 4 output channels, 3x3x3")
    taps = F.pad(w.permute(2, 3, 4, 0, 1).reshape(27, co, ci), (0, (ci + 63) // 64 * 64 - ci, 0, 4 - co))
    out["decoder.conv_out.conv.weight_taps"] = taps.reshape(27 * 4, -1).to(device=device, dtype=torch.bfloat16).contiguous()
    return out


def pack_encoder_weights(sd: Dict[str, torch.Tensor], cfg: dict, device) -> Dict[str, torch.Tensor]:
    """The encoder's state dict (encoder.* + quant_conv, reference names) in the layout of
    pack_decoder_weights: conv weights bf16 [Cout_rows][27 * Cin_pad] (conv_in's 3 input channels padded
    to 64; conv_out and quant_conv padded to 64 rows / 64 input channels so that they chain), norm
    affines and biases fp32, the mid-block attention's q|k|v fused."""
    out: Dict[str, torch.Tensor] = {}

    def pad_to(t, n, dim=0):
        if t.shape[dim] >= n:
            return t
        pad = [0, 0] * (t.dim() - dim - 1) + [0, n - t.shape[dim]]
        return F.pad(t, pad)

    def conv(name, rows=None):
        w = sd[name + ".weight"].float()
        b = sd[name + ".bias"].float()
        co, ci = w.shape[0], w.shape[1]
        cip = (ci + 63) // 64 * 64
        w = pad_to(w.permute(0, 2, 3, 4, 1), cip, dim=4).reshape(co, -1)
        if rows is not None:
            w = pad_to(w, rows, dim=0)
        b = pad_to(b, max(32, (b.shape[0] + 31) // 32 * 32, rows or 0))
        out[name + ".weight"] = w.to(device=device, dtype=torch.bfloat16).contiguous()
        out[name + ".bias"] = b.to(device=device, dtype=torch.float32).contiguous()

    def vec(name):
        out[name] = sd[name].to(device=device, dtype=torch.float32).contiguous()

    def resnet(name):
        for n in ("norm1", "norm2"):
            vec(f"{name}.{n}.weight")
            vec(f"{name}.{n}.bias")
        conv(name + ".conv1.conv")
        conv(name + ".conv2.conv")
        if name + ".conv_shortcut.conv.weight" in sd:
            conv(name + ".conv_shortcut.conv")

    chans = list(cfg["encoder_block_out_channels"])
    conv("encoder.conv_in.conv")
    for i in range(len(chans)):
        for j in range(cfg["encoder_layers_per_block"][i]):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}")
        if cfg["encoder_spatial_down_sample"][i]:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv.conv")
        if cfg["encoder_temporal_down_sample"][i]:
            conv(f"encoder.down_blocks.{i}.temporal_downsamplers.0.conv.conv")
    a = "encoder.mid_block.attentions.0."
    vec(a + "group_norm.weight")
    vec(a + "group_norm.bias")
    out[a + "to_qkv.weight"] = torch.cat([sd[a + "to_q.weight"], sd[a + "to_k.weight"], sd[a + "to_v.weight"]]) \
        .to(device=device, dtype=torch.bfloat16).contiguous()
    out[a + "to_qkv.bias"] = torch.cat([sd[a + "to_q.bias"], sd[a + "to_k.bias"], sd[a + "to_v.bias"]]) \
        .to(device=device, dtype=torch.float32).contiguous()
    out[a + "to_out.0.weight"] = sd[a + "to_out.0.weight"].to(device=device, dtype=torch.bfloat16).contiguous()
    vec(a + "to_out.0.bias")
    resnet("encoder.mid_block.resnets.0")
    resnet("encoder.mid_block.resnets.1")
    vec("encoder.conv_norm_out.weight")
    vec("encoder.conv_norm_out.bias")
    conv("encoder.conv_out.conv", rows=64)
    conv("quant_conv.conv", rows=64)
    return out


class B200LatentDist:
    """`DiagonalGaussianDistribution` of the reference (vae.py:599-628) over device moments."""

    def __init__(self, lib, moments: torch.Tensor, dtype):
        self._lib, self.parameters, self._dtype = lib, moments, dtype
        self.mean, lv = torch.chunk(moments, 2, dim=1)
        self.logvar = torch.clamp(lv, -30.0, 20.0)

    def sample(self, generator=None) -> torch.Tensor:
        noise = torch.randn(self.mean.shape, generator=generator, device=self.parameters.device, dtype=torch.float32)
        return self.sample_with_noise(noise)

    def sample_with_noise(self, noise: torch.Tensor) -> torch.Tensor:
        out = torch.empty(self.mean.shape, device=self.parameters.device, dtype=self._dtype)
        noise = noise.to(device=self.parameters.device, dtype=torch.float32).contiguous()
        check(self._lib.dv_gaussian_sample(self.parameters.data_ptr(), noise.data_ptr(), out.data_ptr(),
                                           out.numel(), _lib.dtype_code(self._dtype), _lib.stream_ptr()),
              "dv_gaussian_sample")
        self._keep = noise
        return out

    def mode(self) -> torch.Tensor:
        return self.mean


class B200VAE:
    """Drop-in for reference `CausalVideoVAE`: decode, and encode when the encoder's weights are given."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], config: dict, device="cuda",
                 dtype=torch.bfloat16, reference_encoder=None):
        self.lib = _lib.load()
        self.config = dict(config)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.DeepVError("B200VAE needs a CUDA device (sm_100a); no CPU fallback")
        self.dtype = dtype
        self.use_tiling = False
        self.reference_encoder = reference_encoder
        self._packed = pack_decoder_weights(state_dict, config, self.device)
        self.has_encoder = "encoder.conv_in.conv.weight" in state_dict
        if self.has_encoder:
            self._packed.update(pack_encoder_weights(state_dict, config, self.device))
        self._names = [n.encode() for n in self._packed]
        refs = (TensorRef * len(self._packed))()
        for i, (n, t) in enumerate(self._packed.items()):
            refs[i] = TensorRef(self._names[i], t.data_ptr(), t.numel())
        ch = config["decoder_block_out_channels"]
        c = VAEConfig()
        c.latent_channels = config["decoder_in_channels"]
        c.out_channels = 3
        for i in range(4):
            c.block_channels[i] = ch[i]
            c.layers_per_block[i] = config["decoder_layers_per_block"][i]
            c.spatial_up[i] = int(config["decoder_spatial_up_sample"][i])
            c.temporal_up[i] = int(config["decoder_temporal_up_sample"][i])
        c.norm_groups = config.get("decoder_norm_num_groups", 32)
        if self.has_encoder:
            c.enc_in_channels = config.get("encoder_in_channels", 3)
            for i in range(4):
                c.enc_block_channels[i] = config["encoder_block_out_channels"][i]
                c.enc_layers_per_block[i] = config["encoder_layers_per_block"][i]
                c.enc_spatial_down[i] = int(config["encoder_spatial_down_sample"][i])
                c.enc_temporal_down[i] = int(config["encoder_temporal_down_sample"][i])
        self._handle = C.c_void_p()
        check(self.lib.dv_vae_create(C.byref(c), refs, len(self._packed), C.byref(self._handle)),
              "dv_vae_create")
        self._plans: Dict[tuple, int] = {}

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    def enable_tiling(self, use_tiling: bool = True):
        self.use_tiling = use_tiling

    def _plan(self, T, h, w, tile):
        key = (T, h, w, tile)
        p = self._plans.get(key)
        if p is None:
            hdl = C.c_void_p()
            check(self.lib.dv_vae_plan_create(self._handle, T, h, w, tile, C.byref(hdl)),
                  "dv_vae_plan_create")
            p = hdl.value
            self._plans[key] = p
        return p

    def plan_flops(self, T, h, w, tile=32) -> float:
        return self.lib.dv_vae_plan_flops(self._plan(T, h, w, tile))

    def decode(self, z: torch.Tensor, return_dict: bool = True, is_init_image=True,
               temporal_chunk=False, window_size=2, tile_sample_min_size=256, out_dtype=None, first_frame: int = 0):
        """Reference signature vae.py:885-886.  temporal_chunk / window_size do not change the
        result and are accepted for call compatibility.  `first_frame` (extension): compute
        output frames [first_frame, T_out) only — bit-identical to the full decode there, zeros in front."""
        _lib.require_cuda(z)
        if z.shape[0] != 1:
            outs = [self.decode(z[i:i + 1], True, is_init_image, temporal_chunk, window_size,
                                tile_sample_min_size, out_dtype, first_frame).sample for i in range(z.shape[0])]
            return SimpleNamespace(sample=torch.cat(outs, dim=0))
        if z.dtype not in (torch.float32, torch.bfloat16):
            z = z.float()
        z = z.contiguous()
        _, _, T, h, w = z.shape
        tile = int(tile_sample_min_size / 8)
        if not self.use_tiling:
            tile = max(h, w)
        plan = self._plan(T, h, w, tile)
        od = out_dtype or z.dtype
        out = torch.empty((1, 3, 8 * (T - 1) + 1, 8 * h, 8 * w), device=z.device, dtype=od)
        check(self.lib.dv_vae_plan_set_first_frame(plan, int(first_frame)), "dv_vae_plan_set_first_frame")
        if first_frame:
            out[:, :, :first_frame].zero_()
        check(self.lib.dv_vae_decode(plan, z.data_ptr(), _lib.dtype_code(z.dtype), out.data_ptr(),
                                     _lib.dtype_code(od), _lib.stream_ptr()), "dv_vae_decode")
        self._last = z
        if not return_dict:
            return (out,)
        return SimpleNamespace(sample=out)

    # ---- multi-GPU: tiles x modalities dealt over the ranks (parallel.Shard) -----------------
    def _sharded_plan(self, T, h, w, tile, slot):
        """A plan per modality slot whose tile outputs live in torch buffers (so that
        torch.distributed can broadcast them)."""
        key = (T, h, w, tile, "shard", slot)
        ent = self._plans.get(key)
        if ent is None:
            hdl = C.c_void_p()
            check(self.lib.dv_vae_plan_create(self._handle, T, h, w, tile, C.byref(hdl)), "dv_vae_plan_create")
            rows, cols, tout = C.c_int(), C.c_int(), C.c_int()
            check(self.lib.dv_vae_plan_geometry(hdl, C.byref(rows), C.byref(cols), C.byref(tout)))
            bufs = []
            for i in range(rows.value * cols.value):
                H, W = C.c_int(), C.c_int()
                check(self.lib.dv_vae_plan_tile_info(hdl, i, C.byref(H), C.byref(W)))
                b = torch.empty((tout.value, H.value, W.value, 3), device=self.device, dtype=torch.bfloat16)
                check(self.lib.dv_vae_plan_bind_tile(hdl, i, b.data_ptr()), "dv_vae_plan_bind_tile")
                bufs.append(b)
            self._plans[key] = hdl.value
            self._shard_bufs = getattr(self, "_shard_bufs", {})
            self._shard_bufs[key] = bufs
        return self._plans[key], self._shard_bufs[key]

    def decode_many(self, zs, shard, tile_sample_min_size=256, out_dtype=None, first_frame: int = 0):
        """Decode several latent videos [1,16,T,h,w] of the same shape (rgb, disparity, ...) with
        the (modality, tile) work items dealt over the ranks of `shard`; every rank returns all
        decoded videos.  `first_frame`: as in `decode`."""
        from .parallel import decode_items
        zs = [z.contiguous() if z.dtype in (torch.float32, torch.bfloat16) else z.float().contiguous() for z in zs]
        _lib.require_cuda(*zs)
        _, _, T, h, w = zs[0].shape
        tile = int(tile_sample_min_size / 8) if self.use_tiling else max(h, w)
        plans = [self._sharded_plan(T, h, w, tile, m) for m in range(len(zs))]
        n_tiles = len(plans[0][1])
        items = decode_items(len(zs), n_tiles)
        for pl, _ in plans:
            check(self.lib.dv_vae_plan_set_first_frame(pl, int(first_frame)), "dv_vae_plan_set_first_frame")
        for i in shard.my_items(len(items)):
            m, t = items[i]
            check(self.lib.dv_vae_decode_tiles(plans[m][0], zs[m].data_ptr(), _lib.dtype_code(zs[m].dtype),
                                               1 << t, _lib.stream_ptr()), "dv_vae_decode_tiles")
        shard.exchange_tiles([plans[m][1][t] for (m, t) in items])
        outs = []
        for m, z in enumerate(zs):
            od = out_dtype or z.dtype
            out = torch.empty((1, 3, 8 * (T - 1) + 1, 8 * h, 8 * w), device=z.device, dtype=od)
            if first_frame:
                out[:, :, :first_frame].zero_()
            check(self.lib.dv_vae_blend(plans[m][0], out.data_ptr(), _lib.dtype_code(od), _lib.stream_ptr()),
                  "dv_vae_blend")
            outs.append(out)
        self._last = zs
        return outs

    def encode(self, x, return_dict: bool = True, is_init_image=True, temporal_chunk=False, window_size=16,
               tile_sample_min_size=256):
        """Reference signature vae.py:844-847; the rollout calls `vae.encode(x).latent_dist.sample()`
        (pipeline.py:250-251,569,574), i.e. the tiled, un-chunked encode.  x: [1,3,T,H,W] on the GPU."""
        if not self.has_encoder:
            if self.reference_encoder is None:
                raise _lib.DeepVError("B200VAE.encode: no encoder weights were given (and no reference_encoder= "
                                      "to delegate to)")
            return self.reference_encoder.encode(x, return_dict, is_init_image, temporal_chunk, window_size,
                                                 tile_sample_min_size)
        if temporal_chunk:
            raise _lib.DeepVError("B200VAE.encode: temporal_chunk=True (window cache) is not built; the rollout "
                                  "never uses it for encoding")
        _lib.require_cuda(x)
        if x.shape[0] != 1:
            parts = [self.encode(x[i:i + 1], True, is_init_image, False, window_size, tile_sample_min_size)
                     .latent_dist.parameters for i in range(x.shape[0])]
            return SimpleNamespace(latent_dist=B200LatentDist(self.lib, torch.cat(parts, dim=0), self.dtype))
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        x = x.contiguous()
        _, _, T, H, W = x.shape
        tile = int(tile_sample_min_size) if self.use_tiling else max(H, W)
        key = ("enc", T, H, W, tile)
        plan = self._enc_plans.get(key) if hasattr(self, "_enc_plans") else None
        if plan is None:
            if not hasattr(self, "_enc_plans"):
                self._enc_plans = {}
            hdl = C.c_void_p()
            check(self.lib.dv_vae_enc_plan_create(self._handle, T, H, W, tile, C.byref(hdl)), "dv_vae_enc_plan_create")
            plan = self._enc_plans[key] = hdl.value
        t, h, w = C.c_int(), C.c_int(), C.c_int()
        check(self.lib.dv_vae_enc_plan_latent_dims(plan, C.byref(t), C.byref(h), C.byref(w)))
        zc = self.config["encoder_out_channels"]
        moments = torch.empty((1, 2 * zc, t.value, h.value, w.value), device=x.device, dtype=torch.float32)
        check(self.lib.dv_vae_encode(plan, x.data_ptr(), _lib.dtype_code(x.dtype), moments.data_ptr(), None, None, 0,
                                     _lib.stream_ptr()), "dv_vae_encode")
        self._last_enc = x
        dist_ = B200LatentDist(self.lib, moments, self.dtype)
        return SimpleNamespace(latent_dist=dist_) if return_dict else (dist_,)

    def enc_plan_flops(self, T, H, W, tile=256) -> float:
        hdl = C.c_void_p()
        check(self.lib.dv_vae_enc_plan_create(self._handle, T, H, W, tile, C.byref(hdl)), "dv_vae_enc_plan_create")
        f = self.lib.dv_vae_enc_plan_flops(hdl)
        self.lib.dv_vae_enc_plan_destroy(hdl)
        return f

    def close(self):
        for p in self._plans.values():
            self.lib.dv_vae_plan_destroy(p)
        self._plans.clear()
        for p in getattr(self, "_enc_plans", {}).values():
            self.lib.dv_vae_enc_plan_destroy(p)
        if hasattr(self, "_enc_plans"):
            self._enc_plans.clear()
        if self._handle:
            self.lib.dv_vae_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
