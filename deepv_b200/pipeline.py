"""Rollout glue around the B200 denoiser / scheduler / VAE: the host side of the hot loop.

Mirrors the call surface of reference `InferencePipeline` for the hot path only:
  generate_one_unit   pipeline.py:439-524   stage x step loop, CFG, stage transition
  decode_latent       pipeline.py:703-725   latent un-normalisation + tiled VAE decode
  pyramid_conditions  pipeline.py:621-658   condition clip lists per stage (host glue)
`prompt_type text|action` and `add_depth` (run.py:374-382) only select which prompt embeddings
are fed and whether the disparity decode is saved; both branches run the same kernels here.
The reference's control plane (CLI, image IO, pose feedback, text encoders) is out of scope
 and can keep calling these methods unchanged — see INTEGRATION.md.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import check
from .mmdit import B200MMDiT
from .scheduler import B200Scheduler
from .vae import B200VAE

VAE_SHIFT, VAE_SCALE = 0.1490, 1 / 1.8415               # pipeline.py:194-195
VAE_VIDEO_SHIFT, VAE_VIDEO_SCALE = -0.2343, 1 / 3.0986   # pipeline.py:196-197

DEFAULT_CFG = dict(stages=[1, 2, 4], frame_per_unit=1, max_temporal_length=8, vae_downsample=8,
                   raymap_dim=6, history_guidance_scale=6.0, history_downsample_ratio=2,
                   num_inference_steps=5)  # run.py:16-50


class B200Pipeline:
    def __init__(self, dit: B200MMDiT, vae: Optional[B200VAE], scheduler: B200Scheduler,
                 model_cfg: Optional[dict] = None, device="cuda", torch_dtype=torch.bfloat16,
                 guidance_scale: float = 4.0, video_guidance_scale: float = 3.5):
        self.model, self.vae, self.scheduler = dit, vae, scheduler
        self.model_cfg = dict(DEFAULT_CFG, **(model_cfg or {}))
        self.device = torch.device(device)
        self.dtype = torch_dtype
        self._guidance_scale = guidance_scale
        self._video_guidance_scale = video_guidance_scale
        self.lib = _lib.load()

    @property
    def do_classifier_free_guidance(self):
        return self._guidance_scale > 0

    def _timestep_vector(self, value: float, dtype, batch: int) -> torch.Tensor:
        """[batch] fp32 device vector of one timestep, rounded through `dtype` first (pipeline.py:473)."""
        cache = self.__dict__.setdefault("_tvecs", {})
        key = (value, dtype, batch)
        t = cache.get(key)
        if t is None:
            if len(cache) > 4096:
                cache.clear()
            tt = torch.tensor(value, dtype=torch.float64).to(dtype).to(torch.float32)
            t = cache[key] = tt.expand(batch).contiguous().to(self.device)
        return t

    # -- pipeline.py:431-437 on the device ------------------------------------------------------
    def sample_block_noise(self, bs, ch, temp, height, width, generator=None, dtype=None):
        """2x2-block correlated noise, cov (1+g) I - g 11^T; distribution-level equivalent of the
        reference's per-block CPU MultivariateNormal loop."""
        dtype = dtype or self.dtype
        planes = bs * ch * temp
        z = torch.randn((planes, height // 2, width // 2, 4), device=self.device, dtype=torch.float32,
                        generator=generator)
        out = torch.empty((bs, ch, temp, height, width), device=self.device, dtype=dtype)
        check(self.lib.dv_block_noise(z.data_ptr(), out.data_ptr(), planes, height, width,
                                      float(self.scheduler.config.gamma), _lib.dtype_code(dtype),
                                      _lib.stream_ptr()), "dv_block_noise")
        return out

    # -- pipeline.py:439-524 ------------------------------------------------------------------------
    @torch.no_grad()
    def generate_one_unit(self, latents, input_history, past_conditions, prompt_embeds,
                          prompt_attention_mask, pooled_prompt_embeds, num_inference_steps,
                          height=None, width=None, temp=1, device=None, dtype=None, generator=None,
                          is_first_frame: bool = False, block_noise: Optional[Sequence[torch.Tensor]] = None,
                          timestep_dtype=None, shard=None):
        """`shard` (parallel.Shard): run only this rank's CFG branch (batch 1) and all-gather the
        branch predictions before the fused CFG + Euler step."""
        stages = self.model_cfg["stages"]
        n_branch = 3 if input_history is not None else (2 if self.do_classifier_free_guidance else 1)
        w_text = self._guidance_scale if is_first_frame else self._video_guidance_scale
        w_hist = self.model_cfg["history_guidance_scale"]
        latents = latents.contiguous()
        dt = latents.dtype
        history = None
        history_mask = None
        if input_history is not None:
            history = torch.cat([input_history] * 3).to(dt)
            hlen = int((input_history.size(-1) / self.model_cfg["history_downsample_ratio"] / 2) *
                       (input_history.size(-2) / self.model_cfg["history_downsample_ratio"] / 2))
            b = input_history.size(0)
            history_mask = torch.cat([torch.zeros(2 * b, hlen), torch.ones(b, hlen)]).to(self.device)
        enc = prompt_embeds.to(self.device)
        enc_mask = prompt_attention_mask.to(self.device)
        pooled = pooled_prompt_embeds.to(self.device)
        sharded = False
        if shard is not None and shard.active:
            # rollout group = branch groups x Ulysses ranks (parallel.Shard.setup_sp)
            g_nb, sp_world, exchange = shard.layout(n_branch)
            sharded = g_nb > 1
            if hasattr(self.model, "set_sequence_parallel"):
                self.model.set_sequence_parallel(shard.rank % sp_world, sp_world, exchange)
        elif getattr(self.model, "_sp", (0, 1))[1] > 1:
            self.model.set_sequence_parallel(0, 1, None)   # back to the single-rank forward
        if sharded:
            if latents.shape[0] != 1:
                raise _lib.DeepVError("CFG-branch sharding expects one sample per rollout")
            mb = shard.my_branch(n_branch)
            enc, enc_mask, pooled = enc[mb:mb + 1], enc_mask[mb:mb + 1], pooled[mb:mb + 1]
            past_conditions = [[c[mb:mb + 1] for c in st] for st in past_conditions]
            if history is not None:
                history, history_mask = history[mb:mb + 1], history_mask[mb:mb + 1]
        n_local = 1 if sharded else n_branch
        if self.model_cfg.get("no_need_depth", False):
            # pipeline.py:476-478 zeroes channels 16.. (disparity + ray map) of every clip fed to the denoiser;
            # the condition clips never change inside a unit, so they are cleared once, in place like the reference
            for st in past_conditions:
                for c in st:
                    c[:, 16:] = 0
        intermed = []
        for i_s in range(len(stages)):
            self.scheduler.set_timesteps(num_inference_steps[i_s], i_s, device=None)
            timesteps = self.scheduler._timesteps_host
            if i_s > 0:
                b, c, t, h, w = latents.shape
                ori = 1 - self.scheduler.ori_start_sigmas[i_s]
                gamma = self.scheduler.config.gamma
                alpha = 1 / (math.sqrt(1 + (1 / gamma)) * (1 - ori) + ori)
                beta = alpha * (1 - ori) / math.sqrt(gamma)
                if block_noise is not None:
                    noise = block_noise[i_s - 1].to(device=self.device, dtype=dt).contiguous()
                else:
                    noise = self.sample_block_noise(b, c, t, 2 * h, 2 * w, generator, dt)
                up = torch.empty((b, c, t, 2 * h, 2 * w), device=self.device, dtype=dt)
                check(self.lib.dv_stage_renoise(latents.data_ptr(), noise.data_ptr(), up.data_ptr(),
                                                b * c * t, h, w, alpha, beta, _lib.dtype_code(dt),
                                                _lib.stream_ptr()), "dv_stage_renoise")
                latents = up
            for idx in range(len(timesteps)):
                x_in = torch.cat([latents] * n_local) if n_local > 1 else latents
                if self.model_cfg.get("no_need_depth", False):
                    if n_branch == 1:
                        latents[:, 16:] = 0      # without CFG the reference's model input IS `latents` (pipeline.py:468)
                    else:
                        if n_local == 1:
                            x_in = x_in.clone()  # a branch-sharded rank: the reference cleared a torch.cat copy
                        x_in[:, 16:] = 0
                # pipeline.py:473 casts the timestep to the latent dtype; timestep_dtype=float32
                # keeps it unrounded for comparisons against the fp32 oracle.
                # The 15 values of a unit are uploaded once per (value, dtype, batch) and reused by every unit
                # (the reference builds and copies one per step, a host round trip inside the hot loop)
                tvec = self._timestep_vector(float(timesteps[idx]), timestep_dtype or dt, x_in.shape[0])
                noise_pred = self.model(
                    sample=[list(past_conditions[i_s]) + [x_in]], timestep_ratio=tvec,
                    encoder_hidden_states=enc, encoder_attention_mask=enc_mask,
                    pooled_projections=pooled, history=history,
                    history_downsample_ratio=self.model_cfg["history_downsample_ratio"] if history is not None else None,
                    history_mask=history_mask)[0]
                if sharded:
                    noise_pred = shard.gather_branches(noise_pred, n_branch)
                latents = self.scheduler.cfg_step(noise_pred, latents, n_branch, w_text, w_hist)
            intermed.append(latents)
        return intermed

    # -- pipeline.py:226-240 / 554-557 --------------------------------------------------------------
    def resize_half(self, x: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
        """F.interpolate(x, size=(h//2, w//2), mode='bilinear') (* scale) of [..., h, w] on the GPU,
        bit-equal to ATen (dv_resize_half)."""
        _lib.require_cuda(x)
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        x = x.contiguous()
        h, w = x.shape[-2], x.shape[-1]
        out = torch.empty(x.shape[:-2] + (h // 2, w // 2), device=x.device, dtype=x.dtype)
        check(self.lib.dv_resize_half(x.data_ptr(), out.data_ptr(), x.numel() // (h * w), h, w, float(scale),
                                      _lib.dtype_code(x.dtype), _lib.stream_ptr()), "dv_resize_half")
        return out

    def get_pyramid_latent(self, x: torch.Tensor, stage_num: int) -> List[torch.Tensor]:
        """pipeline.py:226-240: [x/2^stage_num, ..., x/2, x] (per-frame bilinear halvings)."""
        out = [x]
        for _ in range(stage_num):
            x = self.resize_half(x)
            out.append(x)
        return list(reversed(out))

    def noise_pyramid_base(self, latents: torch.Tensor) -> torch.Tensor:
        """pipeline.py:554-557: the stage-0 noise = repeated (bilinear half) * 2 of the full-resolution draw."""
        for _ in range(len(self.model_cfg["stages"]) - 1):
            latents = self.resize_half(latents, 2.0)
        return latents

    # -- pipeline.py:621-658 ------------------------------------------------------------------------
    def pyramid_conditions(self, generated: torch.Tensor, unit_index: int, firstframe_mask: bool,
                           n_branch: int) -> List[List[torch.Tensor]]:
        """Condition clip lists per stage: last frame at the stage's resolution, older frames one
        stage lower each, everything older than that at stage 0; oldest first."""
        nst = len(self.model_cfg["stages"])
        fpu = self.model_cfg["frame_per_unit"]
        pyr = self.get_pyramid_latent(generated, nst - 1)
        rep = (lambda z: torch.cat([z] * n_branch)) if n_branch > 1 else (lambda z: z)
        out = []
        for i_s in range(nst):
            clips = [rep(pyr[i_s][:, :, -fpu:])]
            cur_stage, ptx = i_s, 1
            while ptx < unit_index - firstframe_mask:
                cur_stage = max(cur_stage - 1, 0)
                if cur_stage == 0:
                    break
                ptx += 1
                clips.append(rep(pyr[cur_stage][:, :, -(ptx * fpu):-((ptx - 1) * fpu)]))
            if cur_stage == 0 and ptx < unit_index - firstframe_mask:
                clips.append(rep(pyr[0][:, :, int(firstframe_mask):-(ptx * fpu)]))
            out.append(list(reversed(clips)))
        return out

    # -- pipeline.py:703-725 ------------------------------------------------------------------------
    @staticmethod
    def _unnormalise(latents: torch.Tensor) -> torch.Tensor:
        latents = latents.clone()
        if latents.shape[2] == 1:
            return (latents / VAE_SCALE) + VAE_SHIFT
        latents[:, :, :1] = (latents[:, :, :1] / VAE_SCALE) + VAE_SHIFT
        latents[:, :, 1:] = (latents[:, :, 1:] / VAE_VIDEO_SCALE) + VAE_VIDEO_SHIFT
        return latents

    @torch.no_grad()
    def decode_latents_sharded(self, latent_list, shard, out_dtype=None, first_frame: int = 0):
        """pipeline.py:695-696 (image + disparity decodes) with the tiles x modalities dealt over
        the ranks of `shard`; returns one video per entry of latent_list on every rank."""
        zs = [self._unnormalise(z) for z in latent_list]
        return self.vae.decode_many(zs, shard, tile_sample_min_size=256, out_dtype=out_dtype, first_frame=first_frame)

    @torch.no_grad()
    def decode_latent(self, latents: torch.Tensor, save_memory: bool = True, out_dtype=None, first_frame: int = 0):
        """pipeline.py:703-725.  `first_frame` (extension, default 0 = the reference's behaviour): only frames
        [first_frame, T) are computed (bit-identical there, zeros in front) — `B200Rollout` passes 25 in continuation
        iterations, whose first 25 decoded frames the reference throws away (pipeline.py:327-328)."""
        latents = self._unnormalise(latents)
        if not save_memory:
            raise _lib.DeepVError("decode_latent: only the save_memory=True (256 px tile) branch exists; "
                                  "the reference's 512 px branch crashes (SURVEY.md App. E.2)")
        return self.vae.decode(latents, temporal_chunk=True, window_size=1, tile_sample_min_size=256,
                               out_dtype=out_dtype, first_frame=first_frame).sample
