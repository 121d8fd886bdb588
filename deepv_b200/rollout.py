"""The autoregressive rollout around the hot path, resident on the GPU.

Mirrors reference `InferencePipeline.generate` (pipeline.py:264-424) and `generate_i2v`
(pipeline.py:526-700) with the same arguments' meaning and the same result dictionary, but
  * decoded frames feed the next iteration through `dv_frames_requantise` (the uint8/PIL round trip
    as arithmetic) instead of GPU -> numpy -> PIL -> GPU (pipeline.py:339-344,564-568);
  * disparity post-processing / renormalisation, ray map <-> camera pose conversion and the history
    frame selection run on the device with device-resident scalars and indices: no `.cpu()`,
    `.item()` or `.max()` read-back between iterations (pipeline.py:311-313,346-414,692);
  * prompt embeddings are looked up (action mode) or encoded (text mode) once per distinct prompt and
    kept on the device, where the reference re-encodes the same text every unit and moves the
    embeddings host -> device at every denoising step (pipeline.py:596-603,489-492).
The heavy work is `B200Pipeline.generate_one_unit`, `B200VAE.encode` and `decode_latent`.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional, Sequence

import torch

from . import _lib
from ._lib import check
from .pipeline import B200Pipeline, VAE_SCALE, VAE_SHIFT, VAE_VIDEO_SCALE, VAE_VIDEO_SHIFT

NUM_INPUT_IMAGE, NUM_INPUT_UNIT = 25, 4          # pipeline.py:269-270
_CHECK_REPLICAS = bool(os.environ.get("DV_CHECK_REPLICAS"))   # debug: all-rank equality of every unit's latents


class PromptCache:
    """prompt -> (prompt_embeds, prompt_attention_mask, pooled_prompt_embeds) on the device.

    `text_embeds` is the table the reference loads from `model_cfg['text_embeds_path']`
    (pipeline.py:199; action prompts + 'empty'); `text_encoder(prompt, device)` is the reference's
    SD3TextEncoderWithMask call (pipeline.py:603), used for prompts the table does not hold and
    evaluated once per distinct prompt."""

    def __init__(self, text_embeds: Optional[Dict[str, Dict[str, torch.Tensor]]] = None,
                 text_encoder: Optional[Callable] = None, device="cuda"):
        self.device = torch.device(device)
        self.table = text_embeds or {}
        self.text_encoder = text_encoder
        self._one: Dict[str, tuple] = {}
        self._cat: Dict[tuple, tuple] = {}
        self.encoder_calls = 0

    def get(self, prompt: str, use_table: bool = True):
        key = str(prompt)
        hit = self._one.get(key)
        if hit is not None:
            return hit
        if use_table and key in self.table:
            e = self.table[key]
            enc, mask, pooled = e["prompt_embeds"], e["prompt_attention_mask"], e["pooled_prompt_embeds"]
        elif self.text_encoder is not None:
            self.encoder_calls += 1
            enc, mask, pooled = self.text_encoder(key, self.device)      # pipeline.py:603 return order
        else:
            raise _lib.DeepVError(f"PromptCache: no embedding for prompt {key!r} and no text_encoder")
        hit = (enc.to(self.device), mask.to(self.device), pooled.to(self.device))
        self._one[key] = hit
        return hit

    def branches(self, prompt: str, n_branch: int, use_table: bool = True):
        """[negative, prompt(, prompt)] stacked along the CFG batch (pipeline.py:605-617)."""
        key = (str(prompt), n_branch, use_table)
        hit = self._cat.get(key)
        if hit is None:
            rows = [self.get("empty")] + [self.get(prompt, use_table)] * (n_branch - 1)
            hit = tuple(torch.cat([r[i] for r in rows], dim=0) for i in range(3))
            self._cat[key] = hit
        return hit


class DeviceNoise:
    """Default noise source: every draw comes from a device generator."""

    def __init__(self, pipe: B200Pipeline, generator: Optional[torch.Generator] = None):
        self.pipe, self.generator = pipe, generator

    def randn(self, shape):
        return torch.randn(tuple(shape), device=self.pipe.device, dtype=torch.float32, generator=self.generator)

    def block(self, bs, ch, temp, height, width, gamma):
        return self.pipe.sample_block_noise(bs, ch, temp, height, width, self.generator)


def noise_for(pipe: B200Pipeline, noise, shard):
    """The noise source of a rollout.  Un-sharded: whatever was given, else the device's global generator.
    Sharded (rollout group > 1 rank): latents must stay replicated over the group (CFG branches of different
    ranks are combined, Ulysses ranks exchange q/k/v of what must be the same input), so a default source is
    seeded from ONE seed broadcast over the group, and an unseeded DeviceNoise is refused."""
    if shard is None or not shard.active:
        return noise or DeviceNoise(pipe)
    if noise is None:
        seed = shard.shared_seed(pipe.device)
        return DeviceNoise(pipe, torch.Generator(device=pipe.device).manual_seed(seed))
    if isinstance(noise, DeviceNoise) and noise.generator is None:
        raise _lib.DeepVError("sharded rollout: DeviceNoise needs a generator seeded identically on every rank of the "
                              "group (or pass noise=None to have one seeded from a broadcast seed)")
    return noise


def plan_prompts(prompts: Sequence[str], actual_unit: int):
    """pipeline.py:275-279: pad the prompt list with its last entry; number of iterations."""
    p = [str(x) for x in prompts]
    step = actual_unit - NUM_INPUT_UNIT
    while (len(p) - actual_unit) % step != 0 or len(p) < actual_unit:
        p.append(p[-1])
    return p, (len(p) - actual_unit) // step + 1


class B200Rollout:
    def __init__(self, pipe: B200Pipeline, prompts: PromptCache):
        if pipe.vae is None or not pipe.vae.has_encoder:
            raise _lib.DeepVError("B200Rollout needs a B200VAE built with encoder weights (row f1)")
        self.pipe, self.prompts = pipe, prompts
        self.lib = pipe.lib
        self.device, self.dtype = pipe.device, pipe.dtype
        # continuation iterations decode only the frames they keep (bit-identical there; False = decode all 57 as the
        # reference does before it drops the first 25)
        self.trim_continuation_decode = True
        self.cfg = pipe.model_cfg

    # -- small device helpers ---------------------------------------------------------------------
    def _s(self):
        return _lib.stream_ptr()

    def frames_from_uint8(self, frames_u8: torch.Tensor) -> torch.Tensor:
        """uint8 [n,H,W,3] (the PIL frames) -> ToTensor + Normalize(0.5, 0.5) -> [1,3,n,H,W] (pipeline.py:564-568).
        The 256 possible values are converted once on the host exactly as torchvision does (ATen's CUDA division
        by a scalar multiplies by the reciprocal and is 1 ulp off) and gathered on the device, so a frame that is
        already resident never goes back to the host."""
        lut = getattr(self, "_u8_lut", None)
        if lut is None:
            q = torch.arange(256, dtype=torch.float32).div(255)
            lut = self._u8_lut = ((q - 0.5) / 0.5).to(self.device, self.dtype)
        idx = frames_u8.to(self.device, non_blocking=True).permute(3, 0, 1, 2).to(torch.int64)
        return lut[idx].unsqueeze(0).contiguous()

    def requantise(self, images: torch.Tensor, t0: int, n: int, want_u8: bool = False):
        """Frames [t0, t0+n) of a decoded video as the next iteration's input frames (pipeline.py:339-344,564-568)."""
        _lib.require_cuda(images)
        images = images.contiguous()
        _, _, T, H, W = images.shape
        out = torch.empty((1, 3, n, H, W), device=images.device, dtype=self.dtype)
        u8 = torch.empty((n, H, W, 3), device=images.device, dtype=torch.uint8) if want_u8 else None
        check(self.lib.dv_frames_requantise(images.data_ptr(), _lib.dtype_code(images.dtype), T, H, W, t0, n,
                                            out.data_ptr(), _lib.dtype_code(out.dtype),
                                            u8.data_ptr() if want_u8 else None, self._s()), "dv_frames_requantise")
        return (out, u8) if want_u8 else out

    def disparity_post(self, raw: torch.Tensor, scale: Optional[torch.Tensor]) -> torch.Tensor:
        """pipeline.py:311-313; `scale` is the device scalar of the previous iteration (None = 1)."""
        raw = raw.contiguous()
        _, _, T, H, W = raw.shape
        out = torch.empty((1, 3, T, H, W), device=raw.device, dtype=torch.float32)
        check(self.lib.dv_disparity_post(raw.data_ptr(), _lib.dtype_code(raw.dtype), T, H, W,
                                         scale.data_ptr() if scale is not None else None, out.data_ptr(), self._s()),
              "dv_disparity_post")
        return out

    def disparity_renorm(self, disp: torch.Tensor, t0: int, n: int, scale: torch.Tensor, compute_scale: bool,
                         clamp: bool) -> torch.Tensor:
        """pipeline.py:346-350 (compute_scale) / :399-401 (clamp); `scale` is a 1-element fp32 device tensor."""
        disp = disp.contiguous()
        _, _, T, H, W = disp.shape
        out = torch.empty((1, 3, n, H, W), device=disp.device, dtype=self.dtype)
        check(self.lib.dv_disparity_renorm(disp.data_ptr(), T, H, W, t0, n, scale.data_ptr(), int(compute_scale),
                                           int(clamp), out.data_ptr(), _lib.dtype_code(out.dtype), self._s()),
              "dv_disparity_renorm")
        return out

    def raymap_to_pose(self, latents: torch.Tensor):
        """pipeline.py:688-692: ray-map channels of the generated latents -> (trans3d, trans2d) [1,T,4,4] fp32."""
        latents = latents.contiguous()
        _, C, T, h, w = latents.shape
        ray = self.cfg["raymap_dim"]
        t3 = torch.empty((1, T, 4, 4), device=latents.device, dtype=torch.float32)
        t2 = torch.empty((1, T, 4, 4), device=latents.device, dtype=torch.float32)
        check(self.lib.dv_raymap_to_pose(latents.data_ptr(), _lib.dtype_code(latents.dtype), C, C - ray, T, h, w,
                                         self.cfg["vae_downsample"], t3.data_ptr(), t2.data_ptr(), self._s()),
              "dv_raymap_to_pose")
        return t3, t2

    def camera_raymap(self, trans2d: torch.Tensor, trans3d: torch.Tensor, H: int, W: int, normalise: bool = True):
        """pipeline.py:29-75 (+ the (x - mean) / std of :300-301 / :258-259): [1,n,4,4] x2 -> [1,6,n,H/8,W/8]."""
        k2 = trans2d.reshape(-1, 4, 4).to(torch.float32).contiguous()
        k3 = trans3d.reshape(-1, 4, 4).to(torch.float32).contiguous()
        n, ds = k2.shape[0], self.cfg["vae_downsample"]
        out = torch.empty((1, 6, n, H // ds, W // ds), device=k2.device, dtype=self.dtype)
        check(self.lib.dv_camera_raymap(k2.data_ptr(), k3.data_ptr(), n, H, W, ds, int(normalise), out.data_ptr(),
                                        _lib.dtype_code(out.dtype), self._s()), "dv_camera_raymap")
        return out

    def _encode(self, x: torch.Tensor, noise) -> torch.Tensor:
        return self._encode_many([x], noise, None)[0]

    def _encode_many(self, xs, noise, shard=None) -> List[torch.Tensor]:
        """`vae.encode(x).latent_dist.sample()` for each clip of `xs` (pipeline.py:569,574 / :250-251), the normal
        draws taken in that order.  Under a shard the clips (image, disparity) are dealt over the ranks of the
        rollout group — each is encoded once and its latent broadcast — while EVERY rank takes every draw, so the
        replicated generators stay in step."""
        vae = self.pipe.vae
        zc = vae.config["encoder_out_channels"]
        shapes = [(x.shape[0], zc, (x.shape[2] - 1) // 8 + 1, x.shape[3] // 8, x.shape[4] // 8) for x in xs]
        draws = [noise.randn(list(s)).to(self.device, torch.float32) for s in shapes]
        split = shard is not None and shard.active and len(xs) > 1

        def enc(i):
            return vae.encode(xs[i]).latent_dist.sample_with_noise(draws[i])

        if not split:
            return [enc(i) for i in range(len(xs))]
        import torch.distributed as dist
        outs, works = [], []
        for i in range(len(xs)):
            owner = i % shard.world
            z = enc(i).contiguous() if shard.rank == owner else torch.empty(shapes[i], device=self.device, dtype=self.dtype)
            outs.append(z)
            works.append(dist.broadcast(z, src=shard._global(owner), group=shard.group, async_op=True))
        for w in works:
            w.wait()
        return outs

    @staticmethod
    def _normalise(z: torch.Tensor, first_only: bool = False) -> torch.Tensor:
        z[:, :, :1] = (z[:, :, :1] - VAE_SHIFT) * VAE_SCALE                            # pipeline.py:570
        if not first_only:
            z[:, :, 1:] = (z[:, :, 1:] - VAE_VIDEO_SHIFT) * VAE_VIDEO_SCALE            # pipeline.py:571
        return z

    # -- pipeline.py:526-700 ------------------------------------------------------------------------
    @torch.no_grad()
    def generate_i2v(self, motion_prompt, use_motion_prompt: bool = True, input_image: torch.Tensor = None,
                     input_disparity: Optional[torch.Tensor] = None, input_raymap: Optional[torch.Tensor] = None,
                     input_history: Optional[torch.Tensor] = None, temp: int = 8, num_inference_steps=5,
                     guidance_scale: float = 4.0, video_guidance_scale: float = 3.5, min_guidance_scale: float = 1.1,
                     use_linear_guidance: bool = False, alpha: float = 1.0,
                     noise=None, shard=None, return_latents: bool = False):
        """input_image: the input frames as [1,3,n,H,W] in [-1,1] on the device (`frames_from_uint8` /
        `requantise`); input_disparity [1,3,n,H,W]; input_raymap [1,6,n_lat,h,w] normalised;
        input_history [1,38,1,h,w].  Returns image, disparity (raw decodes), trans3d, trans2d."""
        pipe, cfg = self.pipe, self.cfg
        fpu, nst, ray = cfg["frame_per_unit"], len(cfg["stages"]), cfg["raymap_dim"]
        noise = noise_for(pipe, noise, shard)
        first = input_disparity is None
        if temp % fpu != 0:
            raise _lib.DeepVError("generate_i2v: temp must be a multiple of frame_per_unit")
        steps = [num_inference_steps] * nst if isinstance(num_inference_steps, int) else list(num_inference_steps)
        pipe._guidance_scale, pipe._video_guidance_scale = guidance_scale, video_guidance_scale       # :548-549
        ramp = [max(guidance_scale - alpha * t_, min_guidance_scale) for t_ in range(temp + 1)] if use_linear_guidance else None
        _lib.require_cuda(input_image)
        H, W = input_image.shape[-2], input_image.shape[-1]
        C = pipe.model.in_channels
        ds = cfg["vae_downsample"]

        lat_noise = noise.randn((1, C, temp + int(first), H // ds, W // ds)).to(self.device, self.dtype)  # :551
        lat_noise = pipe.noise_pyramid_base(lat_noise)                                                    # :554-557
        num_units = lat_noise.shape[2] // fpu

        if first:
            z_img = self._normalise(self._encode(input_image.to(self.dtype), noise))                      # :569-571
            z_disp = torch.zeros_like(z_img)
        else:
            zi, zd = self._encode_many([input_image.to(self.dtype), input_disparity.to(self.dtype)], noise, shard)
            z_img, z_disp = self._normalise(zi), self._normalise(zd)                                      # :569-576
        z_ray = torch.zeros_like(z_img[:, :ray, :1]) if input_raymap is None else input_raymap.to(z_img)
        generated = [torch.cat([z_img, z_disp, z_ray], dim=1).to(self.dtype)]                             # :578-582

        n_branch = 3 if input_history is not None else (2 if pipe.do_classifier_free_guidance else 1)
        n_frames = input_image.shape[2]
        start = 1 if first else (n_frames - 1) // 8 + 1                                                   # :587
        gamma = float(pipe.scheduler.config.gamma)
        for unit in range(start, num_units):
            if ramp is not None:                                                                         # :592-594
                pipe._guidance_scale = pipe._video_guidance_scale = ramp[unit]
            enc, mask, pooled = self.prompts.branches(motion_prompt[unit - int(first)], n_branch, use_motion_prompt)
            conds = pipe.pyramid_conditions(torch.cat(generated, dim=2), unit, first, n_branch)
            lat = lat_noise[:, :, unit * fpu:(unit + 1) * fpu].contiguous()
            h0, w0 = lat.shape[-2], lat.shape[-1]
            block = [noise.block(1, C, fpu, h0 * 2 ** s, w0 * 2 ** s, gamma) for s in range(1, nst)]
            outs = pipe.generate_one_unit(lat, input_history, conds, enc, mask, pooled, steps, temp=fpu,
                                          is_first_frame=False, block_noise=block, shard=shard)
            if _CHECK_REPLICAS and shard is not None and shard.active:
                shard.assert_replicated(outs[-1], f"latent of unit {unit}")
            generated.append(outs[-1])
        if first:
            generated = generated[1:]                                                                     # :680-681
        lat = torch.cat(generated, dim=2)
        half = (lat.shape[1] - ray) // 2
        z_image, z_disparity = lat[:, :half].contiguous(), lat[:, half:2 * half].contiguous()             # :685-686
        trans3d, trans2d = self.raymap_to_pose(lat)                                                       # :688-692
        # a continuation iteration keeps frames [25:] of both videos only (pipeline.py:327-328,339,346): decode just those
        ff = NUM_INPUT_IMAGE if (not first and self.trim_continuation_decode) else 0
        if shard is not None and shard.active:
            image, disparity = pipe.decode_latents_sharded([z_image, z_disparity], shard, first_frame=ff)  # :694-695
        else:
            image = pipe.decode_latent(z_image, first_frame=ff)
            disparity = pipe.decode_latent(z_disparity, first_frame=ff)
        if cfg.get("no_need_depth", False):
            disparity = torch.zeros_like(disparity)                                                       # :696-697
        if return_latents:
            return image, disparity, trans3d, trans2d, lat
        return image, disparity, trans3d, trans2d

    # -- pipeline.py:243-262 ------------------------------------------------------------------------
    def history_latent(self, rgb, disparity, raymap_normalised, noise, shard=None) -> torch.Tensor:
        zv, zd = self._encode_many([rgb.to(self.dtype), disparity.to(self.dtype)], noise, shard)
        video = self._normalise(zv, first_only=True)
        disp = self._normalise(zd, first_only=True)
        return torch.cat([video, disp, raymap_normalised.to(video)], dim=1)

    # -- pipeline.py:264-424 ------------------------------------------------------------------------
    @torch.no_grad()
    def generate(self, batch_dict: Dict, noise=None, shard=None, trace: Optional[list] = None,
                 events: Optional[list] = None) -> Dict:
        """batch_dict: 'img' (PIL image, ndarray or uint8 tensor [H,W,3]), 'prompt' (sequence of action keys
        or texts), 'prompt_type' ('action' | 'text').  Returns the reference's result dictionary.
        `events` (optional list) receives per iteration three CUDA events: start, after generate_i2v, after the
        feedback — recorded on the current stream, nothing is synchronised here."""
        cfg = self.cfg
        units = cfg["max_temporal_length"]
        noise = noise_for(self.pipe, noise, shard)
        total, iters = plan_prompts(batch_dict["prompt"], units)
        use_table = batch_dict.get("prompt_type", "action") == "action"
        img = batch_dict["img"]
        if not isinstance(img, torch.Tensor):
            import numpy as np
            img = torch.from_numpy(np.array(img, dtype=np.uint8))
        frames = self.frames_from_uint8(img.unsqueeze(0))
        state = _Feedback(self)
        in_disp = in_ray = in_hist = None
        start_unit = 0
        for it in range(iters):
            motion = total[0:1] + total[start_unit + 1:start_unit + units]                                # :296
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if events is not None else None
            if ev:
                ev[0].record()
            image, disparity, t3, t2 = self.generate_i2v(
                motion, use_table, frames, in_disp, in_ray, in_hist, temp=units,
                num_inference_steps=cfg.get("num_inference_steps", 10), guidance_scale=4.0, video_guidance_scale=3.5,
                use_linear_guidance=False, alpha=1.0, min_guidance_scale=1.1, noise=noise, shard=shard)   # :302-309
            if trace is not None:
                trace.append(dict(motion_prompt=motion, frames=frames, input_disparity=in_disp, input_raymap=in_ray,
                                  input_history=in_hist, images=image, disparity=disparity, trans3d=t3, trans2d=t2))
            start_unit += units - NUM_INPUT_UNIT
            if ev:
                ev[1].record()
            disp = state.absorb(it, image, disparity, t3, t2, motion)
            frames, in_disp, in_ray, in_hist = state.next_inputs(image, disp, noise, shard)
            if ev:
                ev[2].record()
                events.append(ev)
        return {"pred_img": torch.cat(state.images, dim=2), "pred_disparity": torch.cat(state.disparitys, dim=2),
                "motion_prompt_list": state.prompts, "trans3d": torch.cat(state.trans3d, dim=1),
                "trans2d": torch.cat(state.trans2d, dim=1)}


def _signed_sqrt(x):
    return torch.sign(x) * torch.sqrt(x.abs())


def _inv(m):
    return torch.linalg.inv_ex(m).inverse          # no info read-back, so no host synchronisation


class _Feedback:
    """What `generate` carries from one iteration to the next (pipeline.py:282-414), on the device."""

    def __init__(self, ro: B200Rollout):
        self.ro = ro
        self.images: List[torch.Tensor] = []
        self.disparitys: List[torch.Tensor] = []
        self.trans3d: List[torch.Tensor] = []
        self.trans2d: List[torch.Tensor] = []
        self.prompts: List[List[str]] = []
        self.key_images: List[torch.Tensor] = []      # every vae_downsample-th frame of the kept video (:370-371)
        self.key_disps: List[torch.Tensor] = []
        self.n_frames = 0
        self.scale: Optional[torch.Tensor] = None      # device scalar; None = 1.0 (first iteration)
        self.history_index: Optional[torch.Tensor] = None

    def _keep(self, images, disp):
        ds = self.ro.cfg["vae_downsample"]
        off = (-self.n_frames) % ds
        self.images.append(images)
        self.disparitys.append(disp)
        self.key_images.append(images[:, :, off::ds])
        self.key_disps.append(disp[:, :, off::ds])
        self.n_frames += images.shape[2]

    def absorb(self, it, images, disparity_raw, trans3d, trans2d, motion):
        disp = self.ro.disparity_post(disparity_raw, self.scale)                           # :311-313
        if self.scale is not None:
            trans3d = trans3d.clone()
            trans3d[:, :, :3, 3] = trans3d[:, :, :3, 3] * self.scale                       # :314
        if it == 0:
            self._keep(images, disp)
            self.prompts.append(list(motion))
            self.trans3d.append(trans3d)
            self.trans2d.append(trans2d)
        else:
            self._keep(images[:, :, NUM_INPUT_IMAGE:], disp[:, :, NUM_INPUT_IMAGE:])        # :327-328
            self.prompts.append(list(motion[NUM_INPUT_UNIT:]))
            pre = self.trans3d[-1][:, -NUM_INPUT_UNIT]
            trans3d = torch.matmul(pre.unsqueeze(1), trans3d)                               # :330-332
            self.trans3d.append(trans3d[:, NUM_INPUT_UNIT:])
            self.trans2d.append(trans2d[:, NUM_INPUT_UNIT:])
        return disp

    def next_inputs(self, images, disp, noise, shard=None):
        ro = self.ro
        T, H, W = images.shape[2], images.shape[3], images.shape[4]
        t0 = T - NUM_INPUT_IMAGE
        frames = ro.requantise(images, t0, NUM_INPUT_IMAGE)                                 # :339-344
        if ro.cfg.get("no_need_depth", False):                                              # :345-350 skipped
            scale = self.scale if self.scale is not None else torch.ones(1, device=images.device, dtype=torch.float32)
            in_disp = disp[:, :, t0:].to(ro.dtype).contiguous()
        else:
            scale = torch.empty(1, device=images.device, dtype=torch.float32)
            in_disp = ro.disparity_renorm(disp, t0, NUM_INPUT_IMAGE, scale, True, False)    # :346-350
            self.scale = scale

        cur = torch.cat(self.trans3d, dim=1)[:, -NUM_INPUT_UNIT:]                            # :352-358
        cur = torch.matmul(_inv(cur[:, 0]).unsqueeze(1), cur)
        rel = cur.clone()
        rel[:, 1:] = torch.matmul(_inv(cur[:, :-1]), cur[:, 1:])
        rel[:, :, :3, 3] = _signed_sqrt(rel[:, :, :3, 3] / scale)                           # :360-361
        in_ray = ro.camera_raymap(self.trans2d[-1][:, -NUM_INPUT_UNIT:], rel, H, W)         # :362-368 (+ :300-301)

        t3 = torch.cat(self.trans3d, dim=1)                                                  # :372-377
        t2 = torch.cat(self.trans2d, dim=1)
        t3 = torch.matmul(_inv(t3[:, -NUM_INPUT_UNIT]).unsqueeze(1), t3)
        c2w = t3[0]
        dist = torch.norm(c2w[:-1, :3, 3] - c2w[-1, :3, 3], dim=1)                           # :382-393
        _, near = torch.topk(-dist, k=5)
        dots = torch.sum(c2w[near, :3, 2] * c2w[-1, :3, 2], dim=1)
        k = near[torch.argmin(torch.acos(torch.clamp(dots, -1.0, 1.0)))].reshape(1)          # stays on the device
        self.history_index = k
        h_img = torch.cat(self.key_images, dim=2).index_select(2, k)                         # :394-397
        h_disp = torch.cat(self.key_disps, dim=2).index_select(2, k).contiguous()
        h_disp = ro.disparity_renorm(h_disp, 0, 1, scale, False, True)                       # :399-401
        h3 = t3.index_select(1, k).clone()
        h3[:, :, :3, 3] = _signed_sqrt(h3[:, :, :3, 3] / scale)                              # :403-404
        h_ray = ro.camera_raymap(t2.index_select(1, k), h3, H, W)                            # :406-410 (+ :258-259)
        history = ro.history_latent(h_img, h_disp, h_ray, noise, shard)
        return frames, in_disp, in_ray, history
