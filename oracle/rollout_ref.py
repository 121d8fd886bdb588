"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of the reference rollout around the hot path.

Restates /root/reference/pipeline.py:
  * `get_pyramid_latent`                         :226-240   -> pyramid_latent
  * `get_history_vae_latent`                     :243-262   -> history_latent
  * condition clip lists of `generate_i2v`       :621-658   -> condition_clips
  * `generate_i2v`                               :526-700   -> generate_i2v
  * `raymap_to_trans_matrix`                     :77-163    -> raymap_to_pose
  * `get_raymap_from_camera_parameters(_batch)`  :29-75     -> camera_raymap
  * the feedback between iterations of `generate`:264-424   -> Feedback / generate
with every random draw taken from an injected tape (tests/golden/rollout_cases.NoiseTape) in the
reference's call order.  The denoiser, scheduler and VAE are the restatements of mmdit_ref,
scheduler_ref and vae_ref.

Pinned against the unmodified `InferencePipeline.generate` run on the same tape:
tests/golden/rollout_golden.pt (made by tests/golden/make_rollout_golden.py), checked in
tests/test_oracle_rollout.py.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

from . import mmdit_ref, scheduler_ref, vae_ref

Tensor = torch.Tensor

RAYMAP_MEAN = (-0.0016, -0.0010, 0.9015, 0.0313, -0.0538, 0.2079)   # pipeline.py:200
RAYMAP_STD = (0.3333, 0.2567, 0.0927, 0.4338, 0.1746, 0.5802)       # pipeline.py:201
NUM_INPUT_IMAGE, NUM_INPUT_UNIT = 25, 4                              # pipeline.py:269-270


def _stat(v, like: Tensor) -> Tensor:
    return torch.tensor(v, dtype=like.dtype, device=like.device).view(1, 6, 1, 1, 1)


@dataclass
class RolloutModels:
    dit_cfg: dict
    dit_W: Dict[str, Tensor]
    vae_cfg: dict
    vae_W: Dict[str, Tensor]
    tables: dict                      # scheduler_ref.pyramid_tables(...)
    text_embeds: Dict[str, Dict[str, Tensor]]
    model_cfg: dict
    video_guidance_scale: float = 3.5  # pipeline.py:307
    _pos: Optional[Tensor] = field(default=None, repr=False)

    def pos_table(self):
        if self._pos is None:
            c = self.dit_cfg
            d = c["num_attention_heads"] * c["attention_head_dim"]
            self._pos = mmdit_ref.sincos_2d_table(d, c["pos_embed_max_size"], c["sample_size"] // c["patch_size"])
        return self._pos


# ---------------------------------------------------------------------------------------------
def pyramid_latent(x: Tensor, stage_num: int) -> List[Tensor]:
    """pipeline.py:226-240: per-frame bilinear halvings, coarsest first."""
    out = [x]
    for _ in range(stage_num):
        b, c, t, h, w = x.shape
        y = F.interpolate(x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w), size=(h // 2, w // 2), mode="bilinear")
        x = y.view(b, t, c, h // 2, w // 2).permute(0, 2, 1, 3, 4)
        out.append(x)
    return out[::-1]


def condition_clips(pyr: Sequence[Tensor], unit_index: int, firstframe_mask: bool, fpu: int,
                    n_branch: int) -> List[List[Tensor]]:
    """pipeline.py:621-658: for stage i_s the newest clean frame at that stage's resolution, each
    older frame one stage coarser, and all frames older than that (minus the masked first frame)
    as one stage-0 clip; oldest first, every clip repeated once per CFG branch."""
    offset = int(firstframe_mask)
    out = []
    for i_s in range(len(pyr)):
        clips = [pyr[i_s][:, :, -fpu:]]
        stage, taken = i_s, 1
        while taken < unit_index - offset:
            stage = max(stage - 1, 0)
            if stage == 0:
                break
            taken += 1
            clips.append(pyr[stage][:, :, -(taken * fpu): -((taken - 1) * fpu)])
        if stage == 0 and taken < unit_index - offset:
            clips.append(pyr[0][:, :, offset: -(taken * fpu)])
        out.append([torch.cat([c] * n_branch) for c in reversed(clips)])
    return out


def normalise_latent(z: Tensor) -> Tensor:
    """pipeline.py:570-571 (frame 0 with the image statistics, the rest with the video ones)."""
    z = z.clone()
    z[:, :, :1] = (z[:, :, :1] - vae_ref.VAE_SHIFT) * vae_ref.VAE_SCALE
    z[:, :, 1:] = (z[:, :, 1:] - vae_ref.VAE_VIDEO_SHIFT) * vae_ref.VAE_VIDEO_SCALE
    return z


def encode_sample(m: RolloutModels, x: Tensor, tape) -> Tensor:
    """`self.vae.encode(x).latent_dist.sample()` (pipeline.py:250,569) with the draw from the tape."""
    moments = vae_ref.tiled_encode(m.vae_W, m.vae_cfg, x)
    return vae_ref.gaussian_sample(moments, tape.randn(torch.chunk(moments, 2, dim=1)[0].shape))


def frames_to_input(frames_u8: Tensor) -> Tensor:
    """ToTensor + Normalize(0.5, 0.5) of uint8 [n,H,W,3] frames -> [1,3,n,H,W] (pipeline.py:564-568)."""
    x = frames_u8.permute(3, 0, 1, 2).float().div(255)
    return ((x - 0.5) / 0.5).unsqueeze(0)


# ---------------------------------------------------------------------------------------------
def raymap_to_pose(raymap: Tensor, vae_downsample: int = 8):
    """pipeline.py:77-163 as `generate_i2v` calls it (append_first_reference, relative -> absolute,
    scale 1).  raymap [b,6,t,h,w] -> camera-to-world [b,t+1,4,4], intrinsics [b,t+1,4,4].
    (The reference's dim-less torch.cross picks the last axis for every t != 3.)"""
    b, _, t, h, w = raymap.shape
    d = raymap[:, :3]
    ref = d.mean(dim=(-1, -2), keepdim=True)
    ref = ref / ref.norm(dim=1, keepdim=True)
    d = d / (d * ref).sum(dim=1, keepdim=True)
    d = d.permute(0, 2, 3, 4, 1)                         # [b,t,h,w,3]
    o = raymap[:, 3:].permute(0, 2, 3, 4, 1)
    o = torch.sign(o) * (o.abs() ** 2)

    location = o.reshape(b, t, -1, 3).mean(dim=-2)
    image_location = (o + d).reshape(b, t, -1, 3).mean(dim=-2)
    z_dir = image_location - location
    focal = torch.norm(z_dir, dim=-1)

    left = d[:, :, :, :1].reshape(b, t, -1, 3).mean(dim=-2)
    right = d[:, :, :, -1:].reshape(b, t, -1, 3).mean(dim=-2)
    w_real = torch.norm(torch.linalg.cross(right - left, z_dir, dim=-1), dim=-1) / (w - 1) * w
    up = d[:, :, :1].reshape(b, t, -1, 3).mean(dim=-2)
    down = d[:, :, -1:].reshape(b, t, -1, 3).mean(dim=-2)
    h_real = torch.norm(torch.linalg.cross(up - down, z_dir, dim=-1), dim=-1) / (h - 1) * h

    x_dir = right - left
    y_dir = torch.linalg.cross(z_dir, x_dir, dim=-1)
    x_dir = torch.linalg.cross(y_dir, z_dir, dim=-1)
    pose = torch.zeros(b, t, 4, 4, device=raymap.device)
    pose[:, :, :3, 0] = x_dir / torch.norm(x_dir, dim=-1, keepdim=True)
    pose[:, :, :3, 1] = y_dir / torch.norm(y_dir, dim=-1, keepdim=True)
    pose[:, :, :3, 2] = z_dir / torch.norm(z_dir, dim=-1, keepdim=True)
    pose[:, :, :3, 3] = location
    pose[:, :, 3, 3] = 1.0

    rescale = (w / w_real + h / h_real) / 2 * vae_downsample
    intr = torch.zeros(b, t, 4, 4, device=raymap.device)
    intr[:, :, 0, 0] = focal * rescale
    intr[:, :, 1, 1] = focal * rescale
    intr[:, :, 0, 2] = w / 2 * vae_downsample
    intr[:, :, 1, 2] = h / 2 * vae_downsample
    intr[:, :, 2, 2] = 1.0
    intr[:, :, 3, 3] = 1.0

    pose = torch.cat([torch.eye(4, device=raymap.device).expand(b, 1, 4, 4), pose], dim=1).to(raymap)
    intr = torch.cat([intr[:, :1], intr], dim=1).to(raymap)
    for i in range(t):
        pose[:, i + 1] = torch.bmm(pose[:, i], pose[:, i + 1])
    return pose, intr


def camera_raymap(trans2d: Tensor, trans3d: Tensor, depth_shape, vae_downsample: int = 8) -> Tensor:
    """pipeline.py:29-75: [b,t,4,4] intrinsics / camera-to-world -> ray map [b,t,6,H/ds,W/ds]:
    unit ray directions (8x8-averaged pixel rays, rotated to the world frame) and the ray origin."""
    H, W = depth_shape
    out = []
    for k2, k3 in zip(trans2d, trans3d):
        fu, fv = k2[:, 0, 0].view(-1, 1, 1), k2[:, 1, 1].view(-1, 1, 1)
        cu, cv = k2[:, 0, 2].view(-1, 1, 1), k2[:, 1, 2].view(-1, 1, 1)
        u = torch.arange(W, device=k2.device).view(1, 1, W).expand(k2.shape[0], H, W)
        v = torch.arange(H, device=k2.device).view(1, H, 1).expand(k2.shape[0], H, W)
        one = torch.ones_like(u)
        rays = torch.stack(((u - cu) / fu, (v - cv) / fv, one, one), dim=1).to(k3)      # [t,4,H,W]
        rays = F.avg_pool2d(rays, kernel_size=vae_downsample, stride=vae_downsample)
        t, _, hh, ww = rays.shape
        rot = k3.clone()
        rot[:, :3, 3] = 0.0
        dirs = torch.bmm(rot, rays.reshape(t, 4, hh * ww)).view(t, 4, hh, ww)[:, :3]
        dirs = dirs / dirs.norm(dim=1, keepdim=True)
        origin = torch.ones_like(dirs) * k3[:, :3, 3].view(t, 3, 1, 1)
        out.append(torch.cat([dirs, origin], dim=1))
    return torch.stack(out, dim=0)


# ---------------------------------------------------------------------------------------------
def generate_i2v(m: RolloutModels, motion_prompt: Sequence[str], frames_u8: Tensor,
                 input_disparity: Optional[Tensor], input_raymap: Optional[Tensor],
                 input_history: Optional[Tensor], tape, steps: Sequence[int]):
    """pipeline.py:526-700 in action-prompt mode with classifier-free guidance.

    frames_u8 [n,H,W,3] uint8 (the PIL frames); input_disparity [1,3,n,H,W] in [-1,1] or None (first
    iteration); input_raymap [1,6,n_lat,h,w] already normalised; input_history [1,38,1,h,w].
    Returns image, disparity (raw decodes), trans3d, trans2d and the generated latents."""
    cfg = m.model_cfg
    fpu, nst, ray = cfg["frame_per_unit"], len(cfg["stages"]), cfg["raymap_dim"]
    first = input_disparity is None
    temp = cfg["max_temporal_length"]
    H, W = frames_u8.shape[1], frames_u8.shape[2]
    C = m.dit_cfg["in_channels"]

    noise = tape.randn((1, C, temp + int(first), H // 8, W // 8))                      # :551
    for _ in range(nst - 1):                                                             # :554-557
        b, c, t, h, w = noise.shape
        y = F.interpolate(noise.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w), size=(h // 2, w // 2),
                          mode="bilinear") * 2
        noise = y.view(b, t, c, h // 2, w // 2).permute(0, 2, 1, 3, 4)
    num_units = noise.shape[2] // fpu

    z_img = normalise_latent(encode_sample(m, frames_to_input(frames_u8), tape))         # :569-571
    if first:
        z_disp = torch.zeros_like(z_img)
    else:
        z_disp = normalise_latent(encode_sample(m, input_disparity, tape))               # :574-576
    z_ray = torch.zeros_like(z_img[:, :ray, :1]) if input_raymap is None else input_raymap
    generated = [torch.cat([z_img, z_disp, z_ray], dim=1)]                               # :578-582

    n_branch = 2 if input_history is None else 3
    history = hmask = None
    if input_history is not None:                                                        # :479-483,495
        hlen = int((input_history.size(-1) / cfg["history_downsample_ratio"] / 2) *
                   (input_history.size(-2) / cfg["history_downsample_ratio"] / 2))
        history = torch.cat([input_history] * 3)
        hmask = torch.cat([torch.zeros(2, hlen), torch.ones(1, hlen)]).to(input_history.device)
    neg = m.text_embeds["empty"]
    start = 1 if first else (frames_u8.shape[0] - 1) // 8 + 1                            # :587
    for unit in range(start, num_units):
        pos = m.text_embeds[motion_prompt[unit - int(first)]]                            # :596-601
        order = [neg] + [pos] * (n_branch - 1)                                           # :609-617
        enc = torch.cat([e["prompt_embeds"] for e in order])
        pooled = torch.cat([e["pooled_prompt_embeds"] for e in order])
        mask = torch.cat([e["prompt_attention_mask"] for e in order])
        pyr = pyramid_latent(torch.cat(generated, dim=2), nst - 1)
        conds = condition_clips(pyr, unit, first, fpu, n_branch)

        def model_fn(clips, tt):
            return mmdit_ref.mmdit_forward(m.dit_W, m.dit_cfg, clips, tt.float(), enc, mask, pooled,
                                           history=history, history_mask=hmask,
                                           history_downsample_ratio=cfg["history_downsample_ratio"] if history is not None else None,
                                           pos_table=m.pos_table())

        lat = noise[:, :, unit * fpu:(unit + 1) * fpu]
        _, _, _, h0, w0 = lat.shape
        block = [tape.block(1, C, fpu, h0 * 2 ** s, w0 * 2 ** s, m.tables["gamma"]) for s in range(1, nst)]
        outs = scheduler_ref.generate_one_unit(model_fn, m.tables, lat, conds, block, n_branch, steps,
                                               m.video_guidance_scale, cfg["history_guidance_scale"])
        generated.append(outs[-1])
    if first:
        generated = generated[1:]                                                        # :680-681
    lat = torch.cat(generated, dim=2)
    z_image, z_disparity = torch.chunk(lat[:, :-ray], 2, dim=1)                          # :685-686
    raymap = lat[:, -ray:] * _stat(RAYMAP_STD, lat) + _stat(RAYMAP_MEAN, lat)            # :688-690
    trans3d, trans2d = raymap_to_pose(raymap[:, :, 1:].clone())                          # :692
    image = vae_ref.decode_latent(m.vae_W, m.vae_cfg, z_image)                           # :694-695
    disparity = vae_ref.decode_latent(m.vae_W, m.vae_cfg, z_disparity)
    return image, disparity, trans3d, trans2d, lat


def history_latent(m: RolloutModels, rgb: Tensor, disparity: Tensor, raymap: Tensor, tape) -> Tensor:
    """pipeline.py:243-262: one history frame -> [1,38,1,h,w] (only frame 0 exists, image statistics)."""
    video = encode_sample(m, rgb, tape)
    disp = encode_sample(m, disparity, tape)
    video[:, :, :1] = (video[:, :, :1] - vae_ref.VAE_SHIFT) * vae_ref.VAE_SCALE
    disp[:, :, :1] = (disp[:, :, :1] - vae_ref.VAE_SHIFT) * vae_ref.VAE_SCALE
    raymap = raymap.clone()
    raymap[:, :3] = raymap[:, :3] / raymap[:, :3].norm(dim=1, keepdim=True)
    raymap = (raymap - _stat(RAYMAP_MEAN, raymap)) / _stat(RAYMAP_STD, raymap)
    return torch.cat([video, disp, raymap], dim=1)


def signed_sqrt(x: Tensor) -> Tensor:
    return torch.sign(x) * torch.sqrt(x.abs())


class Feedback:
    """The state `generate` carries between iterations (pipeline.py:282-414)."""

    def __init__(self, m: RolloutModels):
        self.m = m
        self.images: List[Tensor] = []
        self.disparitys: List[Tensor] = []
        self.trans3d: List[Tensor] = []
        self.trans2d: List[Tensor] = []
        self.prompts: List[List[str]] = []
        self.scale = 1.0
        self.history_index = None

    def absorb(self, it: int, images, disparity_raw, trans3d, trans2d, motion_prompt):
        """pipeline.py:311-337: disparity post-processing, pose chaining, what is kept of an iteration."""
        disp = disparity_raw.mean(dim=1, keepdim=True).repeat(1, 3, 1, 1, 1) * 0.5 + 0.5
        disp = torch.clamp(disp, 0, 1) ** 2
        disp = disp / self.scale / 0.95
        trans3d = trans3d.clone()
        trans3d[:, :, :3, 3] = trans3d[:, :, :3, 3] * self.scale
        if it == 0:
            self.images.append(images)
            self.disparitys.append(disp)
            self.prompts.append(list(motion_prompt))
            self.trans3d.append(trans3d)
            self.trans2d.append(trans2d)
        else:
            self.images.append(images[:, :, NUM_INPUT_IMAGE:])
            self.disparitys.append(disp[:, :, NUM_INPUT_IMAGE:])
            self.prompts.append(list(motion_prompt[NUM_INPUT_UNIT:]))
            pre = self.trans3d[-1][:, -NUM_INPUT_UNIT]
            trans3d = torch.matmul(pre.unsqueeze(1), trans3d)
            self.trans3d.append(trans3d[:, NUM_INPUT_UNIT:])
            self.trans2d.append(trans2d[:, NUM_INPUT_UNIT:])
        return images, disp

    def next_inputs(self, images, disp, tape):
        """pipeline.py:339-414: uint8 frames, renormalised disparity, relative-pose ray map and the
        history frame for the next `generate_i2v`."""
        m = self.m
        ds = m.model_cfg["vae_downsample"]
        last = images[0, :, -NUM_INPUT_IMAGE:].permute(1, 2, 3, 0)
        frames_u8 = (torch.clamp(last * 0.5 + 0.5, 0, 1).to(torch.float32) * 255).to(torch.uint8)    # :341 (truncation)

        d = disp[:, :, -NUM_INPUT_IMAGE:]
        self.scale = 1 / d[:, :, 0].max()                                                             # :347
        d = torch.sqrt(d * self.scale * 0.95) * 2 - 1

        cur = torch.cat(self.trans3d, dim=1)[:, -NUM_INPUT_UNIT:].clone()                             # :352-358
        cur = torch.matmul(torch.inverse(cur[:, 0]).unsqueeze(1), cur)
        rel = cur.clone()
        rel[:, 1:] = torch.matmul(torch.inverse(cur[:, :-1]), cur[:, 1:])
        rel[:, :, :3, 3] = signed_sqrt(rel[:, :, :3, 3] / self.scale)                                 # :360-361
        raymap = camera_raymap(self.trans2d[-1][:, -NUM_INPUT_UNIT:], rel.to(d), d.shape[-2:], 8)     # :362-368
        raymap = raymap.permute(0, 2, 1, 3, 4)

        imgs = torch.cat(self.images, dim=2)[:, :, ::ds]                                              # :370-377
        disps = torch.cat(self.disparitys, dim=2)[:, :, ::ds]
        t3 = torch.cat(self.trans3d, dim=1)
        t2 = torch.cat(self.trans2d, dim=1)
        t3 = torch.matmul(torch.inverse(t3[:, -NUM_INPUT_UNIT]).unsqueeze(1), t3)
        c2w = t3.squeeze(0)
        dist = torch.norm(c2w[:-1, :3, 3] - c2w[-1, :3, 3], dim=1)                                    # :382-393
        _, near = torch.topk(-dist, k=5)
        dots = torch.sum(c2w[near, :3, 2] * c2w[-1, :3, 2], dim=1)
        k = near[torch.argmin(torch.acos(torch.clamp(dots, -1.0, 1.0)))].item()
        self.history_index = k

        h_img = imgs[:, :, k:k + 1]
        h_disp = torch.clamp(torch.sqrt(disps[:, :, k:k + 1] * self.scale * 0.95) * 2 - 1, -1, 1)      # :399-401
        h3 = t3[:, k:k + 1].clone()
        h3[:, :, :3, 3] = signed_sqrt(h3[:, :, :3, 3] / self.scale)                                    # :403-404
        h_ray = camera_raymap(t2[:, k:k + 1], h3, h_disp.shape[-2:], ds).permute(0, 2, 1, 3, 4)        # :406-410
        history = history_latent(m, h_img, h_disp, h_ray, tape)
        return frames_u8, d, raymap, history


def plan_prompts(prompts: Sequence[str], actual_unit: int = 8):
    """pipeline.py:275-279: pad the prompt list with its last entry, count the iterations."""
    p = list(prompts)
    step = actual_unit - NUM_INPUT_UNIT
    while (len(p) - actual_unit) % step != 0 or len(p) < actual_unit:
        p.append(p[-1])
    return p, (len(p) - actual_unit) // step + 1


def generate(m: RolloutModels, first_frame_u8: Tensor, prompts: Sequence[str], tape, steps: Sequence[int],
             trace: Optional[list] = None):
    """pipeline.py:264-424 (`prompt_type == 'action'`): the whole autoregressive rollout."""
    total, iters = plan_prompts(prompts, m.model_cfg["max_temporal_length"])
    fb = Feedback(m)
    frames = first_frame_u8.unsqueeze(0)
    in_disp = in_ray = in_hist = None
    start_unit = 0
    for it in range(iters):
        motion = total[0:1] + total[start_unit + 1: start_unit + m.model_cfg["max_temporal_length"]]  # :296
        if in_ray is not None:
            in_ray = (in_ray - _stat(RAYMAP_MEAN, in_ray)) / _stat(RAYMAP_STD, in_ray)                # :300-301
        image, disparity, t3, t2, lat = generate_i2v(m, motion, frames, in_disp, in_ray, in_hist, tape, steps)
        if trace is not None:
            trace.append(dict(motion_prompt=motion, frames=frames, input_disparity=in_disp, input_raymap=in_ray,
                              input_history=in_hist, images=image, disparity=disparity, trans3d=t3, trans2d=t2,
                              latents=lat))
        start_unit += m.model_cfg["max_temporal_length"] - NUM_INPUT_UNIT
        image, disp = fb.absorb(it, image, disparity, t3, t2, motion)
        frames, in_disp, in_ray, in_hist = fb.next_inputs(image, disp, tape)
    return dict(pred_img=torch.cat(fb.images, dim=2), pred_disparity=torch.cat(fb.disparitys, dim=2),
                trans3d=torch.cat(fb.trans3d, dim=1), trans2d=torch.cat(fb.trans2d, dim=1),
                motion_prompt_list=fb.prompts)
