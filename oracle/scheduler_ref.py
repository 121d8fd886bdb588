"""TEST INFRASTRUCTURE ONLY — restatement of the reference scheduler and sampler loop.

* pyramid_tables / stage_schedule / euler_step restate
  /root/reference/model/scheduler.py:70-149 (init), :179-206 (set_timesteps), :274-289 (step).
* cfg_combine / renoise_coefficients / generate_one_unit restate
  /root/reference/pipeline.py:439-524 (stage x step loop, CFG, stage transition).

Golden pins: SURVEY.md App. D values (printed by the real reference) live in
tests/golden/scheduler_golden.json; tests/test_oracle_vs_reference.py re-derives them from the
real reference when /root/reference is present.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F


def pyramid_tables(num_train_timesteps: int = 1000, shift: float = 1.0, stages: int = 3,
                   stage_range: Sequence[float] = (0, 1 / 3, 2 / 3, 1), gamma: float = 1 / 3) -> Dict:
    """scheduler.py:70-149.  Returns start/end/ori_start sigmas, timestep ratios and the per-stage
    (timestep_max, timestep_min) pairs that set_timesteps interpolates between."""
    n = num_train_timesteps
    # init_sigmas (:77-88): fp32 linspace 1..n reversed, sigma = t/n with the SD3 shift
    ts = np.linspace(1, n, n, dtype=np.float32)[::-1].copy()
    sig = torch.from_numpy(ts) / n
    sig = shift * sig / (1 + (shift - 1) * sig)
    timesteps = sig * n
    start, end, ori, dist = {}, {}, {}, []
    for i in range(stages):
        lo = max(int(stage_range[i] * n), 0)
        hi = min(int(stage_range[i + 1] * n), n)
        s0 = sig[lo].item()
        s1 = sig[hi].item() if hi < n else 0.0
        ori[i] = s0
        if i != 0:  # gamma-corrected start point (:112-117)
            o = 1 - s0
            s0 = 1 - (1 / (math.sqrt(1 + (1 / gamma)) * (1 - o) + o)) * o
        dist.append(s0 - s1)
        start[i], end[i] = s0, s1
    tot = sum(dist)
    ratios, t_range = {}, {}
    for i in range(stages):
        r0 = 0.0 if i == 0 else sum(dist[:i]) / tot
        r1 = 1.0 if i == stages - 1 else sum(dist[:i + 1]) / tot
        ratios[i] = (r0, r1)
    for i in range(stages):
        tmax = timesteps[int(ratios[i][0] * n)]
        tmin = timesteps[min(int(ratios[i][1] * n), n - 1)]
        # timesteps_per_stage = linspace(tmax, tmin, n+1)[:-1] (:142-145); set_timesteps only reads
        # its first and last element
        per_stage = np.linspace(tmax, tmin, n + 1)[:-1]
        t_range[i] = (float(per_stage[0]), float(per_stage[-1]))
    # sigmas_per_stage = linspace(1, 0, n+1)[:-1] (:146-149): first 1.0, last 1/n
    s_per_stage = np.linspace(1, 0, n + 1)[:-1]
    return dict(start_sigmas=start, end_sigmas=end, ori_start_sigmas=ori, timestep_ratios=ratios,
                stage_t_range=t_range, stage_sigma_range=(float(s_per_stage[0]), float(s_per_stage[-1])),
                gamma=gamma)


def stage_schedule(tables: Dict, num_inference_steps: int, stage_index: int):
    """scheduler.py:179-206: fp64 timesteps [n] and sigmas [n+1] (last = 0)."""
    tmax, tmin = tables["stage_t_range"][stage_index]
    timesteps = np.linspace(tmax, tmin, num_inference_steps)
    smax, smin = tables["stage_sigma_range"]
    sigmas = np.concatenate([np.linspace(smax, smin, num_inference_steps), np.zeros(1)])
    return timesteps, sigmas


def euler_step(sample: torch.Tensor, model_output: torch.Tensor, sigma: float, sigma_next: float):
    """scheduler.py:278-286 with sigma a 0-dim fp64 tensor (does not promote)."""
    ds = torch.tensor(sigma_next, dtype=torch.float64) - torch.tensor(sigma, dtype=torch.float64)
    prev = sample.to(torch.float32) + ds * model_output
    return prev.to(model_output.dtype)


def cfg_combine(pred: torch.Tensor, w_text: float, w_hist: float) -> torch.Tensor:
    """pipeline.py:502-513.  pred: [n_branch * b, ...] stacked (uncond, text[, text+history])."""
    nb = pred.shape[0]
    if nb == 1:
        return pred
    if nb == 2:
        u, t = pred.chunk(2)
        return u + w_text * (t - u)
    u, t, h = pred.chunk(3)
    return u + w_text * (t - u) + w_hist * (h - t)


def renoise_coefficients(ori_start_sigma: float, gamma: float):
    """pipeline.py:457-460."""
    o = 1 - ori_start_sigma
    alpha = 1 / (math.sqrt(1 + (1 / gamma)) * (1 - o) + o)
    beta = alpha * (1 - o) / math.sqrt(gamma)
    return alpha, beta


def block_noise_cov(gamma: float) -> torch.Tensor:
    """pipeline.py:433: covariance of one 2x2 noise block."""
    return torch.eye(4) * (1 + gamma) - torch.ones(4, 4) * gamma


def generate_one_unit(model_fn: Callable, tables: Dict, latents: torch.Tensor,
                      past_conditions: List[List[torch.Tensor]], block_noise: List[torch.Tensor],
                      n_branch: int, steps: Sequence[int], w_text: float, w_hist: float,
                      timestep_dtype=None, no_need_depth: bool = False):
    """pipeline.py:439-524 with injected block noise (SURVEY.md §7 vi).

    model_fn(clips, timestep[B]) -> [B, C, 1, h, w]; latents [1, C, 1, h0, w0];
    past_conditions[i_s]: clips (already repeated n_branch times) for stage i_s;
    block_noise[i_s - 1]: [1, C, 1, h, w] for stages >= 1.  Returns the per-stage latents.
    """
    outs = []
    for i_s in range(len(steps)):
        timesteps, sigmas = stage_schedule(tables, steps[i_s], i_s)
        if i_s > 0:
            b, c, t, h, w = latents.shape
            up = F.interpolate(latents.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w),
                               size=(2 * h, 2 * w), mode="nearest")
            latents = up.view(b, t, c, 2 * h, 2 * w).permute(0, 2, 1, 3, 4)
            alpha, beta = renoise_coefficients(tables["ori_start_sigmas"][i_s], tables["gamma"])
            latents = alpha * latents + beta * block_noise[i_s - 1].to(latents.dtype)
        for idx in range(steps[i_s]):
            x_in = torch.cat([latents] * n_branch)
            clips = list(past_conditions[i_s])
            if no_need_depth:                       # pipeline.py:476-478 (with CFG the model input is a copy)
                clips = [c.clone() for c in clips]
                for c in clips + [x_in]:
                    c[:, 16:] *= 0
            tval = torch.tensor(timesteps[idx], dtype=torch.float64)
            tt = tval.expand(x_in.shape[0]).to(device=x_in.device, dtype=timestep_dtype or x_in.dtype)  # pipeline.py:473
            pred = model_fn(clips + [x_in], tt)
            guided = cfg_combine(pred, w_text, w_hist)
            latents = euler_step(latents, guided, float(sigmas[idx]), float(sigmas[idx + 1]))
        outs.append(latents)
    return outs
