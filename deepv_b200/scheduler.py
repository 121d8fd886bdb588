"""Host-side mirror of `PyramidFlowMatchEulerDiscreteScheduler` (reference model/scheduler.py).

Same constructor arguments, attributes and call surface as the pipeline uses
(pipeline.py:218,432,449-450,457-458,515-520): `set_timesteps(n, stage_index, device=)`,
`.timesteps`, `.sigmas`, `.ori_start_sigmas`, `.config.gamma`, `step(...).prev_sample`.
The tables are tiny host float math (done once per stage); the per-step tensor update runs in
the `dv_cfg_euler_step` kernel, bit-exact with the reference's ATen arithmetic, optionally fused
with the classifier-free-guidance combine (`cfg_step`).
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check


class B200Scheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, shift: float = 1.0, stages: int = 3,
                 stage_range: Sequence[float] = (0, 1 / 3, 2 / 3, 1), gamma: float = 1 / 3):
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, shift=shift,
                                      stages=stages, stage_range=list(stage_range), gamma=gamma)
        self.gamma = gamma
        self.timestep_ratios, self.start_sigmas, self.end_sigmas, self.ori_start_sigmas = {}, {}, {}, {}
        self._stage_t, self._stage_s = {}, {}
        self._build_tables()
        self.timesteps = None
        self.sigmas = None
        self._sigmas_host: Optional[np.ndarray] = None
        self._step_index: Optional[int] = None
        self.num_inference_steps = None

    # scheduler.py:70-149 ----------------------------------------------------------------
    def _global_sigmas(self) -> np.ndarray:
        n, shift = self.config.num_train_timesteps, self.config.shift
        t = np.linspace(1, n, n, dtype=np.float32)[::-1].copy()
        s = (t / np.float32(n)).astype(np.float32)
        s = (np.float32(shift) * s / (np.float32(1) + np.float32(shift - 1) * s)).astype(np.float32)
        return s

    def _build_tables(self) -> None:
        n, stages, rng, gamma = (self.config.num_train_timesteps, self.config.stages,
                                 self.config.stage_range, self.config.gamma)
        sig = self._global_sigmas()
        tsteps = (sig * np.float32(n)).astype(np.float32)
        dist: List[float] = []
        for i in range(stages):
            lo = max(int(rng[i] * n), 0)
            hi = min(int(rng[i + 1] * n), n)
            s0 = float(sig[lo])
            s1 = float(sig[hi]) if hi < n else 0.0
            self.ori_start_sigmas[i] = s0
            if i != 0:
                o = 1 - s0
                s0 = 1 - (1 / (math.sqrt(1 + (1 / gamma)) * (1 - o) + o)) * o
            dist.append(s0 - s1)
            self.start_sigmas[i], self.end_sigmas[i] = s0, s1
        tot = sum(dist)
        for i in range(stages):
            r0 = 0.0 if i == 0 else sum(dist[:i]) / tot
            r1 = 1.0 if i == stages - 1 else sum(dist[:i + 1]) / tot
            self.timestep_ratios[i] = (r0, r1)
        for i in range(stages):
            r0, r1 = self.timestep_ratios[i]
            tmax = tsteps[int(r0 * n)]                 # fp32 scalars: the reference's linspace
            tmin = tsteps[min(int(r1 * n), n - 1)]     # runs in fp32 (scheduler.py:140-145)
            per = np.linspace(tmax, tmin, n + 1)[:-1]
            self._stage_t[i] = (float(per[0]), float(per[-1]))
            sp = np.linspace(1, 0, n + 1)[:-1]
            self._stage_s[i] = (float(sp[0]), float(sp[-1]))
        self.sigma_min, self.sigma_max = float(sig[-1]), float(sig[0])

    # scheduler.py:179-206 -------------------------------------------------------------------
    def set_timesteps(self, num_inference_steps: int, stage_index: int, device=None):
        self.num_inference_steps = num_inference_steps
        tmax, tmin = self._stage_t[stage_index]
        ts = np.linspace(tmax, tmin, num_inference_steps)
        smax, smin = self._stage_s[stage_index]
        sg = np.concatenate([np.linspace(smax, smin, num_inference_steps), np.zeros(1)])
        self._timesteps_host, self._sigmas_host = ts, sg
        self.timesteps = torch.from_numpy(ts).to(device=device)
        self.sigmas = torch.from_numpy(sg).to(device=device)
        self._step_index = None

    @property
    def step_index(self):
        return self._step_index

    # scheduler.py:230-294 ---------------------------------------------------------------------
    def step(self, model_output: torch.Tensor, timestep=None, sample: torch.Tensor = None,
             generator=None, return_dict: bool = True):
        if isinstance(timestep, int) or isinstance(timestep, (torch.IntTensor, torch.LongTensor)):
            raise ValueError("Passing integer indices as timesteps to `step()` is not supported. "
                             "Make sure to pass one of the `scheduler.timesteps` as a timestep.")
        prev = self.cfg_step(model_output, sample, n_branch=1)
        if not return_dict:
            return (prev,)
        return SimpleNamespace(prev_sample=prev)

    def cfg_step(self, noise_pred: torch.Tensor, sample: torch.Tensor, n_branch: int,
                 w_text: float = 0.0, w_hist: float = 0.0) -> torch.Tensor:
        """Fused CFG combine (pipeline.py:502-513) + Euler update.  noise_pred: [n_branch*b, ...]
        stacked as (uncond, text[, text+history]); sample: [b, ...] in the same dtype."""
        if self._step_index is None:
            self._step_index = 0
        _lib.require_cuda(noise_pred, sample)
        lib = _lib.load()
        if noise_pred.dtype != sample.dtype:
            sample = sample.to(noise_pred.dtype)
        noise_pred = noise_pred.contiguous()
        sample = sample.contiguous()
        numel = sample.numel()
        if noise_pred.numel() != n_branch * numel:
            raise _lib.DeepVError(f"cfg_step: noise_pred has {noise_pred.numel()} elements, expected "
                                  f"{n_branch} x {numel}")
        out = torch.empty_like(sample)
        i = self._step_index
        check(lib.dv_cfg_euler_step(noise_pred.data_ptr(), n_branch, sample.data_ptr(), out.data_ptr(),
                                    numel, float(w_text), float(w_hist), float(self._sigmas_host[i]),
                                    float(self._sigmas_host[i + 1]), _lib.dtype_code(sample.dtype),
                                    _lib.stream_ptr()), "dv_cfg_euler_step")
        self._step_index += 1
        return out

    def __len__(self):
        return self.config.num_train_timesteps
