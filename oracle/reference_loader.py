"""TEST INFRASTRUCTURE ONLY — the UNMODIFIED reference (model/mmdit.py, model/vae.py, model/scheduler.py,
pipeline.py) instantiated with seeded synthetic weights, on any device.

The reference cannot be pip-installed (no setup.py / pyproject.toml) and /root/reference does not exist
on the GPU box, so `stage()` copies its five Python sources into the git-ignored `baseline/_ref/`
(never into history; `gpurun` snapshots carry it).  `oracle/_shim.py` then imports them from
wherever they are found.  Users: the boundary test (tests/test_gpu_boundary.py: the reference's own
`generate_one_unit` / `decode_latent` loops over the B200 objects), `bench.py --impl reference`
(CPU arm) and `bench.py --same-box-eager` (the reference's own bf16 eager CUDA path as a reported bar).
Nothing under deepv_b200/ imports this.
"""
from __future__ import annotations

import importlib
import os
import shutil
from pathlib import Path

import torch

from . import _shim

ROOT = Path(__file__).resolve().parents[1]
STAGED = ROOT / "baseline" / "_ref"
FILES = ("pipeline.py", "run.py", "model/mmdit.py", "model/vae.py", "model/scheduler.py")

MMDIT_KEYS = ("sample_size", "patch_size", "in_channels", "num_layers", "attention_head_dim",
              "num_attention_heads", "caption_projection_dim", "pooled_projection_dim",
              "pos_embed_max_size", "max_num_frames", "qk_norm", "pos_embed_type",
              "temp_pos_embed_type", "joint_attention_dim", "use_temporal_causal",
              "add_temp_pos_embed", "interp_condition_pos")
VAE_KEYS = ("encoder_out_channels", "decoder_in_channels", "encoder_block_out_channels",
            "decoder_block_out_channels", "encoder_layers_per_block", "decoder_layers_per_block",
            "encoder_spatial_down_sample", "decoder_spatial_up_sample",
            "encoder_temporal_down_sample", "decoder_temporal_up_sample", "interpolate")


def stage(src: str = "/root/reference") -> bool:
    """Copy the reference's sources to baseline/_ref/ (git-ignored).  Returns False when `src` is absent."""
    if not os.path.isfile(os.path.join(src, "model", "mmdit.py")):
        return False
    for rel in FILES:
        dst = STAGED / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), dst)
    return True


def available() -> bool:
    return _shim.reference_available()


def modules():
    """(mmdit, scheduler, vae, pipeline) modules of the real reference."""
    mm, sc, va = _shim.import_reference()
    return mm, sc, va, importlib.import_module("pipeline")


def build_mmdit(cfg: dict, W: dict, device="cpu", dtype=torch.float32):
    """The reference MMDiT with the given state dict.  Constructed on the meta device (its own random init of
    2 B parameters costs ~1 min of CPU and is overwritten anyway) and the tensors of `W` are assigned, not copied;
    the sincos table the constructor computes with numpy (mmdit.py:820-824) stays a real buffer."""
    mm, _, _, _ = modules()
    with torch.device("meta"):
        dit = mm.MMDiT(**{k: cfg[k] for k in MMDIT_KEYS}).eval()
    missing, unexpected = dit.load_state_dict(W, strict=False, assign=True)
    assert missing == ["pos_embed.pos_embed"] and not unexpected, (missing, unexpected)
    assert not any(p.is_meta for p in dit.parameters()) and not any(b.is_meta for b in dit.buffers())
    dit.in_channels = cfg["in_channels"]     # diffusers' ModelMixin resolves this from .config (pipeline.py:551)
    return dit.to(device=device, dtype=dtype)


def build_vae(cfg: dict, W: dict, device="cpu", dtype=torch.float32):
    _, _, va, _ = modules()
    vae = va.CausalVideoVAE(**{k: cfg[k] for k in VAE_KEYS}).eval()
    missing, unexpected = vae.load_state_dict(W, strict=False)
    assert not unexpected and all(k.startswith(("encoder.", "quant_conv.")) for k in missing), (missing, unexpected)
    vae.enable_tiling()
    return vae.to(device=device, dtype=dtype)


class _TextEncoderStub(torch.nn.Module):
    """`_create_models` returns a text encoder that `__init__` moves to the device (pipeline.py:192); the hot
    path never calls it in action mode (pipeline.py:596-601)."""

    def forward(self, *a, **k):
        raise RuntimeError("text encoder weights are unavailable offline")


def build_pipeline(models, scheduler_kw: dict, model_cfg: dict, text_embeds: dict, device="cpu",
                   dtype=torch.float32, subclass_hook=None):
    """An `InferencePipeline` whose `_create_models` (pipeline.py:203-223, the drop-in seam) returns
    `models = (dit, vae, scheduler | None)`; the reference's own `__init__` runs unmodified.
    `subclass_hook(cls)` may return a subclass to instantiate instead."""
    import tempfile
    _, sc, _, pl = modules()
    dit, vae, sched = models
    if sched is None:
        sched = sc.PyramidFlowMatchEulerDiscreteScheduler(**scheduler_kw)

    class _Injected(pl.InferencePipeline):
        def _create_models(self):
            return dit, vae, sched, _TextEncoderStub()

    cls = subclass_hook(_Injected) if subclass_hook else _Injected
    with tempfile.NamedTemporaryFile(suffix=".pt") as f:
        torch.save(text_embeds, f.name)
        cfg = dict(model_cfg, text_embeds_path=f.name)
        pipe = cls(cfg, device=str(device), torch_dtype=dtype)
    return pl, pipe
