"""Golden rollout from the REAL reference (run in the build container only).

    PYTHONPATH=/root/repo python tests/golden/make_rollout_golden.py

Runs the unmodified `InferencePipeline.generate` (pipeline.py:264-424: two iterations of
`generate_i2v`, the uint8 / disparity / pose feedback between them and the history selection) on
CPU in fp32 on the tiny case of tests/golden/rollout_cases.py.  Only `_create_models`' results are
substituted (seeded synthetic weights instead of checkpoints) and the three noise sources are
redirected to a seeded tape; everything else is the reference's own code.
"""
import sys
import time
from pathlib import Path

import numpy as np
import torch
from PIL import Image

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))

from oracle import _shim, weights  # noqa: E402
from tests.golden import cases, rollout_cases as rc  # noqa: E402

MMDIT_KEYS = ("sample_size", "patch_size", "in_channels", "num_layers", "attention_head_dim",
              "num_attention_heads", "caption_projection_dim", "pooled_projection_dim",
              "pos_embed_max_size", "max_num_frames", "qk_norm", "pos_embed_type",
              "temp_pos_embed_type", "joint_attention_dim", "use_temporal_causal",
              "add_temp_pos_embed", "interp_condition_pos")
VAE_KEYS = ("encoder_out_channels", "decoder_in_channels", "encoder_block_out_channels",
            "decoder_block_out_channels", "encoder_layers_per_block", "decoder_layers_per_block",
            "encoder_spatial_down_sample", "decoder_spatial_up_sample",
            "encoder_temporal_down_sample", "decoder_temporal_up_sample", "interpolate")


def build_reference_pipeline(case, tape):
    import importlib
    mm, sc, va = _shim.import_reference()
    pl = importlib.import_module("pipeline")
    cfg, W = weights.mmdit_weights(case["dit"]["cfg"], seed=case["dit"]["wseed"])
    dit = mm.MMDiT(**{k: cfg[k] for k in MMDIT_KEYS}).eval()
    missing, unexpected = dit.load_state_dict(W, strict=False)
    assert missing == ["pos_embed.pos_embed"] and not unexpected
    dit.in_channels = cfg["in_channels"]     # diffusers' ModelMixin resolves this from .config (pipeline.py:551)
    vcfg, VW = weights.vae_weights(case["vae"]["cfg"], seed=case["vae"]["wseed"], encoder=True)
    vae = va.CausalVideoVAE(**{k: vcfg[k] for k in VAE_KEYS}).eval()
    vae.load_state_dict(VW, strict=True)
    vae.enable_tiling()

    pipe = object.__new__(pl.InferencePipeline)          # __init__ only loads checkpoints (pipeline.py:171-201)
    pipe.device, pipe.dtype = torch.device("cpu"), torch.float32
    pipe.model_cfg = dict(case["model_cfg"])
    pipe.downsample = 8
    pipe.model, pipe.vae = dit, vae
    pipe.scheduler = sc.PyramidFlowMatchEulerDiscreteScheduler(**cases.SCHEDULER_KW)
    pipe.text_encoder = None
    pipe.vae_shift_factor, pipe.vae_scale_factor = 0.1490, 1 / 1.8415
    pipe.vae_video_shift_factor, pipe.vae_video_scale_factor = -0.2343, 1 / 3.0986
    pipe.text_embeds = rc.text_embeds(case)
    pipe.raymap_mean = torch.tensor([-0.0016, -0.0010, 0.9015, 0.0313, -0.0538, 0.2079]).view(6, 1, 1, 1)
    pipe.raymap_std = torch.tensor([0.3333, 0.2567, 0.0927, 0.4338, 0.1746, 0.5802]).view(6, 1, 1, 1)

    def taped_randn(shape, generator=None, device=None, dtype=None, layout=None):
        return tape.randn(shape).to(device=device, dtype=dtype)

    pl.randn_tensor = taped_randn
    va.randn_tensor = taped_randn
    gamma = pipe.scheduler.config.gamma
    pipe.sample_block_noise = lambda bs, ch, t, h, w: tape.block(bs, ch, t, h, w, gamma)
    return pl, pipe


def main():
    torch.set_grad_enabled(False)
    case = rc.ROLLOUT
    tape = rc.NoiseTape(case["seed"] + 3)
    pl, pipe = build_reference_pipeline(case, tape)

    # record what generate() feeds each generate_i2v call and what comes back
    calls = []
    inner = pipe.generate_i2v

    def spy(motion_prompt, use_motion_prompt, input_image, input_disparity, input_raymap, input_history, **kw):
        rec = dict(motion_prompt=[str(p) for p in motion_prompt], n_images=len(input_image),
                   input_disparity=None if input_disparity is None else rc.digest_frames(input_disparity),
                   input_raymap=None if input_raymap is None else input_raymap.clone(),
                   input_history=None if input_history is None else input_history.clone())
        frames = torch.stack([torch.from_numpy(np.asarray(im)) for im in input_image])      # [n, H, W, 3] uint8
        rec["input_frames_sum"] = frames.long().sum(dim=(1, 2, 3))
        rec["input_frames_sub"] = frames[:, ::8, ::8].clone()
        out = inner(motion_prompt, use_motion_prompt, input_image, input_disparity, input_raymap, input_history, **kw)
        rec["images"] = rc.digest_frames(out[0])
        rec["disparity"] = rc.digest_frames(out[1])
        rec["trans3d"], rec["trans2d"] = out[2].clone(), out[3].clone()
        calls.append(rec)
        return out

    pipe.generate_i2v = spy
    t0 = time.time()
    res = pipe.generate(dict(img=Image.fromarray(rc.first_frame(case).numpy()), prompt=np.array(rc.prompts(case)),
                             prompt_type="action"))
    print(f"reference generate(): {time.time() - t0:.1f} s, {len(calls)} iterations, {len(tape.calls)} noise draws")
    gold = dict(calls=calls, tape_calls=tape.calls,
                pred_img=rc.digest_frames(res["pred_img"]), pred_disparity=rc.digest_frames(res["pred_disparity"]),
                trans3d=res["trans3d"].clone(), trans2d=res["trans2d"].clone(),
                motion_prompt_list=[[str(p) for p in m] for m in res["motion_prompt_list"]])
    torch.save(gold, HERE / "rollout_golden.pt")
    print("pred_img", tuple(res["pred_img"].shape), "pred_disparity", tuple(res["pred_disparity"].shape),
          "trans3d", tuple(res["trans3d"].shape))
    import os
    print("rollout_golden.pt", os.path.getsize(HERE / "rollout_golden.pt"))


if __name__ == "__main__":
    main()
