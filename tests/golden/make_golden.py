"""Generate the golden fixtures from the REAL reference (run in the build container only).

    PYTHONPATH=/root/repo python tests/golden/make_golden.py

Imports /root/reference/model/*.py unmodified through oracle/_shim.py, feeds it the seeded
synthetic weights of oracle/weights.py and seeded inputs (tests/golden/cases.py), and stores the
reference's outputs.  The fixtures pin oracle/ (and through it the CUDA path) on machines where
/root/reference does not exist.
"""
import json
import os
import sys
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))

from oracle import _shim, weights  # noqa: E402
from tests.golden import cases  # noqa: E402


def main():
    mm, sc, va = _shim.import_reference()
    torch.set_grad_enabled(False)

    # ---- scheduler tables ------------------------------------------------
    s = sc.PyramidFlowMatchEulerDiscreteScheduler(**cases.SCHEDULER_KW)
    g = {"kw": cases.SCHEDULER_KW,
         "start_sigmas": {str(k): v for k, v in s.start_sigmas.items()},
         "end_sigmas": {str(k): v for k, v in s.end_sigmas.items()},
         "ori_start_sigmas": {str(k): v for k, v in s.ori_start_sigmas.items()},
         "timestep_ratios": {str(k): list(v) for k, v in s.timestep_ratios.items()},
         "stages": {}}
    for n in (1, 5):
        for i in range(3):
            s.set_timesteps(n, i)
            g["stages"][f"{n}_{i}"] = {"timesteps": s.timesteps.tolist(), "sigmas": s.sigmas.tolist()}
    (HERE / "scheduler_golden.json").write_text(json.dumps(g, indent=1))

    # ---- scheduler.step in bf16 and fp32 --------------------------------------------------------
    step = {}
    for dt in (torch.bfloat16, torch.float32):
        x, v = cases.step_inputs(dt)
        s.set_timesteps(5, 0)
        outs = []
        cur = x
        for k in range(5):
            cur = s.step(model_output=v, timestep=s.timesteps[k], sample=cur).prev_sample
            outs.append(cur.clone())
        step[str(dt)] = torch.stack(outs)
    torch.save(step, HERE / "scheduler_step_golden.pt")

    # ---- MMDiT forwards ---------------------------------------------------------------------------
    out = {}
    for name, case in cases.MMDIT_CASES.items():
        cfg, W = weights.mmdit_weights(case["cfg"], seed=case["wseed"])
        keys = ("sample_size", "patch_size", "in_channels", "num_layers", "attention_head_dim",
                "num_attention_heads", "caption_projection_dim", "pooled_projection_dim",
                "pos_embed_max_size", "max_num_frames", "qk_norm", "pos_embed_type",
                "temp_pos_embed_type", "joint_attention_dim", "use_temporal_causal",
                "add_temp_pos_embed", "interp_condition_pos")
        model = mm.MMDiT(**{k: cfg[k] for k in keys}).eval()
        missing, unexpected = model.load_state_dict(W, strict=False)
        assert missing == ["pos_embed.pos_embed"] and not unexpected, (missing, unexpected)
        inp = cases.mmdit_inputs(case)
        y = model(sample=[inp["clips"]], timestep_ratio=inp["t"], encoder_hidden_states=inp["enc"],
                  encoder_attention_mask=inp["mask"], pooled_projections=inp["pooled"],
                  history=inp["hist"], history_mask=inp["hmask"],
                  history_downsample_ratio=2 if inp["hist"] is not None else None)[0]
        out[name] = y.clone()
        print(name, tuple(y.shape), float(y.abs().max()))
    torch.save(out, HERE / "mmdit_golden.pt")

    # ---- VAE decode ---------------------------------------------------------------------------------
    out = {}
    for name, case in cases.VAE_CASES.items():
        cfg, W = weights.vae_weights(case["cfg"], seed=case["wseed"])
        keys = ("encoder_out_channels", "decoder_in_channels", "encoder_block_out_channels",
                "decoder_block_out_channels", "encoder_layers_per_block", "decoder_layers_per_block",
                "encoder_spatial_down_sample", "decoder_spatial_up_sample",
                "encoder_temporal_down_sample", "decoder_temporal_up_sample", "interpolate")
        vae = va.CausalVideoVAE(**{k: cfg[k] for k in keys}).eval()
        missing, unexpected = vae.load_state_dict(W, strict=False)
        assert not unexpected and all(m.startswith(("encoder", "quant_conv")) for m in missing)
        vae.enable_tiling()
        z = cases.vae_latent(case)
        y = vae.decode(z.clone(), temporal_chunk=True, window_size=1, tile_sample_min_size=256).sample
        out[name] = cases.vae_digest(y)
        print(name, tuple(y.shape), float(y.abs().max()))
    torch.save(out, HERE / "vae_golden.pt")

    # ---- VAE encode: moments (mean | logvar) of the tiled, un-chunked encode ---------------------------
    out = {}
    for name, case in cases.VAE_ENC_CASES.items():
        cfg, W = weights.vae_weights(case["cfg"], seed=case["wseed"], encoder=True)
        vae = va.CausalVideoVAE(**{k: cfg[k] for k in keys}).eval()
        vae.load_state_dict(W, strict=True)
        vae.enable_tiling()
        m = vae.encode(cases.vae_video(case)).latent_dist.parameters
        out[name] = m.clone().float()
        print(name, tuple(m.shape), float(m.abs().max()))
    torch.save(out, HERE / "vae_encode_golden.pt")
    for f in sorted(HERE.glob("*golden*")):
        print(f.name, os.path.getsize(f))


if __name__ == "__main__":
    main()
