"""Row f1 measurement: the tiled VAE encode of the rollout (pipeline.py:250-251,569-576) at the demo
shape, full-width encoder, random-init weights: ms and TFLOP/s (algorithmic FLOPs from the plan)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from deepv_b200 import synthetic as synth  # noqa: E402
from deepv_b200.vae import B200VAE  # noqa: E402

cfg, W = synth.vae_weights(None, seed=2, encoder=True)
vae = B200VAE(W, cfg, device="cuda", dtype=torch.bfloat16)
vae.enable_tiling()
for T in (1, 25):
    x = torch.randn(1, 3, T, 384, 512, device="cuda").bfloat16()
    for _ in range(2):
        vae.encode(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 3
    e0.record()
    for _ in range(n):
        d = vae.encode(x).latent_dist
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = vae.enc_plan_flops(T, 384, 512, 256)
    print(f"encode [1,3,{T},384,512] -> moments {tuple(d.parameters.shape)}: {ms:.2f} ms, {fl / 1e12:.2f} TFLOP, "
          f"{fl / ms / 1e9:.0f} TFLOP/s", flush=True)
