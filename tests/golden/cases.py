"""Seeded inputs shared by the golden generator and the tests (CPU generators only, so the
same tensors are reproduced on every machine)."""
import torch

SCHEDULER_KW = dict(num_train_timesteps=1000, shift=1.0, stages=3,
                    stage_range=[0, 1 / 3, 2 / 3, 1], gamma=0.3333)  # run.py:27-31


def _g(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def step_inputs(dtype):
    g = _g(11)
    x = torch.randn(1, 38, 1, 12, 16, generator=g).to(dtype)
    v = torch.randn(1, 38, 1, 12, 16, generator=g).to(dtype)
    return x, v


# name -> config overrides, weight seed, clip dims (t, h, w) oldest first, batch, history, text lens
MMDIT_CASES = {
    "two_block_b2": dict(cfg=dict(num_layers=2), wseed=1, clips=[(1, 12, 16), (1, 12, 16)], B=2,
                         hist=None, lens=[1, 12], t=936.0639953613281, seed=21),
    "two_block_b3_hist": dict(cfg=dict(num_layers=2), wseed=1,
                              clips=[(2, 8, 8), (1, 16, 16), (1, 16, 16)], B=3, hist=(16, 16),
                              lens=[1, 12, 12], t=654.5895004272461, seed=22),
    "three_block_mixed": dict(cfg=dict(num_layers=3), wseed=5,
                              clips=[(3, 12, 16), (1, 24, 32), (1, 24, 32)], B=2, hist=None,
                              lens=[1, 20], t=97.53875732421875, seed=23),
}


def mmdit_inputs(case):
    g = _g(case["seed"])
    B = case["B"]
    clips = [torch.randn(B, 38, t, h, w, generator=g) for (t, h, w) in case["clips"]]
    enc = torch.randn(B, 77, 4096, generator=g)
    pooled = torch.randn(B, 2048, generator=g)
    mask = torch.zeros(B, 77, dtype=torch.long)
    for b, n in enumerate(case["lens"]):
        mask[b, :n] = 1
    t = torch.full((B,), case["t"], dtype=torch.float32)
    hist = hmask = None
    if case["hist"] is not None:
        hh, hw = case["hist"]
        hist = torch.randn(B, 38, 1, hh, hw, generator=g)
        ntok = (hh // 4) * (hw // 4)
        hmask = torch.cat([torch.zeros(B - 1, ntok), torch.ones(1, ntok)])  # [neg, ..., pos]
    return dict(clips=clips, enc=enc, mask=mask, pooled=pooled, t=t, hist=hist, hmask=hmask)


VAE_CASES = {
    # reduced widths keep the CPU reference fast; geometry (tiles, blends, windows) is the real one
    "tiled_2x1_T2": dict(cfg=dict(decoder_block_out_channels=(32, 32, 64, 64),
                                  encoder_block_out_channels=(32, 32, 64, 64),
                                  decoder_layers_per_block=(1, 1, 1, 1)), wseed=2,
                         latent=(1, 16, 2, 40, 16), seed=31),
    "untiled_T3": dict(cfg=dict(decoder_block_out_channels=(32, 32, 64, 64),
                                encoder_block_out_channels=(32, 32, 64, 64),
                                decoder_layers_per_block=(2, 1, 1, 1)), wseed=3,
                       latent=(1, 16, 3, 16, 16), seed=32),
    "tiled_2x2_T2": dict(cfg=dict(decoder_block_out_channels=(32, 32, 32, 32),
                                  encoder_block_out_channels=(32, 32, 32, 32),
                                  decoder_layers_per_block=(1, 1, 1, 1)), wseed=4,
                         latent=(1, 16, 2, 40, 40), seed=33),
}


def vae_latent(case):
    return torch.randn(*case["latent"], generator=_g(case["seed"]))


def vae_digest(y):
    """Compact, position-sensitive digest of a decoded video [1,3,T,H,W] (keeps fixtures small)."""
    return dict(shape=tuple(y.shape), sub=y[:, :, ::2, ::8, ::8].clone().float(),
                rows=y.double().mean(dim=(1, 2, 4)).float(), cols=y.double().mean(dim=(1, 2, 3)).float(),
                seam_h=y[:, :, :, 180:260:4, ::8].clone().float() if y.shape[3] > 260 else None,
                seam_w=y[:, :, :, ::8, 180:260:4].clone().float() if y.shape[4] > 260 else None)


# ---- VAE encode: moments of vae.encode(x) with tiling on (vae.py:844-883) ----
VAE_ENC_CASES = {
    # reduced widths keep the CPU reference fast; tile geometry (256-px tiles every 192 px, 8-latent
    # blends) and the stride-2 causal convolutions are the real ones
    "image_untiled": dict(cfg=dict(decoder_block_out_channels=(32, 32, 64, 64), encoder_block_out_channels=(32, 32, 64, 64),
                                   decoder_layers_per_block=(1, 1, 1, 1), encoder_layers_per_block=(1, 2, 1, 1)),
                          wseed=5, video=(1, 3, 1, 64, 96), seed=41),
    "clip9_tiled_2x2": dict(cfg=dict(decoder_block_out_channels=(32, 32, 64, 64), encoder_block_out_channels=(32, 32, 64, 64),
                                     decoder_layers_per_block=(1, 1, 1, 1), encoder_layers_per_block=(1, 1, 1, 1)),
                            wseed=6, video=(1, 3, 9, 320, 272), seed=42),
    "clip17_tiled_1x2": dict(cfg=dict(decoder_block_out_channels=(32, 32, 32, 32), encoder_block_out_channels=(32, 32, 32, 32),
                                      decoder_layers_per_block=(1, 1, 1, 1), encoder_layers_per_block=(2, 1, 1, 1)),
                             wseed=7, video=(1, 3, 17, 200, 328), seed=43),
}


def vae_video(case):
    return torch.randn(*case["video"], generator=_g(case["seed"]))
