"""The drop-in boundary exercised through the REFERENCE'S OWN CODE.

The unmodified `InferencePipeline` (pipeline.py, staged git-ignored under baseline/_ref/ by
oracle/reference_loader.stage(), imported through oracle/_shim.py) is subclassed exactly as
INTEGRATION.md §1 describes — only `_create_models` (pipeline.py:203-223) is overridden, returning
`B200MMDiT / B200VAE / B200Scheduler` — and its own `__init__`, `generate_one_unit`
(pipeline.py:439-524: stage loop, nearest up-sampling + re-noise in ATen, `torch.cat` CFG batches, the
bf16-rounded timestep, `autocast`, separate CFG arithmetic, `scheduler.step`) and `decode_latent`
(:703-725) run over those objects.  The result must equal `B200Pipeline`'s fused mirror of the same loop
BIT FOR BIT on the same noise: both call the same denoiser on the same inputs, and the fused
CFG + Euler / re-noise kernels reproduce ATen's rounding sequence.
"""
import pytest
import torch

from oracle import reference_loader as rl
from oracle import weights
from tests.golden import cases, rollout_cases as rc

pytestmark = pytest.mark.gpu

MODEL_CFG = dict(stages=[1, 2, 4], frame_per_unit=1, max_temporal_length=8, vae_downsample=8, raymap_dim=6,
                 history_guidance_scale=6.0, history_downsample_ratio=2)


@pytest.fixture(scope="module")
def both():
    if not rl.available():
        pytest.skip("reference sources not staged (oracle/reference_loader.stage() needs /root/reference)")
    from deepv_b200.mmdit import B200MMDiT
    from deepv_b200.pipeline import B200Pipeline
    from deepv_b200.scheduler import B200Scheduler
    from deepv_b200.vae import B200VAE
    dtype = torch.bfloat16
    cfg, W = weights.mmdit_weights(dict(num_layers=2), seed=1)
    vcfg, VW = weights.vae_weights(dict(decoder_block_out_channels=(128, 128, 128, 128),
                                        encoder_block_out_channels=(128, 128, 128, 128),
                                        decoder_layers_per_block=(1, 1, 1, 1)), seed=7)
    dit = B200MMDiT(W, cfg)
    vae = B200VAE(VW, vcfg, dtype=dtype)

    def as_plugin(base):
        class B200InferencePipeline(base):
            pass
        return B200InferencePipeline

    pl, ref_pipe = rl.build_pipeline((dit, vae, B200Scheduler(**cases.SCHEDULER_KW)), cases.SCHEDULER_KW, MODEL_CFG,
                                     rc.text_embeds(rc.ROLLOUT), device="cuda", dtype=dtype, subclass_hook=as_plugin)
    assert isinstance(ref_pipe, pl.InferencePipeline) and ref_pipe.model is dit and ref_pipe.vae is vae
    ours = B200Pipeline(dit, vae, B200Scheduler(**cases.SCHEDULER_KW), model_cfg=MODEL_CFG, torch_dtype=dtype)
    return ref_pipe, ours, dtype


@pytest.mark.parametrize("with_history", [False, True])
def test_reference_generate_one_unit_loop_over_b200_objects(both, with_history):
    ref_pipe, ours, dtype = both
    dev = "cuda"
    g = torch.Generator().manual_seed(61 + int(with_history))
    nb = 3 if with_history else 2
    h0, w0 = 8, 16
    lat = (torch.randn(1, 38, 1, h0, w0, generator=g) * 2).to(dtype).to(dev)
    conds = [[torch.randn(nb, 38, 2, h0, w0, generator=g).to(dtype).to(dev),
              torch.randn(nb, 38, 1, h0 * 2 ** i, w0 * 2 ** i, generator=g).to(dtype).to(dev)] for i in range(3)]
    noise = [torch.randn(1, 38, 1, h0 * 2, w0 * 2, generator=g), torch.randn(1, 38, 1, h0 * 4, w0 * 4, generator=g)]
    hist = torch.randn(1, 38, 1, h0 * 4, w0 * 4, generator=g).to(dtype).to(dev) if with_history else None
    enc = torch.randn(nb, 77, 4096, generator=g).to(dtype)
    pooled = torch.randn(nb, 2048, generator=g).to(dtype)
    mask = torch.zeros(nb, 77, dtype=torch.long)
    mask[0, :1] = 1
    mask[1:, :12] = 1
    steps = [2, 3, 2]

    # the reference's own loop; its CPU block-noise sampler (pipeline.py:431-437) replays the injected draws
    draws = iter(noise)
    ref_pipe.sample_block_noise = lambda bs, ch, t, h, w: next(draws)
    ref_pipe._guidance_scale, ref_pipe._video_guidance_scale = 4.0, 3.5      # pipeline.py:548-549
    got = ref_pipe.generate_one_unit(lat.clone(), hist, [[c.clone() for c in st] for st in conds], enc, mask, pooled,
                                     steps, h0, w0, 1, torch.device(dev), dtype, None, is_first_frame=False)
    want = ours.generate_one_unit(lat.clone(), hist, [[c.clone() for c in st] for st in conds], enc, mask, pooled,
                                  steps, block_noise=noise)
    torch.cuda.synchronize()
    assert len(got) == len(want) == 3
    for i in range(3):
        assert got[i].dtype == want[i].dtype == dtype and got[i].shape == want[i].shape
        assert torch.isfinite(got[i]).all()
        assert torch.equal(got[i], want[i]), f"stage {i}: max diff {(got[i].float() - want[i].float()).abs().max().item():.3e}"


def test_reference_decode_latent_over_b200_vae(both):
    ref_pipe, ours, dtype = both
    z = torch.randn(1, 16, 3, 40, 48, generator=torch.Generator().manual_seed(63)).to(dtype).cuda()
    got = ref_pipe.decode_latent(z.clone(), save_memory=True)        # pipeline.py:703-713 (mutates its argument)
    want = ours.decode_latent(z.clone())
    torch.cuda.synchronize()
    assert got.shape == want.shape == (1, 3, 17, 320, 384) and got.dtype == dtype
    assert torch.equal(got, want)
