"""CPU: the UNMODIFIED reference (imported through oracle/_shim.py from /root/reference, or from the staged
git-ignored copy under baseline/_ref/ where that is absent) against the oracle restatement, through
oracle/reference_loader.py — the loader the boundary test, `bench.py --impl reference` and the same-box
eager bar use.  Skipped only when neither location holds the reference sources."""
import pytest
import torch

from oracle import mmdit_ref, reference_loader as rl, scheduler_ref, weights
from tests.golden import cases, rollout_cases as rc

pytestmark = pytest.mark.skipif(not rl.available(), reason="reference sources not found (/root/reference or baseline/_ref)")

MODEL_CFG = dict(stages=[1, 2, 4], frame_per_unit=1, max_temporal_length=8, vae_downsample=8, raymap_dim=6,
                 history_guidance_scale=6.0, history_downsample_ratio=2)


def test_staging_copies_the_five_sources(tmp_path, monkeypatch):
    import os
    if not os.path.isdir("/root/reference"):
        pytest.skip("no mounted reference to stage from")
    monkeypatch.setattr(rl, "STAGED", tmp_path / "_ref")
    assert rl.stage()
    got = sorted(str(p.relative_to(tmp_path / "_ref")) for p in (tmp_path / "_ref").rglob("*.py"))
    assert got == sorted(rl.FILES)


def test_reference_pipeline_loop_equals_oracle_unit():
    """`InferencePipeline.generate_one_unit` (pipeline.py:439-524) with the real MMDiT / scheduler, built by the
    loader with only `_create_models` overridden, against scheduler_ref.generate_one_unit + mmdit_ref."""
    torch.set_grad_enabled(False)
    cfg, W = weights.mmdit_weights(dict(num_layers=2), seed=1)
    vcfg, VW = weights.vae_weights(dict(decoder_block_out_channels=(32, 32, 64, 64), encoder_block_out_channels=(32, 32, 64, 64),
                                        decoder_layers_per_block=(1, 1, 1, 1)), seed=7)
    dit, vae = rl.build_mmdit(cfg, W), rl.build_vae(vcfg, VW)
    pl, pipe = rl.build_pipeline((dit, vae, None), cases.SCHEDULER_KW, MODEL_CFG, rc.text_embeds(rc.ROLLOUT))
    assert type(pipe).__mro__[-2] is pl.InferencePipeline
    g = torch.Generator().manual_seed(61)
    nb, h0, w0 = 2, 8, 16
    lat = torch.randn(1, 38, 1, h0, w0, generator=g)
    conds = [[torch.randn(nb, 38, 1, h0 * 2 ** i, w0 * 2 ** i, generator=g)] for i in range(3)]
    noise = [torch.randn(1, 38, 1, h0 * 2, w0 * 2, generator=g), torch.randn(1, 38, 1, h0 * 4, w0 * 4, generator=g)]
    enc, pooled = torch.randn(nb, 77, 4096, generator=g), torch.randn(nb, 2048, generator=g)
    mask = torch.zeros(nb, 77, dtype=torch.long)
    mask[0, :1] = 1
    mask[1:, :12] = 1
    draws = iter(noise)
    pipe.sample_block_noise = lambda bs, ch, t, h, w: next(draws)
    pipe._guidance_scale, pipe._video_guidance_scale = 4.0, 3.5
    got = pipe.generate_one_unit(lat.clone(), None, conds, enc, mask, pooled, [1, 2, 1], h0, w0, 1, torch.device("cpu"),
                                 torch.float32, None, is_first_frame=False)
    tb = scheduler_ref.pyramid_tables(**cases.SCHEDULER_KW)

    def model_fn(clips, tt):
        return mmdit_ref.mmdit_forward(W, cfg, clips, tt.float(), enc, mask, pooled)

    want = scheduler_ref.generate_one_unit(model_fn, tb, lat, conds, noise, 2, [1, 2, 1], 3.5, 6.0)
    for a, b in zip(got, want):
        assert a.shape == b.shape
        assert (a - b).abs().max().item() <= 2e-5 * b.abs().max().item()
