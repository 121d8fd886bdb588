"""The frame schedule of the trimmed decode, proved on the oracle (CPU): poison, in every causal conv of the decoder, the
output frames the schedule says need not be computed, and check that the kept output frames do not change by a bit.

`deepv_b200/work.decode_frame_schedule` mirrors the backward walk of `csrc/vae.cu:run_tile`; the GPU test
(`tests/test_gpu_fullsize.py::test_vae_trimmed_decode_is_bit_identical_on_the_kept_frames`) checks the kernels, this one
checks the receptive-field argument itself against the reference's decoder structure (vae.py:731-751, 225-252, 293-310).
"""
import re

import pytest
import torch

from deepv_b200 import synthetic, work
from oracle import vae_ref

SMALL = dict(decoder_block_out_channels=(32, 32, 64, 64), encoder_block_out_channels=(32, 32, 64, 64))


def first_frame_of(name, sched, keep):
    res_t, sp_t0, tp_t0, _ = sched
    m = re.match(r"decoder\.up_blocks\.(\d)\.resnets\.(\d)\.(conv1|conv2|conv_shortcut)$", name)
    if m:
        i, j, which = int(m.group(1)), int(m.group(2)), m.group(3)
        return max(0, res_t[i][j] - 2) if which == "conv1" else res_t[i][j]
    m = re.match(r"decoder\.up_blocks\.(\d)\.upsamplers\.0\.conv$", name)
    if m:
        return sp_t0[int(m.group(1))]
    m = re.match(r"decoder\.up_blocks\.(\d)\.temporal_upsamplers\.0\.conv$", name)
    if m:
        return tp_t0[int(m.group(1))]
    if name == "decoder.conv_out":
        return keep
    return 0          # post_quant_conv, conv_in, the mid block: all frames


@pytest.mark.parametrize("t_lat,keep", [(8, 25), (8, 1), (8, 56), (4, 9), (3, 12)])
def test_poisoned_skipped_frames_never_reach_the_kept_frames(monkeypatch, t_lat, keep):
    cfg, W = synthetic.vae_weights(SMALL, seed=5)
    z = torch.randn(1, cfg["decoder_in_channels"] if "decoder_in_channels" in cfg else 16, t_lat, 8, 16,
                    generator=torch.Generator().manual_seed(3))
    clean = vae_ref.full_decode(W, cfg, z)
    assert clean.shape[2] == 8 * (t_lat - 1) + 1 and keep < clean.shape[2]
    sched = work.decode_frame_schedule(keep, tuple(cfg["decoder_layers_per_block"]))
    real = vae_ref.causal_conv
    poisoned = []

    def conv(Wd, name, x, st, is_init):
        y = real(Wd, name, x, st, is_init)
        t0 = first_frame_of(name, sched, keep)
        if t0 > 0:
            y = y.clone()
            y[:, :, :t0] = float("nan")
            poisoned.append((name, t0, y.shape[2]))
        return y

    monkeypatch.setattr(vae_ref, "causal_conv", conv)
    trimmed = vae_ref.full_decode(W, cfg, z)
    assert torch.equal(trimmed[:, :, keep:], clean[:, :, keep:])            # bit-identical, no NaN leaked forward
    assert torch.isnan(trimmed[:, :, :keep]).all()
    if keep >= 25:
        # the schedule is not vacuous: the top level skips frames in every conv, the temporal up-sampler in front of it too
        names = {n for n, _, _ in poisoned}
        assert "decoder.up_blocks.3.resnets.0.conv1" in names and "decoder.up_blocks.2.temporal_upsamplers.0.conv" in names


def test_schedule_is_tight_at_the_top_level(monkeypatch):
    """Starting any top-level conv ONE frame later than scheduled corrupts a kept frame: the schedule skips as much as
    the causal look-back allows."""
    cfg, W = synthetic.vae_weights(SMALL, seed=5)
    z = torch.randn(1, 16, 8, 8, 16, generator=torch.Generator().manual_seed(3))
    clean = vae_ref.full_decode(W, cfg, z)
    keep = 25
    sched = work.decode_frame_schedule(keep, tuple(cfg["decoder_layers_per_block"]))
    real = vae_ref.causal_conv
    for victim in ("decoder.up_blocks.3.resnets.2.conv2", "decoder.up_blocks.3.resnets.0.conv1",
                   "decoder.up_blocks.2.temporal_upsamplers.0.conv"):
        def conv(Wd, name, x, st, is_init, victim=victim):
            y = real(Wd, name, x, st, is_init)
            if name == victim:
                y = y.clone()
                y[:, :, :first_frame_of(name, sched, keep) + 1] = float("nan")
            return y
        monkeypatch.setattr(vae_ref, "causal_conv", conv)
        out = vae_ref.full_decode(W, cfg, z)
        assert torch.isnan(out[:, :, keep:]).any(), victim
        monkeypatch.setattr(vae_ref, "causal_conv", real)
    assert not torch.isnan(clean).any()
