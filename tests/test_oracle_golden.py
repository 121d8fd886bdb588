"""CPU: the oracle restatement (and the host-side scheduler tables) against golden vectors
produced by the REAL reference (tests/golden/make_golden.py)."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import mmdit_ref, scheduler_ref, vae_ref, weights
from tests.golden import cases

G = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def sched_golden():
    return json.loads((G / "scheduler_golden.json").read_text())


def test_scheduler_tables_exact(sched_golden):
    tb = scheduler_ref.pyramid_tables(**cases.SCHEDULER_KW)
    for k in ("start_sigmas", "end_sigmas", "ori_start_sigmas"):
        assert {str(i): v for i, v in tb[k].items()} == sched_golden[k]
    assert {str(i): list(v) for i, v in tb["timestep_ratios"].items()} == sched_golden["timestep_ratios"]
    for key, ref in sched_golden["stages"].items():
        n, i = map(int, key.split("_"))
        ts, sg = scheduler_ref.stage_schedule(tb, n, i)
        assert ts.tolist() == ref["timesteps"] and sg.tolist() == ref["sigmas"]


def test_survey_appendix_d_values(sched_golden):
    # SURVEY.md App. D prints of the real reference
    assert sched_golden["start_sigmas"]["1"] == 0.8002459438271061
    assert sched_golden["stages"]["5_2"]["timesteps"][-1] == 1.385009765625
    a1, b1 = scheduler_ref.renoise_coefficients(0.6669999957084656, 0.3333)
    a2, b2 = scheduler_ref.renoise_coefficients(0.33399999141693115, 0.3333)
    assert (a1, b1) == (0.5998620228185144, 0.6930419797086197)
    assert (a2, b2) == (0.7496111148052451, 0.43367542844777596)


def test_host_scheduler_tables_match_golden(sched_golden):
    from deepv_b200.scheduler import B200Scheduler
    s = B200Scheduler(**cases.SCHEDULER_KW)
    assert {str(i): v for i, v in s.start_sigmas.items()} == sched_golden["start_sigmas"]
    assert {str(i): v for i, v in s.end_sigmas.items()} == sched_golden["end_sigmas"]
    assert {str(i): v for i, v in s.ori_start_sigmas.items()} == sched_golden["ori_start_sigmas"]
    assert {str(i): list(v) for i, v in s.timestep_ratios.items()} == sched_golden["timestep_ratios"]
    for key, ref in sched_golden["stages"].items():
        n, i = map(int, key.split("_"))
        s.set_timesteps(n, i)
        assert s.timesteps.tolist() == ref["timesteps"] and s.sigmas.tolist() == ref["sigmas"]
    with pytest.raises(ValueError):
        s.step(model_output=torch.zeros(1), timestep=3, sample=torch.zeros(1))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_euler_step_bit_exact(dtype, sched_golden):
    gold = torch.load(G / "scheduler_step_golden.pt")[str(dtype)]
    x, v = cases.step_inputs(dtype)
    sig = sched_golden["stages"]["5_0"]["sigmas"]
    cur = x
    for k in range(5):
        cur = scheduler_ref.euler_step(cur, v, sig[k], sig[k + 1])
        assert torch.equal(cur, gold[k]), f"step {k}"


@pytest.mark.parametrize("name", list(cases.MMDIT_CASES))
def test_mmdit_oracle_matches_reference_golden(name):
    gold = torch.load(G / "mmdit_golden.pt")[name]
    case = cases.MMDIT_CASES[name]
    cfg, W = weights.mmdit_weights(case["cfg"], seed=case["wseed"])
    inp = cases.mmdit_inputs(case)
    with torch.no_grad():
        y = mmdit_ref.mmdit_forward(W, cfg, inp["clips"], inp["t"], inp["enc"], inp["mask"],
                                    inp["pooled"], inp["hist"], inp["hmask"],
                                    2 if inp["hist"] is not None else None)
    assert y.shape == gold.shape
    assert (y - gold).abs().max().item() <= 2e-5 * max(1.0, gold.abs().max().item())


@pytest.mark.parametrize("name", list(cases.VAE_CASES))
def test_vae_oracle_matches_reference_golden(name):
    gold = torch.load(G / "vae_golden.pt")[name]
    case = cases.VAE_CASES[name]
    cfg, W = weights.vae_weights(case["cfg"], seed=case["wseed"])
    with torch.no_grad():
        y = vae_ref.tiled_decode(W, cfg, cases.vae_latent(case), 256, 1, True)
    d = cases.vae_digest(y)
    assert d["shape"] == gold["shape"]
    for k in ("sub", "rows", "cols", "seam_h", "seam_w"):
        if gold[k] is None:
            assert d[k] is None
            continue
        assert (d[k] - gold[k]).abs().max().item() <= 2e-5, k


@pytest.mark.parametrize("name", list(cases.VAE_ENC_CASES))
def test_vae_encode_oracle_matches_reference_golden(name):
    gold = torch.load(G / "vae_encode_golden.pt")[name]
    case = cases.VAE_ENC_CASES[name]
    cfg, W = weights.vae_weights(case["cfg"], seed=case["wseed"], encoder=True)
    with torch.no_grad():
        m = vae_ref.tiled_encode(W, cfg, cases.vae_video(case))
    assert m.shape == gold.shape
    assert (m - gold).abs().max().item() <= 2e-5
    # DiagonalGaussianDistribution.sample with the draw injected (vae.py:602-615)
    noise = torch.randn(m.shape[0], m.shape[1] // 2, *m.shape[2:], generator=torch.Generator().manual_seed(9))
    z = vae_ref.gaussian_sample(m, noise)
    mean, logvar = m.chunk(2, dim=1)
    assert torch.allclose(z, mean + torch.exp(0.5 * logvar.clamp(-30, 20)) * noise)


def test_cfg_combine_orders():
    p = torch.randn(3, 4, 5)
    u, t, h = p.chunk(3)
    assert torch.equal(scheduler_ref.cfg_combine(p, 3.5, 6.0), u + 3.5 * (t - u) + 6.0 * (h - t))
    assert torch.equal(scheduler_ref.cfg_combine(p[:2], 3.5, 6.0), u + 3.5 * (t - u))
    assert torch.equal(scheduler_ref.cfg_combine(p[:1], 3.5, 6.0), p[:1])
