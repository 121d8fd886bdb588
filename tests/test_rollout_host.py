"""Host-side logic of the device rollout (deepv_b200/rollout.py) that needs no GPU."""
from types import SimpleNamespace

import torch

from deepv_b200 import rollout
from oracle import rollout_ref
from tests.golden import rollout_cases as rc


def test_plan_prompts_matches_reference_arithmetic():
    """pipeline.py:275-279: padding of the prompt list and the iteration count."""
    for n in range(1, 30):
        p = [f"p{i}" for i in range(n)]
        assert rollout.plan_prompts(p, 8) == rollout_ref.plan_prompts(p, 8)
    assert rollout.plan_prompts(["w"] * 12, 8)[1] == 2
    assert rollout.plan_prompts(["w"] * 3, 8) == (["w"] * 8, 1)


def test_prompt_cache_lookup_and_stacking():
    table = rc.text_embeds()
    pc = rollout.PromptCache(table, None, "cpu")
    enc, mask, pooled = pc.branches("a", 2)
    assert torch.equal(enc, torch.cat([table["empty"]["prompt_embeds"], table["a"]["prompt_embeds"]]))
    assert torch.equal(mask, torch.cat([table["empty"]["prompt_attention_mask"], table["a"]["prompt_attention_mask"]]))
    assert torch.equal(pooled, torch.cat([table["empty"]["pooled_prompt_embeds"], table["a"]["pooled_prompt_embeds"]]))
    assert pc.branches("a", 2)[0] is enc                    # cached, not rebuilt
    assert pc.branches("a", 3)[0].shape[0] == 3
    try:
        pc.get("unknown text")
    except Exception as e:
        assert "no embedding" in str(e)
    else:
        raise AssertionError("expected a loud failure without a text encoder")


def test_key_frames_follow_the_global_stride():
    """pipeline.py:370-371 takes every 8th frame of the concatenated video; the device rollout keeps
    them incrementally (57 frames, then 32 more per iteration)."""
    fb = rollout._Feedback(SimpleNamespace(cfg=dict(vae_downsample=8)))
    vids = [torch.arange(57.0).view(1, 1, 57, 1, 1), 100 + torch.arange(32.0).view(1, 1, 32, 1, 1),
            200 + torch.arange(32.0).view(1, 1, 32, 1, 1)]
    for v in vids:
        fb._keep(v, v + 0.5)
        assert torch.equal(torch.cat(fb.key_images, dim=2), torch.cat(fb.images, dim=2)[:, :, ::8])
        assert torch.equal(torch.cat(fb.key_disps, dim=2), torch.cat(fb.disparitys, dim=2)[:, :, ::8])
    assert torch.cat(fb.key_images, dim=2).shape[2] == 8 + 4 + 4
