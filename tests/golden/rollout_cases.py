"""Seeded tiny rollout shared by the rollout golden generator and the tests (CPU generators only).

One case = the inputs `InferencePipeline.generate` (pipeline.py:264-424) consumes: a first frame,
a list of action prompts, a table of prompt embeddings, plus the configuration of the three
models.  Geometry is the smallest the reference's rollout accepts that both paths can run:
192x256 px (latents 24x32 -> stages 6x8 / 12x16 / 24x32), 8 units = 57 frames per iteration,
2 iterations (12 prompts), one denoising step per stage.
"""
import torch

from tests.golden import cases

ROLLOUT = dict(
    height=192, width=256, n_prompts=12, seed=71,
    model_cfg=dict(stages=[1, 2, 4], frame_per_unit=1, max_temporal_length=8, vae_downsample=8, raymap_dim=6,
                   history_guidance_scale=6.0, history_downsample_ratio=2, num_inference_steps=1),
    dit=dict(cfg=dict(num_layers=2), wseed=1),
    vae=dict(cfg=dict(decoder_block_out_channels=(32, 32, 64, 64), encoder_block_out_channels=(32, 32, 64, 64),
                      decoder_layers_per_block=(1, 1, 1, 1), encoder_layers_per_block=(1, 1, 1, 1)), wseed=8),
)
# the same rollout for the CUDA path: GroupNorm groups of >= 4 channels (128-wide VAE), 128x128 px
ROLLOUT_GPU = dict(ROLLOUT, height=128, width=128, seed=81,
                   vae=dict(cfg=dict(decoder_block_out_channels=(128, 128, 128, 128),
                                     encoder_block_out_channels=(128, 128, 128, 128),
                                     decoder_layers_per_block=(1, 1, 1, 1), encoder_layers_per_block=(1, 1, 1, 1)),
                            wseed=9))
ACTIONS = ("w", "a", "s", "d")


def first_frame(case=ROLLOUT):
    """uint8 [H, W, 3]: smooth colour ramps plus seeded texture (what `batch_dict['img']` holds)."""
    g = cases._g(case["seed"])
    H, W = case["height"], case["width"]
    yy = torch.linspace(0, 1, H).view(H, 1, 1)
    xx = torch.linspace(0, 1, W).view(1, W, 1)
    base = torch.cat([yy.expand(H, W, 1), xx.expand(H, W, 1), (1 - yy * xx).expand(H, W, 1)], dim=2)
    img = (0.8 * base + 0.2 * torch.rand(H, W, 3, generator=g)).clamp(0, 1)
    return (img * 255).to(torch.uint8)


def prompts(case=ROLLOUT):
    g = cases._g(case["seed"] + 1)
    idx = torch.randint(0, len(ACTIONS), (case["n_prompts"],), generator=g).tolist()
    return [ACTIONS[i] for i in idx]


def text_embeds(case=ROLLOUT):
    """prompt -> {prompt_embeds [1,77,4096], pooled_prompt_embeds [1,2048], prompt_attention_mask [1,77]}
    (the layout of the file `model_cfg['text_embeds_path']` points to, pipeline.py:199)."""
    g = cases._g(case["seed"] + 2)
    out = {}
    for k, name in enumerate(("empty",) + ACTIONS):
        mask = torch.zeros(1, 77, dtype=torch.long)
        mask[0, :1 + 3 * k] = 1
        out[name] = dict(prompt_embeds=torch.randn(1, 77, 4096, generator=g),
                         pooled_prompt_embeds=torch.randn(1, 2048, generator=g),
                         prompt_attention_mask=mask)
    return out


class NoiseTape:
    """Every random draw of a rollout, in call order, from one seeded CPU generator.

    The golden generator patches the reference's three noise sources (`randn_tensor` in pipeline.py:428
    and vae.py:614, `sample_block_noise` pipeline.py:431-437) with this tape, and the oracle / CUDA
    rollouts take the same tape, so no noise has to be stored in the fixture.  `block` keeps the
    reference's distribution (cov (1+g) I - g 11^T per 2x2 block) but draws it as z @ chol^T.
    """

    def __init__(self, seed):
        self.g = cases._g(seed)
        self.calls = []

    def randn(self, shape):
        self.calls.append(("randn", tuple(shape)))
        return torch.randn(*shape, generator=self.g)

    def block(self, bs, ch, temp, height, width, gamma):
        self.calls.append(("block", (bs, ch, temp, height, width)))
        cov = torch.eye(4) * (1 + gamma) - torch.ones(4, 4) * gamma
        z = torch.randn(bs * ch * temp * (height // 2) * (width // 2), 4, generator=self.g)
        n = z @ torch.linalg.cholesky(cov).T
        n = n.view(bs, ch, temp, height // 2, width // 2, 2, 2).permute(0, 1, 2, 3, 5, 4, 6)
        return n.reshape(bs, ch, temp, height, width).contiguous()


def digest_frames(v):
    """Compact, position-sensitive digest of a video [1,C,T,H,W]."""
    return dict(shape=tuple(v.shape), sub=v[:, :, ::4, ::8, ::8].clone().float(),
                frame_mean=v.double().mean(dim=(1, 3, 4)).float(), frame_std=v.double().std(dim=(1, 3, 4)).float())


class RecordingTape(NoiseTape):
    """A tape that also keeps every draw, so a later run can replay a suffix of it (teacher forcing)."""

    def __init__(self, seed):
        super().__init__(seed)
        self.draws = []

    def randn(self, shape):
        t = super().randn(shape)
        self.draws.append(t)
        return t

    def block(self, bs, ch, temp, height, width, gamma):
        t = super().block(bs, ch, temp, height, width, gamma)
        self.draws.append(t)
        return t


class ReplayTape:
    """Replays recorded draws from position `start`, checking the requested shapes."""

    def __init__(self, draws, calls, start=0):
        self.draws, self.expect, self.pos = draws, calls, start
        self.calls = []

    def _next(self, kind, shape):
        want_kind, want_shape = self.expect[self.pos]
        assert (kind, tuple(shape)) == (want_kind, tuple(want_shape)), (self.pos, kind, shape, want_kind, want_shape)
        t = self.draws[self.pos]
        self.pos += 1
        self.calls.append((kind, tuple(shape)))
        return t

    def randn(self, shape):
        return self._next("randn", shape)

    def block(self, bs, ch, temp, height, width, gamma):
        return self._next("block", (bs, ch, temp, height, width))


def build_oracle_models(case=ROLLOUT):
    from oracle import rollout_ref, scheduler_ref, weights
    dcfg, DW = weights.mmdit_weights(case["dit"]["cfg"], seed=case["dit"]["wseed"])
    vcfg, VW = weights.vae_weights(case["vae"]["cfg"], seed=case["vae"]["wseed"], encoder=True)
    return rollout_ref.RolloutModels(dcfg, DW, vcfg, VW, scheduler_ref.pyramid_tables(**cases.SCHEDULER_KW),
                                     text_embeds(case), dict(case["model_cfg"]))
