"""TEST INFRASTRUCTURE ONLY — import shim that lets the *unmodified* reference files
(/root/reference/model/*.py, pipeline.py) import in this container, where diffusers / timm /
IPython / fire / imageio are not installed.

It is used by tests/golden/make_golden.py (to generate golden vectors from the real reference)
and by tests/test_oracle_vs_reference.py (to pin oracle/ against the real reference); both are
skipped when /root/reference is absent (e.g. on the GPU box).  Nothing in deepv_b200/ imports it.

The stand-ins restate the handful of diffusers 0.31.0 / timm 1.0.11 symbols the reference uses:
  GELU(approximate="tanh") = F.gelu(Linear(x), approximate="tanh"); get_activation("silu"/"swish")
  = nn.SiLU; Attention(_from_deprecated_attn_block, AttnProcessor2_0) = per-image GroupNorm ->
  q,k,v Linear -> SDPA (heads = C // dim_head) -> out Linear -> + residual -> / rescale;
  ConfigMixin/register_to_config = config dict with attribute access + attribute forwarding.
"""
from __future__ import annotations

import functools
import importlib
import importlib.machinery
import inspect
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def _find_reference() -> str:
    """$DEEPV_REFERENCE_ROOT, else the mounted reference, else the git-ignored copy that
    oracle/reference_loader.stage() puts under baseline/_ref/ so that it travels to the GPU box."""
    env = os.environ.get("DEEPV_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", _STAGED):
        if os.path.isfile(os.path.join(cand, "model", "mmdit.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_reference()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "mmdit.py"))


def _mod(name: str) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    m.__path__ = []  # behave like a package
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(sys.modules[parent], child, m)
    return m


def install() -> None:
    """Install the stand-in modules (idempotent)."""
    if "diffusers" in sys.modules and getattr(sys.modules["diffusers"], "_deepv_shim", False):
        return
    import transformers  # noqa: F401  (must be imported before a fake `timm` appears)
    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    class _Config(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError as e:
                raise AttributeError(k) from e

    def register_to_config(init):
        @functools.wraps(init)
        def wrapper(self, *args, **kwargs):
            sig = inspect.signature(init)
            bound = sig.bind(self, *args, **kwargs)
            bound.apply_defaults()
            cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
            object.__setattr__(self, "_internal_config", _Config(cfg))
            init(self, *args, **kwargs)
        return wrapper

    class ConfigMixin:
        @property
        def config(self):
            return self._internal_config

        def __getattr__(self, name):
            # attribute forwarding of config keys (pipeline.py:551 reads model.in_channels)
            cfg = self.__dict__.get("_internal_config")
            if cfg is not None and name in cfg:
                return cfg[name]
            sup = super()
            if hasattr(sup, "__getattr__"):
                return sup.__getattr__(name)
            raise AttributeError(name)

    class ModelMixin(nn.Module):
        @property
        def device(self):
            return next(self.parameters()).device

        @property
        def dtype(self):
            return next(self.parameters()).dtype

    class SchedulerMixin:
        pass

    class BaseOutput(dict):
        def __init__(self, **kw):
            super().__init__(**kw)
            for k, v in kw.items():
                object.__setattr__(self, k, v)

        def __post_init__(self):
            pass

    # dataclass-decorated subclasses of BaseOutput just need attribute storage
    class _DataOutput:
        pass

    def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
        return torch.randn(shape, generator=generator, device=device, dtype=dtype)

    def is_torch_version(op, ver):
        from packaging import version
        import operator
        ops = {">=": operator.ge, ">": operator.gt, "<": operator.lt, "<=": operator.le, "==": operator.eq}
        return ops[op](version.parse(torch.__version__.split("+")[0]), version.parse(ver))

    class _Logging:
        @staticmethod
        def get_logger(name):
            import logging
            return logging.getLogger(name)

    def get_activation(name):
        name = name.lower()
        if name in ("silu", "swish"):
            return nn.SiLU()
        if name == "gelu":
            return nn.GELU()
        if name == "relu":
            return nn.ReLU()
        raise ValueError(name)

    class GELU(nn.Module):
        def __init__(self, dim_in, dim_out, approximate="none", bias=True):
            super().__init__()
            self.proj = nn.Linear(dim_in, dim_out, bias=bias)
            self.approximate = approximate

        def forward(self, x):
            return F.gelu(self.proj(x), approximate=self.approximate)

    class GEGLU(nn.Module):
        def __init__(self, dim_in, dim_out, bias=True):
            super().__init__()
            self.proj = nn.Linear(dim_in, dim_out * 2, bias=bias)

        def forward(self, x):
            h, g = self.proj(x).chunk(2, dim=-1)
            return h * F.gelu(g)

    class ApproximateGELU(nn.Module):
        def __init__(self, dim_in, dim_out, bias=True):
            super().__init__()
            self.proj = nn.Linear(dim_in, dim_out, bias=bias)

        def forward(self, x):
            x = self.proj(x)
            return x * torch.sigmoid(1.702 * x)

    class Attention(nn.Module):
        """diffusers 0.31.0 Attention as used by the VAE mid block (vae.py:439-445)."""

        def __init__(self, query_dim, heads=8, dim_head=64, rescale_output_factor=1.0, eps=1e-5,
                     norm_num_groups=None, spatial_norm_dim=None, residual_connection=False,
                     bias=False, upcast_softmax=False, _from_deprecated_attn_block=False, **kw):
            super().__init__()
            inner = heads * dim_head
            self.heads = heads
            self.rescale_output_factor = rescale_output_factor
            self.residual_connection = residual_connection
            self.group_norm = (nn.GroupNorm(norm_num_groups, query_dim, eps=eps, affine=True)
                               if norm_num_groups is not None else None)
            self.to_q = nn.Linear(query_dim, inner, bias=bias)
            self.to_k = nn.Linear(query_dim, inner, bias=bias)
            self.to_v = nn.Linear(query_dim, inner, bias=bias)
            self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=True), nn.Dropout(0.0)])

        def forward(self, hidden_states, temb=None, **kw):
            residual = hidden_states
            b, c, h, w = hidden_states.shape
            x = hidden_states.view(b, c, h * w).transpose(1, 2)
            if self.group_norm is not None:
                x = self.group_norm(x.transpose(1, 2)).transpose(1, 2)
            q, k, v = self.to_q(x), self.to_k(x), self.to_v(x)
            hd = q.shape[-1] // self.heads
            q = q.view(b, -1, self.heads, hd).transpose(1, 2)
            k = k.view(b, -1, self.heads, hd).transpose(1, 2)
            v = v.view(b, -1, self.heads, hd).transpose(1, 2)
            x = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)
            x = x.transpose(1, 2).reshape(b, -1, self.heads * hd).to(q.dtype)
            x = self.to_out[1](self.to_out[0](x))
            x = x.transpose(-1, -2).reshape(b, c, h, w)
            if self.residual_connection:
                x = x + residual
            return x / self.rescale_output_factor

    class AutoencoderKLOutput:
        def __init__(self, latent_dist):
            self.latent_dist = latent_dist

    d = _mod("diffusers")
    d._deepv_shim = True
    _mod("diffusers.utils").is_torch_version = is_torch_version
    sys.modules["diffusers.utils"].BaseOutput = BaseOutput
    sys.modules["diffusers.utils"].logging = _Logging
    _mod("diffusers.utils.torch_utils").randn_tensor = randn_tensor
    _mod("diffusers.models")
    _mod("diffusers.models.modeling_utils").ModelMixin = ModelMixin
    cu = _mod("diffusers.configuration_utils")
    cu.ConfigMixin = ConfigMixin
    cu.register_to_config = register_to_config
    act = _mod("diffusers.models.activations")
    act.get_activation = get_activation
    act.GELU, act.GEGLU, act.ApproximateGELU = GELU, GEGLU, ApproximateGELU
    ap = _mod("diffusers.models.attention_processor")
    ap.Attention = Attention
    ap.AttentionProcessor = object
    _mod("diffusers.models.modeling_outputs").AutoencoderKLOutput = AutoencoderKLOutput
    _mod("diffusers.schedulers")
    _mod("diffusers.schedulers.scheduling_utils").SchedulerMixin = SchedulerMixin

    _mod("timm")
    _mod("timm.models")
    _mod("timm.models.layers").trunc_normal_ = lambda w, std=0.02, **kw: nn.init.trunc_normal_(
        w, std=std, a=-2.0, b=2.0)
    _mod("IPython").embed = lambda *a, **k: None
    for name in ("fire", "imageio", "plyfile"):
        if name not in sys.modules:
            _mod(name)


def import_reference():
    """Return (mmdit_module, scheduler_module, vae_module) of the real reference."""
    if not reference_available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT}")
    install()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    mm = importlib.import_module("model.mmdit")
    sc = importlib.import_module("model.scheduler")
    va = importlib.import_module("model.vae")
    return mm, sc, va
