"""ORACLE SUPPORT (test infrastructure) — import the real reference modules from /root/reference in the build
container. `diffusers` is not installed there, so the handful of symbols the model files import
(wan/models/wan_fantasy_transformer3d_1B.py:16-19, wan/models/wan_vae.py:8-14) are stubbed.
Used only by tools/gen_golden.py (which writes tests/golden/) — /root/reference does not exist on the GPU box.
"""
from __future__ import annotations

import functools
import inspect
import logging
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = "/root/reference"


class _Config(dict):
    __getattr__ = dict.__getitem__


def _register_to_config(init):
    @functools.wraps(init)
    def wrapper(self, *args, **kwargs):
        sig = inspect.signature(init)
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = _Config({k: v for k, v in bound.arguments.items() if k != "self"})
        init(self, *args, **kwargs)
        object.__setattr__(self, "_config", cfg)
    return wrapper


class _ConfigMixin:
    @property
    def config(self):
        return self._config

    @classmethod
    def from_config(cls, cfg, **kw):
        keys = inspect.signature(cls.__init__).parameters
        return cls(**{k: v for k, v in {**cfg, **kw}.items() if k in keys})


class _ModelMixin(nn.Module):
    _keys_to_ignore_on_load_unexpected = None
    _supports_gradient_checkpointing = False

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device


class _DecoderOutput:
    def __init__(self, sample):
        self.sample = sample


class _DiagGauss:
    def __init__(self, p):
        self.mean, self.logvar = torch.chunk(p, 2, dim=1)

    def mode(self):
        return self.mean


class _AEOutput:
    def __init__(self, latent_dist):
        self.latent_dist = latent_dist

    def __getitem__(self, i):                      # diffusers' BaseOutput indexes like a tuple of its fields
        return (self.latent_dist,)[i]


def install_stubs():
    if "diffusers" in sys.modules and not getattr(sys.modules["diffusers"], "_sa_stub", False):
        return

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m._sa_stub = True
        sys.modules[name] = m
        return m

    mod("diffusers")
    mod("diffusers.configuration_utils", ConfigMixin=_ConfigMixin, register_to_config=_register_to_config)
    mod("diffusers.models")
    mod("diffusers.models.modeling_utils", ModelMixin=_ModelMixin)
    mod("diffusers.loaders")
    mod("diffusers.loaders.single_file_model", FromOriginalModelMixin=type("FromOriginalModelMixin", (), {}))
    lg = types.SimpleNamespace(get_logger=lambda name=None: logging.getLogger(name or "ref"))
    mod("diffusers.utils", is_torch_version=lambda *a, **k: True, logging=lg)
    mod("diffusers.utils.accelerate_utils", apply_forward_hook=lambda f: f)
    mod("diffusers.models.autoencoders")
    mod("diffusers.models.autoencoders.vae", DecoderOutput=_DecoderOutput, DiagonalGaussianDistribution=_DiagGauss)
    mod("diffusers.models.modeling_outputs", AutoencoderKLOutput=_AEOutput)


def import_reference():
    """Returns (dit_module, adapter_module, vae_module) of the real reference with SDPA forced in the adapter
    (the adapter would otherwise call flash_attn_varlen_func, which asserts CUDA)."""
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import wan.models.wan_fantasy_transformer3d_1B as dit
    import wan.models.vocal_projector_fantasy_1B as vp
    import wan.models.wan_vae as vae
    vp.FLASH_ATTN_2_AVAILABLE = False
    vp.FLASH_ATTN_3_AVAILABLE = False
    return dit, vp, vae


def import_reference_14b():
    """The train_14B model module (wan/models/wan_fantasy_transformer3d_14B.py) with SDPA forced in its adapter."""
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import wan.models.wan_fantasy_transformer3d_14B as dit14
    import wan.models.vocal_projector_fantasy_14B as vp14
    vp14.FLASH_ATTN_2_AVAILABLE = False
    vp14.FLASH_ATTN_3_AVAILABLE = False
    dit14.FLASH_ATTN_2_AVAILABLE = False
    dit14.FLASH_ATTN_3_AVAILABLE = False
    return dit14, vp14
