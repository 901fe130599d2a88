"""Flow-matching Euler scheduler used by the inference pipeline.

The reference takes `FlowMatchEulerDiscreteScheduler` from diffusers==0.30.1 (pyproject.toml:15; call sites
inference.py:491-496, wan/pipeline/wan_inference_long_pipeline.py:645, 715, 754; config
deepspeed_config/wan2.1/wan_civitai.yaml: shift 5.0, use_dynamic_shifting false, 1000 train steps). diffusers is not
installed on the build or GPU boxes, so this is a restatement of that class's published arithmetic (same method
names and attributes the pipeline touches). PARITY UNPINNED: no reference test or vendored source pins it here.
The Euler update itself runs fused with the CFG combine in sa_cfg_euler_step (see pipeline.py).
"""
from __future__ import annotations

import numpy as np
import torch


class FlowMatchEulerDiscreteScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, shift: float = 1.0, use_dynamic_shifting: bool = False, **_):
        if use_dynamic_shifting:
            raise NotImplementedError("use_dynamic_shifting is false in the reference config (wan_civitai.yaml)")
        self.config = type("Config", (), dict(num_train_timesteps=num_train_timesteps, shift=shift,
                                              use_dynamic_shifting=use_dynamic_shifting))()
        timesteps = np.linspace(1, num_train_timesteps, num_train_timesteps, dtype=np.float32)[::-1].copy()
        sigmas = torch.from_numpy(timesteps) / num_train_timesteps
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        self.timesteps = sigmas * num_train_timesteps
        self.sigmas = sigmas.to("cpu")
        self.sigma_min, self.sigma_max = self.sigmas[-1].item(), self.sigmas[0].item()
        self._step_index = self._begin_index = None
        self.num_inference_steps = None

    @property
    def step_index(self):
        return self._step_index

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None, mu=None):
        """sigma_i = shift(linspace(sigma_max, sigma_min, N)) — the shift is applied a second time on top of the
        already shifted sigma_max / sigma_min, as diffusers 0.30.1 does — then a terminal 0 is appended."""
        if sigmas is None:
            self.num_inference_steps = num_inference_steps
            n = self.config.num_train_timesteps
            sigmas = np.linspace(self.sigma_max * n, self.sigma_min * n, num_inference_steps) / n
        s = self.config.shift
        sigmas = s * sigmas / (1 + (s - 1) * sigmas)
        sigmas = torch.from_numpy(np.asarray(sigmas)).to(dtype=torch.float32)
        self.timesteps = (sigmas * self.config.num_train_timesteps).to(device=device)
        self.sigmas = torch.cat([sigmas, torch.zeros(1)])          # kept on the host: dsigma is a launch argument
        self._timesteps_host = (sigmas * self.config.num_train_timesteps).tolist()
        self._step_index = self._begin_index = None

    def index_for_timestep(self, timestep):
        t = float(timestep)
        idx = [i for i, v in enumerate(self._timesteps_host) if v == t]
        if not idx:
            raise IndexError(f"timestep {t} is not on the schedule")
        return idx[1] if len(idx) > 1 else idx[0]

    def dsigma(self, timestep) -> float:
        """sigma_{i+1} - sigma_i as an fp32 value (the fp32 tensor subtraction of diffusers' step)."""
        i = self.index_for_timestep(timestep) if self._step_index is None else self._step_index
        self._step_index = i + 1
        return float((self.sigmas[i + 1] - self.sigmas[i]).item())

    def dsigma_at(self, i: int) -> float:
        """sigma_{i+1} - sigma_i for schedule position i, from the host copy of the table (no device sync)."""
        return float((self.sigmas[i + 1] - self.sigmas[i]).item())

    def step(self, model_output, timestep, sample, return_dict=False, **_):
        """prev = sample.float() + (sigma_next - sigma) * model_output, cast back to model_output.dtype."""
        from . import ops
        d = self.dsigma(timestep)
        if model_output.dtype != torch.bfloat16:
            raise NotImplementedError("scheduler.step: bf16 model output expected on the B200 path")
        out = ops.cfg_euler_step(model_output.contiguous(), sample.to(torch.bfloat16).contiguous(), d, cfg=False)
        return (out,)
