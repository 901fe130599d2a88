"""Golden fixture of the WHOLE inference pipeline from the REAL reference class
(wan/pipeline/wan_inference_long_pipeline.py: WanI2VTalkingInferenceLongPipeline.__call__, :540-806) run on CPU in fp32.

What is real: the pipeline class and its loop (sliding windows, audio slicing, 3-way CFG, overlap blending, mask / y
assembly, VAE encode of the conditioning clip, VAE decode, post-processing), the reference DiT (tiny stand-in width) and
the reference VAE. What is stubbed (none of it is on the hot path): the diffusers base classes the module imports
(`DiffusionPipeline` = attribute registry + cpu execution device), the T5 tokenizer / text encoder, CLIP image encoder
and Wav2Vec2 (deterministic functions of their inputs, defined in `pipeline_stubs` so the tests can build the same
ones), and `FlowMatchEulerDiscreteScheduler` — diffusers is not installed, so the scheduler object is this repo's
restatement with a torch `step` (its arithmetic stays "parity unpinned", see oracle/pipeline.py).

Run in the build container only:   python tools/gen_golden_pipeline.py   -> tests/golden/pipeline_tiny.npz
"""
from __future__ import annotations

import sys
import types
from contextlib import contextmanager
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import pipeline as OP  # noqa: E402
from oracle.refstub import REFERENCE_ROOT, import_reference  # noqa: E402
from stableavatar_b200 import synth  # noqa: E402
from stableavatar_b200.scheduler import FlowMatchEulerDiscreteScheduler as _Sched  # noqa: E402
from tools import pipeline_stubs as S  # noqa: E402


class TorchScheduler(_Sched):
    """The repo's scheduler restatement with diffusers' torch `step` (the product's step is a CUDA kernel)."""

    def step(self, model_output, timestep, sample, return_dict=False, **_):
        i = self.index_for_timestep(timestep) if self._step_index is None else self._step_index
        self._step_index = i + 1
        return (OP.euler_step(model_output, sample, self.sigmas[i], self.sigmas[i + 1]),)


def install_pipeline_stubs():
    """Call AFTER oracle.refstub.import_reference() (which (re)installs the model-level diffusers stubs)."""

    def mod(name, **attrs):
        m = sys.modules.get(name) or types.ModuleType(name)
        m.__dict__.update(attrs)
        m._sa_stub = True
        sys.modules[name] = m
        return m

    class DiffusionPipeline:
        def register_modules(self, **kw):
            for k, v in kw.items():
                setattr(self, k, v)

        @property
        def _execution_device(self):
            return torch.device("cpu")

        @contextmanager
        def progress_bar(self, total=None):
            yield types.SimpleNamespace(update=lambda *a: None)

        def maybe_free_model_hooks(self):
            pass

    class _Proc:
        def __init__(self, *a, **k):
            pass

    class BaseOutput:
        def __init__(self, **kw):
            self.__dict__.update(kw)

        def __init_subclass__(cls, **kw):
            pass

    import dataclasses
    import logging
    mod("diffusers", FlowMatchEulerDiscreteScheduler=TorchScheduler)
    mod("diffusers.callbacks", MultiPipelineCallbacks=type("MultiPipelineCallbacks", (), {}),
        PipelineCallback=type("PipelineCallback", (), {}))
    mod("diffusers.image_processor", VaeImageProcessor=_Proc)
    mod("diffusers.models.embeddings", get_1d_rotary_pos_embed=lambda *a, **k: None)
    mod("diffusers.pipelines")
    mod("diffusers.pipelines.pipeline_utils", DiffusionPipeline=DiffusionPipeline)
    mod("diffusers.schedulers", FlowMatchEulerDiscreteScheduler=TorchScheduler)
    du = sys.modules["diffusers.utils"]
    du.BaseOutput = dataclasses.dataclass(BaseOutput) if False else BaseOutput
    du.replace_example_docstring = lambda *_: (lambda f: f)
    if not hasattr(du, "logging"):
        du.logging = types.SimpleNamespace(get_logger=logging.getLogger)
    mod("diffusers.utils.torch_utils", randn_tensor=lambda shape, generator=None, device=None, dtype=None: torch.randn(
        shape, generator=generator, dtype=dtype))
    mod("diffusers.video_processor", VideoProcessor=_Proc)
    # wan.utils.__init__ pulls fm_solvers -> more diffusers; the pipeline only needs color_correction (skimage) by name
    pkg = mod("wan.utils")
    pkg.__path__ = [str(Path(REFERENCE_ROOT) / "wan" / "utils")]
    mod("wan.utils.color_correction", match_and_blend_colors=lambda *a, **k: None)
    mod("wan.models.wan_image_encoder", CLIPModel=type("CLIPModel", (), {}))
    mod("wan.models.wan_text_encoder", WanT5EncoderModel=type("WanT5EncoderModel", (), {}))


def gen_pipeline():
    dit, _, vae_mod = import_reference()
    install_pipeline_stubs()
    import wan.pipeline.wan_inference_long_pipeline as P

    cfg = synth.DIT_TINY
    keys = ("model_type", "patch_size", "text_len", "in_dim", "dim", "ffn_dim", "freq_dim", "text_dim", "out_dim",
            "num_heads", "num_layers")
    model = dit.WanTransformer3DFantasyModel(**{k: cfg[k] for k in keys}).eval()
    model.load_state_dict(synth.dit_state_dict(cfg), strict=True)
    vae = vae_mod.AutoencoderKLWan().eval()
    vae.load_state_dict(synth.vae_state_dict(encoder=True), strict=True)

    pipe = P.WanI2VTalkingInferenceLongPipeline(
        tokenizer=S.Tokenizer(), text_encoder=S.TextEncoder(cfg["text_dim"]), vae=vae, transformer=model,
        clip_image_encoder=S.ClipEncoder(), scheduler=TorchScheduler(1000, 5.0), wav2vec_processor=S.Wav2VecProcessor(),
        wav2vec=S.Wav2Vec())
    res = {}
    for name in ("windows3", "short_last"):
        c = S.case(name)
        S.write_cond_image(c["cond_path"], c["height"], c["width"])
        kw = dict(prompt=c["prompt"], negative_prompt=c["negative_prompt"], height=c["height"], width=c["width"],
                  num_frames=c["clip_length"], clip_length=c["clip_length"], num_inference_steps=c["steps"], guidance_scale=6.0,
                  text_guide_scale=c["text_scale"], audio_guide_scale=c["audio_scale"], vocal_input_values=c["audio"],
                  fps=c["fps"], sr=c["sr"], cond_file_path=c["cond_path"], overlap_window_length=c["overlap"],
                  overlapping_weight_scheme=c["scheme"])
        with torch.no_grad():
            video = pipe(latents=c["latents"].clone(), output_type="numpy", **kw).videos
            lat = pipe(latents=c["latents"].clone(), output_type="latent", return_dict=True, **kw).videos
        pre = "" if name == "windows3" else name + "_"
        if name == "windows3":                       # frames of one scenario are enough; latents pin the other
            res[pre + "video_f16"] = video.numpy().astype(np.float16)
        res[pre + "latents"] = lat.numpy().astype(np.float32)
        print(name, "video range", float(video.min()), float(video.max()))
    np.savez_compressed(ROOT / "tests" / "golden" / "pipeline_tiny.npz", **res)
    print("wrote pipeline_tiny.npz", {k: (v.shape, v.dtype) for k, v in res.items()})


if __name__ == "__main__":
    torch.set_grad_enabled(False)
    gen_pipeline()
