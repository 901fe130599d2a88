"""CPU checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol that
include/stableavatar_b200.h declares; argument validation answers without touching a GPU."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from stableavatar_b200 import build, _lib
    build.build()
    return _lib.lib()


def declared_symbols():
    text = (ROOT / "include" / "stableavatar_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sa_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported by libsa_b200.so"


def test_version_and_error_string(lib):
    assert lib.sa_version() >= 100
    assert isinstance(lib.sa_last_error(), bytes)


def test_bad_arguments_are_rejected_without_a_gpu(lib):
    from stableavatar_b200 import _lib as L
    g = L.GemmArgs()                       # all-null
    assert lib.sa_gemm_bf16(C.byref(g), None) == -1
    assert b"null" in lib.sa_last_error()
    g = L.GemmArgs(a=16, w=16, out=16, M=4, N=8, K=12, lda=12, ldw=12, ldc=8)
    assert lib.sa_gemm_bf16(C.byref(g), None) == -1   # K not a multiple of 8
    a = L.AttnArgs()
    assert lib.sa_flash_attn_d128(C.byref(a), None) == -1


def test_ops_refuse_cpu_tensors():
    import torch
    from stableavatar_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))


def test_model_state_dict_keys_match_reference_names():
    """The parameter names are part of the drop-in boundary: a reference checkpoint must load."""
    from stableavatar_b200 import synth
    from stableavatar_b200.wan_transformer3d import WanTransformer3DFantasyModel
    cfg = synth.DIT_TINY
    keys = ("model_type", "patch_size", "text_len", "in_dim", "dim", "ffn_dim", "freq_dim", "text_dim", "out_dim",
            "num_heads", "num_layers")
    m = WanTransformer3DFantasyModel(**{k: cfg[k] for k in keys})
    want = synth.dit_param_shapes(cfg)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == {k: tuple(v) for k, v in want.items()}
    assert m.config.patch_size == (1, 2, 2) and m.freqs.shape == (1024, 64) and m.freqs.dtype == torch.complex128


import torch  # noqa: E402
