"""fp8 weight mode — the numerics of wan/utils/fp8_optimization.py on the B200 path.

The reference's `model_cpu_offload_and_qfloat8` mode (inference.py:517-520) stores every parameter except the
`modulation` tables as float8_e4m3fn and casts each module back to bf16 around its forward (`convert_weight_dtype_wrapper`
-> `autocast_model_forward`, fp8_optimization.py:7-17, 44-56): the arithmetic is ordinary bf16 on weights that were
rounded to e4m3. That rounding is what changes results; the fp8 storage only saves memory, which a 180 GB part does not
need. So here `convert_model_weight_to_float8` rounds the parameters through float8_e4m3fn IN PLACE and leaves them in
the model's dtype (the tcgen05 GEMMs keep reading bf16), and `convert_weight_dtype_wrapper` has nothing left to do.
Same function names and arguments as the reference, so inference.py's calls work unchanged.
"""
from __future__ import annotations

import torch


def replace_parameters_by_name(module, name_keywords, device):
    """fp8_optimization.py:19-28: turn the named nn.Parameters into plain tensors on `device`."""
    from torch import nn
    for name, param in list(module.named_parameters(recurse=False)):
        if any(keyword in name for keyword in name_keywords):
            if isinstance(param, nn.Parameter):
                tensor = param.data
                delattr(module, name)
                setattr(module, name, tensor.to(device=device))
    for _, child in module.named_children():
        replace_parameters_by_name(child, name_keywords, device)


@torch.no_grad()
def convert_model_weight_to_float8(model, exclude_module_name=("embed_tokens",)):
    """fp8_optimization.py:30-45 — same selection rule (module name or parameter name containing an excluded keyword is
    skipped); the selected parameters take the values float8_e4m3fn storage would give them."""
    for name, module in model.named_modules():
        if any(k in name for k in exclude_module_name):
            continue
        for param_name, param in module.named_parameters():
            if any(k in param_name for k in exclude_module_name):
                continue
            param.data = param.data.to(torch.float8_e4m3fn).to(param.dtype)
    for m in model.modules():                      # concatenated / re-laid operands are rebuilt from the new values
        if hasattr(m, "_prep"):
            m._prep = None
    model._fp8_weight_values = True
    return model


def convert_weight_dtype_wrapper(module, origin_dtype):
    """fp8_optimization.py:47-56 wraps every forward in an fp8 -> origin_dtype -> fp8 round trip; the parameters here are
    already held in `origin_dtype`, so only the dtype is checked."""
    for p in module.parameters():
        if p.dtype not in (origin_dtype, torch.float32):
            raise TypeError(f"parameter of dtype {p.dtype}: expected {origin_dtype} (call convert_model_weight_to_float8 "
                            "from this module, which keeps the storage dtype)")
    return module
