"""Audio adapter ("vocal projector") of StableAvatar on the B200 C-ABI kernels.

Mirrors wan/models/vocal_projector_fantasy_1B.py:402-450 (FantasyTalkingVocalCondition1BModel), its 14B sibling
wan/models/vocal_projector_fantasy_14B.py:384-449 (FantasyTalkingVocalCondition14BModel: two-stage audio projection,
blocks at the DiT width = 8 heads of 640) and the window helpers of wan/models/vocal_projector_fantasy.py:39-131 — same
constructor arguments, parameter names and forward signature.
The adapter's residual stream is fp32 (nn.LayerNorm output under autocast, default dtype=torch.float32 in
VocalAttentionBlock.forward) while every Linear / attention runs in bf16.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


def split_audio_sequence(audio_proj_length, num_frames=81):
    """Index range of audio tokens centred on each latent frame (vocal_projector_fantasy.py:39-78)."""
    tokens_per_frame = audio_proj_length / num_frames
    half_tokens = int(tokens_per_frame * 4 / 2)
    centres = [0]
    for i in range(1, int((num_frames - 1) / 4) + 1):
        start_token = tokens_per_frame * ((i - 1) * 4 + 1)
        end_token = tokens_per_frame * (i * 4 + 1)
        centres.append(int((start_token + end_token) / 2) - 1)
    ranges = [[c - half_tokens, c + half_tokens] for c in centres]
    ranges[0] = [-(half_tokens * 2 - ranges[1][0]), ranges[1][0]]
    return ranges


def window_gather_table(length, pos_idx_ranges, expand_length=0):
    """Row-index table equivalent to split_tensor_with_padding (vocal_projector_fantasy.py:81-131): table[g][a] is the
    source token of slot a in window g, or -1 for the zero slots, which the reference appends at the END of a window
    whichever side ran out of range. Returns (table [G][A], lens [G])."""
    table, lens = [], []
    last = length - 1
    for start, end in pos_idx_ranges:
        start, end = start - expand_length, end + expand_length
        pad = max(-start, 0) + max(end - last, 0)
        vs, ve = max(start, 0), min(end, last)
        row = list(range(vs, ve + 1)) if vs <= ve else []
        lens.append(len(row))
        table.append(row + [-1] * pad)
    if len({len(r) for r in table}) != 1:
        raise RuntimeError(f"stack expects each tensor to be equal size, but got {[len(r) for r in table]}")
    return table, lens


def _param(*shape):
    return nn.Parameter(torch.empty(*shape), requires_grad=False)


class _Linear(nn.Module):
    def __init__(self, i, o, bias=True):
        super().__init__()
        self.weight = _param(o, i)
        self.bias = _param(o) if bias else None


class _Norm(nn.Module):
    def __init__(self, dim, bias=True):
        super().__init__()
        self.weight = _param(dim)
        if bias:
            self.bias = _param(dim)


class VocalCrossAttention(nn.Module):
    def __init__(self, vocal_dim, dit_dim, num_heads):
        super().__init__()
        self.num_heads, self.head_dim = num_heads, vocal_dim // num_heads
        self.q, self.k, self.v, self.o = (_Linear(vocal_dim, vocal_dim), _Linear(dit_dim, vocal_dim),
                                          _Linear(dit_dim, vocal_dim), _Linear(vocal_dim, vocal_dim))
        self.norm_q, self.norm_k = _Norm(vocal_dim, bias=False), _Norm(vocal_dim, bias=False)


class VocalAttentionBlock(nn.Module):
    def __init__(self, vocal_dim, dit_dim, ffn_dim, num_heads):
        super().__init__()
        self.norm3 = _Norm(vocal_dim)
        self.cross_attn = VocalCrossAttention(vocal_dim, dit_dim, num_heads)
        self.ffn = nn.Sequential(_Linear(vocal_dim, ffn_dim), nn.Identity(), _Linear(ffn_dim, vocal_dim))
        self.modulation = _param(1, 6, vocal_dim)


class Final_Head(nn.Module):
    def __init__(self, dim, out_dim):
        super().__init__()
        self.final_proj = _Linear(dim, out_dim)
        self.modulation = _param(1, 2, dim)


class VocalProjModel(nn.Module):
    def __init__(self, audio_in_dim, cross_attention_dim):
        super().__init__()
        self.proj = _Linear(audio_in_dim, cross_attention_dim, bias=False)
        self.norm = _Norm(cross_attention_dim)


class VocalProjModel14B(nn.Module):
    """vp14B.py:384-399: Linear(768, 2048, no bias) + LayerNorm, Linear(2048, dim, no bias) + LayerNorm."""

    def __init__(self, audio_in_dim, cross_attention_dim):
        super().__init__()
        self.proj_1 = _Linear(audio_in_dim, 2048, bias=False)
        self.norm_1 = _Norm(2048)
        self.proj_2 = _Linear(2048, cross_attention_dim, bias=False)
        self.norm_2 = _Norm(cross_attention_dim)


class FantasyTalkingVocalCondition1BModel(nn.Module):
    _proj_cls = VocalProjModel

    def __init__(self, audio_in_dim: int, audio_proj_dim: int, dit_dim: int):
        super().__init__()
        self.audio_in_dim, self.audio_proj_dim = audio_in_dim, audio_proj_dim
        self.proj_model = self._proj_cls(audio_in_dim, audio_proj_dim)
        self.blocks = nn.ModuleList([VocalAttentionBlock(audio_proj_dim, dit_dim, audio_proj_dim * 2, 8) for _ in range(2)])
        self.final_head = Final_Head(audio_proj_dim, audio_proj_dim)
        self._prep = None
        self._tables = {}

    def _prepare(self):
        if self._prep is None:
            blocks = []
            for b in self.blocks:
                ca = b.cross_attn
                blocks.append(dict(w_kv=torch.cat([ca.k.weight, ca.v.weight]).contiguous(),
                                   b_kv=torch.cat([ca.k.bias, ca.v.bias]).contiguous()))
            mods = torch.stack([b.modulation.reshape(-1) for b in self.blocks]).contiguous()
            self._prep = dict(blocks=blocks, mods=mods, head_mod=self.final_head.modulation.reshape(2, -1).contiguous())
        return self._prep

    def _project(self, a):
        """VocalProjModel (vp1B.py:389-399): bf16 Linear, then nn.LayerNorm whose autocast output is fp32."""
        pm = self.proj_model
        feat = ops.gemm(a, pm.proj.weight)
        return ops.layernorm(feat, weight=pm.norm.weight, bias=pm.norm.bias, eps=1e-5, out_dtype=torch.float32,
                             round_bf16=False)

    def _window_table(self, T, num_frames, device):
        key = (T, num_frames, str(device))
        if key not in self._tables:
            table, lens = window_gather_table(T, split_audio_sequence(T, num_frames), expand_length=4)
            self._tables[key] = (torch.tensor(table, dtype=torch.int32, device=device), len(table), len(table[0]),
                                 torch.tensor(lens, dtype=torch.long))
        return self._tables[key]

    def forward(self, vocal_embeddings=None, video_sample_n_frames=81, latents=None, e0=None, e=None):
        """vocal_embeddings [B,T,768] bf16, latents [B,L,dit_dim] bf16, e0 [B,6,dim] bf16, e [B,dim] bf16 ->
        (context tokens [B, G, A, C] bf16, window lengths [G]) as vp1B.py:433-450."""
        p = self._prepare()
        B, T, _ = vocal_embeddings.shape
        C = self.audio_proj_dim
        feat = self._project(vocal_embeddings.reshape(B * T, -1).to(torch.bfloat16))
        table, G, A, lens = self._window_table(T, video_sample_n_frames, feat.device)
        L = latents.shape[1]
        if L % G != 0:
            raise RuntimeError(f"shape '[{B * G}, -1, 8, {C // 8}]' is invalid for input of size {B * L * C}")
        outs = []
        for b in range(B):                                    # B == 1 on the pipeline path (1B.py:1004-1006)
            x = ops.gather_rows(feat[b * T:(b + 1) * T], table.reshape(-1))              # [G*A, C] fp32
            eb = ops.add_bcast(p["mods"], e0[b].reshape(1, -1).contiguous())             # [2, 1, 6C]
            lat = latents[b]
            for i, blk in enumerate(self.blocks):
                ch = eb[i, 0].view(6, C)
                ca, pb = blk.cross_attn, p["blocks"][i]
                ops.layernorm(x, shift=ch[0], scale=ch[1], gate=ch[2], res=x, out=x, round_bf16=False)
                xn = ops.layernorm(x, weight=blk.norm3.weight, bias=blk.norm3.bias)
                q = ops.gemm(xn, ca.q.weight, ca.q.bias)
                ops.rmsnorm_rope_(q, ca.norm_q.weight)
                kv = ops.gemm(lat, pb["w_kv"], pb["b_kv"])                               # [L, 2C]
                ops.rmsnorm_rope_(kv[:, :C], ca.norm_k.weight)
                nh, hd = ca.num_heads, ca.head_dim
                a = ops.attn_small_q(q.view(G, A, nh, hd), kv[:, :C].view(G, L // G, nh, hd),
                                     kv[:, C:].view(G, L // G, nh, hd))
                ops.gemm(a.view(G * A, C), ca.o.weight, ca.o.bias, res=x, out=x)
                xm = ops.layernorm(x, shift=ch[3], scale=ch[4], round_bf16=False)
                hdn = ops.gemm(xm, blk.ffn[0].weight, blk.ffn[0].bias, act=ops.ACT_GELU_TANH)
                ops.gemm(hdn, blk.ffn[2].weight, blk.ffn[2].bias, res=x, gate=ch[5], gate_ld=0, rows_per_batch=G * A, out=x)
            em = ops.add_bcast(p["head_mod"], e[b].reshape(1, -1).contiguous())          # [2, 1, C]
            xf = ops.layernorm(x, shift=em[0, 0], scale=em[1, 0], round_bf16=False)
            fh = self.final_head.final_proj
            outs.append(ops.gemm(xf, fh.weight, fh.bias).view(G, A, C))
        ctx = torch.stack(outs)
        if B > 1:
            lens = torch.cat([lens] * 3)
        return ctx, lens


class FantasyTalkingVocalCondition14BModel(FantasyTalkingVocalCondition1BModel):
    """vp14B.py:402-449. Same blocks / head / window logic as the 1B adapter; `audio_proj_dim` is the DiT width, so the
    8 attention heads are 640 wide (sa_attn_small_q tiles the keys), and the audio projection has two stages."""
    _proj_cls = VocalProjModel14B

    def _project(self, a):
        pm = self.proj_model
        f1 = ops.gemm(a, pm.proj_1.weight)
        f1 = ops.layernorm(f1, weight=pm.norm_1.weight, bias=pm.norm_1.bias, eps=1e-5, round_bf16=False)   # -> bf16 for proj_2
        f2 = ops.gemm(f1, pm.proj_2.weight)
        return ops.layernorm(f2, weight=pm.norm_2.weight, bias=pm.norm_2.bias, eps=1e-5, out_dtype=torch.float32,
                             round_bf16=False)
