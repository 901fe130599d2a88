"""ORACLE (test infrastructure, not product code) — CPU restatement of StableAvatar's audio-conditioned Wan2.1 DiT.

Plain PyTorch fp32 on CPU (fp64 where the reference uses it), written functionally over a reference-named state
dict. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Pinned against outputs of the real reference modules (tests/golden/*.npz, made by tools/gen_golden.py by importing
/root/reference in the build container): see tests/test_oracle_golden.py.

Every function cites the reference lines it restates (paths relative to /root/reference):
  1B  = wan/models/wan_fantasy_transformer3d_1B.py
  vp1B = wan/models/vocal_projector_fantasy_1B.py
  vp  = wan/models/vocal_projector_fantasy.py
  tc  = wan/models/cache_utils.py
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------ small pieces
def sinusoidal_embedding_1d(dim, position):
    """1B:210-220 — [cos | sin] of position * 10000^(-i/half), computed in float64."""
    half = dim // 2
    position = position.to(torch.float64)
    sinusoid = torch.outer(position, torch.pow(10000, -torch.arange(half).to(position).div(half)))
    return torch.cat([torch.cos(sinusoid), torch.sin(sinusoid)], dim=1)


def rope_params(max_seq_len, dim, theta=10000):
    """1B:224-231 — complex128 table exp(i * pos * theta^(-2j/dim))."""
    freqs = torch.outer(torch.arange(max_seq_len),
                        1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float64).div(dim)))
    return torch.polar(torch.ones_like(freqs), freqs)


def rope_freqs(head_dim):
    """1B:855-862 — 64 complex pairs per head split 22 (frame) / 21 (row) / 21 (col) for head_dim 128."""
    d = head_dim
    return torch.cat([rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)),
                      rope_params(1024, 2 * (d // 6))], dim=1)


def rope_apply(x, grid_sizes, freqs):
    """1B:296-323 — x [B, L, N, D]; adjacent pairs (2j, 2j+1) rotated; tokens past f*h*w are left untouched."""
    n, c = x.size(2), x.size(3) // 2
    fr = freqs.split([c - 2 * (c // 3), c // 3, c // 3], dim=1)
    out = []
    for i, (f, h, w) in enumerate(grid_sizes):
        seq_len = f * h * w
        x_i = torch.view_as_complex(x[i, :seq_len].to(torch.float32).reshape(seq_len, n, -1, 2))
        freqs_i = torch.cat([fr[0][:f].view(f, 1, 1, -1).expand(f, h, w, -1),
                             fr[1][:h].view(1, h, 1, -1).expand(f, h, w, -1),
                             fr[2][:w].view(1, 1, w, -1).expand(f, h, w, -1)], dim=-1).reshape(seq_len, 1, -1)
        x_i = torch.view_as_real(x_i * freqs_i).flatten(2)
        out.append(torch.cat([x_i, x[i, seq_len:]]))
    return torch.stack(out).float()


def rms_norm(x, weight, eps=1e-6):
    """1B:326-342 — over the full channel dim (all heads jointly)."""
    xf = x.float()
    return (xf * torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)).type_as(x) * weight


def layer_norm(x, weight=None, bias=None, eps=1e-6):
    """1B:345-355."""
    return F.layer_norm(x.float(), (x.shape[-1],), weight, bias, eps).type_as(x)


def attention(q, k, v):
    """1B:158-207, SDPA branch: softmax(q k^T / sqrt(d)) v, no mask (k_lens ignored), layout [B, L, N, D]."""
    q, k, v = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(q.shape[-1])
    return torch.matmul(torch.softmax(s, dim=-1), v).transpose(1, 2).contiguous()


def linear(x, sd, name):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


# ------------------------------------------------------------------------------------------------ audio windows
def split_audio_sequence(audio_proj_length, num_frames=81):
    """vp:39-78."""
    tokens_per_frame = audio_proj_length / num_frames
    half_tokens = int(tokens_per_frame * 4 / 2)
    pos = []
    for i in range(int((num_frames - 1) / 4) + 1):
        if i == 0:
            pos.append(0)
        else:
            start_token = tokens_per_frame * ((i - 1) * 4 + 1)
            end_token = tokens_per_frame * (i * 4 + 1)
            pos.append(int((start_token + end_token) / 2) - 1)
    ranges = [[p - half_tokens, p + half_tokens] for p in pos]
    ranges[0] = [-(half_tokens * 2 - ranges[1][0]), ranges[1][0]]
    return ranges


def split_tensor_with_padding(x, pos_idx_ranges, expand_length=0):
    """vp:81-131 — windows gathered from x [1, T, C]; out-of-range slots become zeros appended at the END."""
    ranges = [[a - expand_length, b + expand_length] for a, b in pos_idx_ranges]
    max_valid = x.size(1) - 1
    subs, lens = [], []
    for start, end in ranges:
        pad = max(-start, 0) + max(end - max_valid, 0)
        vs, ve = max(start, 0), min(end, max_valid)
        part = x[:, vs:ve + 1] if vs <= ve else x.new_zeros((1, 0, x.size(2)))
        sub = F.pad(part, (0, 0, 0, pad, 0, 0))
        lens.append(sub.size(-2) - pad)
        subs.append(sub)
    return torch.stack(subs, dim=1), torch.tensor(lens, dtype=torch.long)


# ------------------------------------------------------------------------------------------------ audio adapter
def vocal_block(sd, pre, x, e0, latents, G, num_heads=8):
    """vp1B:338-362 (+ VocalCrossAttention vp1B:245-277): x [B, G, A, C] or [B, G*A, C]."""
    e = (sd[pre + "modulation"] + e0).chunk(6, dim=1)
    if x.dim() == 4:
        x = x.flatten(1, 2)
    temp = layer_norm(x) * (1 + e[1]) + e[0]
    x = x + temp * e[2]                                    # "pseudo self-attention": no attention at all
    b, C = x.size(0), x.size(2)
    d = C // num_heads
    xn = layer_norm(x, sd[pre + "norm3.weight"], sd[pre + "norm3.bias"])
    q = rms_norm(linear(xn, sd, pre + "cross_attn.q"), sd[pre + "cross_attn.norm_q.weight"]).view(b * G, -1, num_heads, d)
    k = rms_norm(linear(latents, sd, pre + "cross_attn.k"), sd[pre + "cross_attn.norm_k.weight"]).view(b * G, -1, num_heads, d)
    v = linear(latents, sd, pre + "cross_attn.v").view(b * G, -1, num_heads, d)
    a = attention(q, k, v).view(b, -1, num_heads, d).flatten(2)
    x = x + linear(a, sd, pre + "cross_attn.o")
    temp = layer_norm(x) * (1 + e[4]) + e[3]
    y = linear(F.gelu(linear(temp, sd, pre + "ffn.0"), approximate="tanh"), sd, pre + "ffn.2")
    return x + y * e[5]


def vocal_projector(sd, vocal_embeddings, video_sample_n_frames, latents, e0, e, pre="vocal_projector."):
    """vp1B:433-450 — returns ([B, G, A, C], lens [G]). A state dict with proj_model.proj_1 is the 14B adapter
    (vp14B:384-399: two Linear+LayerNorm stages 768 -> 2048 -> dim; blocks and head identical, vp14B:431-449)."""
    if pre + "proj_model.proj_1.weight" in sd:
        feat = vocal_embeddings
        for n in ("1", "2"):
            feat = linear(feat, sd, f"{pre}proj_model.proj_{n}")
            feat = F.layer_norm(feat, (feat.shape[-1],), sd[f"{pre}proj_model.norm_{n}.weight"],
                                sd[f"{pre}proj_model.norm_{n}.bias"], 1e-5)
    else:
        feat = linear(vocal_embeddings, sd, pre + "proj_model.proj")
        feat = F.layer_norm(feat, (feat.shape[-1],), sd[pre + "proj_model.norm.weight"], sd[pre + "proj_model.norm.bias"], 1e-5)
    ranges = split_audio_sequence(feat.size(1), num_frames=video_sample_n_frames)
    x, lens = split_tensor_with_padding(feat, ranges, expand_length=4)
    G = x.size(1)
    n_blocks = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith(pre + "blocks."))
    for i in range(n_blocks):
        x = vocal_block(sd, f"{pre}blocks.{i}.", x, e0, latents, G)
    em = (sd[pre + "final_head.modulation"] + e.unsqueeze(1)).chunk(2, dim=1)
    x = linear(layer_norm(x) * (1 + em[1]) + em[0], sd, pre + "final_head.final_proj")
    return x.view(x.size(0), G, -1, x.size(-1)), lens


# ------------------------------------------------------------------------------------------------ DiT block
def self_attention(sd, pre, x, grid_sizes, freqs, num_heads):
    """1B:383-413."""
    b, s, C = x.shape
    d = C // num_heads
    q = rms_norm(linear(x, sd, pre + "q"), sd[pre + "norm_q.weight"]).view(b, s, num_heads, d)
    k = rms_norm(linear(x, sd, pre + "k"), sd[pre + "norm_k.weight"]).view(b, s, num_heads, d)
    v = linear(x, sd, pre + "v").view(b, s, num_heads, d)
    a = attention(rope_apply(q, grid_sizes, freqs), rope_apply(k, grid_sizes, freqs), v)
    return linear(a.flatten(2), sd, pre + "o")


def cross_attention(sd, pre, x, context, vocal_context, G, num_heads):
    """1B:534-605 — text + CLIP-image + audio attentions share q and are summed before o. Audio: tokens are grouped
    by view(b*G, -1, ...) (group = token_index // (L/G)); when vocal_context is 3-D every token sees all audio tokens."""
    b, L, C = x.shape
    d = C // num_heads
    ctx_img, ctx_txt = context[:, :257], context[:, 257:]
    q = rms_norm(linear(x, sd, pre + "q"), sd[pre + "norm_q.weight"]).view(b, -1, num_heads, d)
    k = rms_norm(linear(ctx_txt, sd, pre + "k"), sd[pre + "norm_k.weight"]).view(b, -1, num_heads, d)
    v = linear(ctx_txt, sd, pre + "v").view(b, -1, num_heads, d)
    k_img = rms_norm(linear(ctx_img, sd, pre + "k_img"), sd[pre + "norm_k_img.weight"]).view(b, -1, num_heads, d)
    v_img = linear(ctx_img, sd, pre + "v_img").view(b, -1, num_heads, d)
    img_x = attention(q, k_img, v_img)
    txt_x = attention(q, k, v)
    if vocal_context.dim() == 4:
        vq = q.view(b * G, -1, num_heads, d)
        vk = linear(vocal_context, sd, pre + "k_vocal").view(b * G, -1, num_heads, d)
        vv = linear(vocal_context, sd, pre + "v_vocal").view(b * G, -1, num_heads, d)
        voc_x = attention(vq, vk, vv).view(b, L, num_heads, d)
    else:
        vk = linear(vocal_context, sd, pre + "k_vocal").view(b, -1, num_heads, d)
        vv = linear(vocal_context, sd, pre + "v_vocal").view(b, -1, num_heads, d)
        voc_x = attention(q, vk, vv)
    return linear(txt_x.flatten(2) + img_x.flatten(2) + voc_x.flatten(2), sd, pre + "o")


def dit_block(sd, pre, x, e0, grid_sizes, freqs, context, vocal_context, G, num_heads):
    """1B:650-695."""
    e = (sd[pre + "modulation"] + e0).chunk(6, dim=1)
    temp = layer_norm(x) * (1 + e[1]) + e[0]
    x = x + self_attention(sd, pre + "self_attn.", temp, grid_sizes, freqs, num_heads) * e[2]
    xn = layer_norm(x, sd[pre + "norm3.weight"], sd[pre + "norm3.bias"])
    x = x + cross_attention(sd, pre + "cross_attn.", xn, context, vocal_context, G, num_heads)
    temp = layer_norm(x) * (1 + e[4]) + e[3]
    y = linear(F.gelu(linear(temp, sd, pre + "ffn.0"), approximate="tanh"), sd, pre + "ffn.2")
    return x + y * e[5]


# ------------------------------------------------------------------------------------------------ TeaCache
class TeaCache:
    """tc:19-74 + the bookkeeping in 1B:1021-1103 (cond_flag=True path used by the pipeline)."""

    def __init__(self, coefficients, num_steps, rel_l1_thresh=0.0, num_skip_start_steps=0):
        self.rescale = np.poly1d(coefficients)
        self.num_steps, self.thresh, self.skip_start = num_steps, rel_l1_thresh, num_skip_start_steps
        self.reset()

    def reset(self):
        self.cnt, self.acc, self.prev_inp, self.prev_residual = 0, 0, None, None

    def decide(self, e0):
        skip = self.cnt < self.skip_start
        if self.cnt == 0 or self.cnt == self.num_steps - 1 or skip:
            calc, self.acc = True, 0
        else:
            rel = ((e0 - self.prev_inp).abs().mean() / self.prev_inp.abs().mean()).item()
            self.acc += self.rescale(rel)
            if self.acc < self.thresh:
                calc = False
            else:
                calc, self.acc = True, 0
        self.prev_inp = e0
        self.cnt += 1
        if self.cnt == self.num_steps:
            prev = self.prev_residual
            self.reset()
            self.prev_residual = prev if not calc else None
        return calc


# ------------------------------------------------------------------------------------------------ full forward
def dit_forward(sd, cfg, x, t, context, seq_len, clip_fea, y, vocal_embeddings, video_sample_n_frames=81,
                is_clip_level_modeling=False, hooks=None, teacache=None):
    """1B:928-1159 (sp_world_size == 1). x [B,16,F,H,W], y [B,20,F,H,W], t [B], context list of [Li, text_dim],
    clip_fea [B,257,1280], vocal_embeddings [B,T,768]. cfg: dict(dim, num_heads, num_layers, freq_dim, text_len,
    patch_size, out_dim). hooks: optional dict collecting per-block outputs under 'block{i}' / 'vocal_context'."""
    dim, nh, nl = cfg["dim"], cfg["num_heads"], cfg["num_layers"]
    ps = tuple(cfg.get("patch_size", (1, 2, 2)))
    freqs = rope_freqs(dim // nh)
    xs = [torch.cat([u, v], dim=0) for u, v in zip(x, y)]
    xs = [F.conv3d(u.unsqueeze(0), sd["patch_embedding.weight"], sd["patch_embedding.bias"], stride=ps) for u in xs]
    grid_sizes = [tuple(u.shape[2:]) for u in xs]
    xs = [u.flatten(2).transpose(1, 2) for u in xs]
    assert max(u.size(1) for u in xs) <= seq_len
    h = torch.cat([torch.cat([u, u.new_zeros(1, seq_len - u.size(1), u.size(2))], dim=1) for u in xs])

    e = linear(F.silu(linear(sinusoidal_embedding_1d(cfg["freq_dim"], t).float(), sd, "time_embedding.0")), sd,
               "time_embedding.2")
    e0 = linear(F.silu(e), sd, "time_projection.1").unflatten(1, (6, dim))

    ctx = torch.stack([torch.cat([u, u.new_zeros(cfg["text_len"] - u.size(0), u.size(1))]) for u in context])
    ctx = linear(F.gelu(linear(ctx, sd, "text_embedding.0"), approximate="tanh"), sd, "text_embedding.2")
    c = F.layer_norm(clip_fea, (clip_fea.shape[-1],), sd["img_emb.proj.0.weight"], sd["img_emb.proj.0.bias"], 1e-5)
    c = linear(F.gelu(linear(c, sd, "img_emb.proj.1")), sd, "img_emb.proj.3")
    c = F.layer_norm(c, (dim,), sd["img_emb.proj.4.weight"], sd["img_emb.proj.4.bias"], 1e-5)
    ctx = torch.cat([c, ctx], dim=1)

    if cfg.get("variant") == "14B":                       # 14B:1008 — adapter on every sample, always 81 frames / 21 groups
        video_sample_n_frames = 81
        vc, _ = vocal_projector(sd, vocal_embeddings, 81, h, e0, e)
    elif vocal_embeddings.size(0) > 1:                    # 1B:1004-1007 — adapter once, replicated [0, vc, vc]
        vc, _ = vocal_projector(sd, vocal_embeddings[-1:], video_sample_n_frames, h[-1:], e0[-1:], e[-1:])
        vc = torch.cat([torch.zeros_like(vc), vc, vc])
    else:
        vc, _ = vocal_projector(sd, vocal_embeddings, video_sample_n_frames, h, e0, e)
    G = (video_sample_n_frames - 1) // 4 + 1
    if is_clip_level_modeling:
        vc = vc.flatten(1, 2)
    if hooks is not None:
        hooks["vocal_context"] = vc
        hooks["e0"] = e0

    def run_blocks(h):
        for i in range(nl):
            h = dit_block(sd, f"blocks.{i}.", h, e0, grid_sizes, freqs, ctx, vc, G, nh)
            if hooks is not None:
                hooks[f"block{i}"] = h
        return h

    if teacache is not None:
        if teacache.decide(e0):
            ori = h.clone()
            h = run_blocks(h)
            teacache.prev_residual = h - ori
        else:
            h = h + teacache.prev_residual
    else:
        h = run_blocks(h)

    em = (sd["head.modulation"] + e.unsqueeze(1)).chunk(2, dim=1)
    h = linear(layer_norm(h) * (1 + em[1]) + em[0], sd, "head.head")
    out = []
    for u, g in zip(h, grid_sizes):                       # 1B:1161-1184 unpatchify
        u = u[:math.prod(g)].view(*g, *ps, cfg.get("out_dim", 16))
        u = torch.einsum("fhwpqrc->cfphqwr", u)
        out.append(u.reshape(cfg.get("out_dim", 16), *[a * b for a, b in zip(g, ps)]))
    return torch.stack(out)
