"""Wan causal-3D VAE decode and encode on hand-written sm_100a kernels.

Drop-in for `AutoencoderKLWan` (wan/models/wan_vae.py:619-704): same constructor arguments,
`.config.{latent_channels, temporal_compression_ratio, spacial_compression_ratio}`, the reference's state-dict key
names (`model.decoder...`, `model.conv2...`, `model.encoder...`, `model.conv1...`),
`decode(z, return_dict=True) -> DecoderOutput(.sample)` with z [B,16,T,h,w] -> [B,3,1+4(T-1),8h,8w] fp32 in [-1,1], and
`encode(x, return_dict=True) -> AutoencoderKLOutput(.latent_dist)` with x [B,3,T,H,W] -> a diagonal Gaussian over
[B,16,1+(T-1)//4,H/8,W/8] whose `.mode()` is what the pipeline consumes (pipe.py:402-403).

Internals are B200-first rather than a transcription of Decoder3d.forward (:426-475): activations are channels-last
bf16, every 3x3x3 causal conv is an implicit GEMM on tcgen05 fed by TMA (ops.conv3d_cl), and the 33-entry feature cache
(:549-574 — `x[:, :, -2:].clone()` + `torch.cat` per conv per chunk) becomes one persistent ring buffer per conv whose
two leading frames are the cache: the producer of a conv's input writes straight behind them. Chunking follows the
reference exactly (one latent frame per chunk, chunk 0 skips the temporal upsamplers and yields a single frame).

`encode` (the conditioning clip) reuses the same kernels: chunks of 1, 4, 4, ... frames
(:523-538); the stride-2 Conv2d of Resample('downsample*') runs as a 2x2-tap conv over a space-to-depth copy, the
stride-(2,1,1) time_conv as the same implicit GEMM with a temporal stride over a one-frame ring buffer.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn as nn

from . import ops
from .synth import vae_decoder_layout, vae_decoder_param_shapes, vae_encoder_layout, vae_encoder_param_shapes

LATENT_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508, 0.4134, -0.0715, 0.5517, -0.3632,
               -0.1922, -0.9497, 0.2503, -0.2921]
LATENT_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743, 3.2687, 2.1526, 2.8652, 1.5579, 1.6382,
              1.1253, 2.8251, 1.9160]


class DecoderOutput:
    def __init__(self, sample):
        self.sample = sample


class DiagonalGaussianDistribution:
    """The slice of diffusers' class the reference uses on the encode result (wan_vae.py:655; pipe.py:402-403)."""

    def __init__(self, parameters):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std, self.var = torch.exp(0.5 * self.logvar), torch.exp(self.logvar)

    def mode(self):
        return self.mean

    def sample(self, generator=None):
        noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise


class AutoencoderKLOutput:
    def __init__(self, latent_dist):
        self.latent_dist = latent_dist

    def __getitem__(self, i):
        return (self.latent_dist,)[i]


def _build_param_tree(root: nn.Module, shapes: dict):
    for name, shape in shapes.items():
        mod = root
        *path, leaf = name.split(".")
        for part in path:
            if not hasattr(mod, part):
                mod.add_module(part, nn.Module())
            mod = getattr(mod, part)
        mod.register_parameter(leaf, nn.Parameter(torch.empty(*shape), requires_grad=False))


class _Conv:
    """One causal conv with its persistent input ring buffer [2 + Tmax, H, W, Cin] (two leading cache frames)."""

    use_halo = True        # A/B switch of tools/vae_bench.py; the per-tap kernel computes the same convolution
    generation = 0         # bumped whenever any ring buffer is (re)allocated: captured CUDA graphs hold their addresses

    def __init__(self, weight, bias, dev, lead=None, stride_t=1, pad=None):
        if weight.dim() == 4:
            weight = weight.unsqueeze(2)
        cout, cin, kt, kh, kw = weight.shape
        self.lead = kt - 1 if lead is None else lead           # cache frames kept in front of the chunk
        self.stride_t, self.pad = stride_t, pad
        self.cout, self.cin, self.k = cout, (cin + 31) // 32 * 32, (kt, kh, kw)
        cout_pad = (cout + 15) // 16 * 16
        w = torch.zeros(cout_pad, kt, kh, kw, self.cin, device=dev, dtype=torch.float32)
        w[:cout, ..., :cin] = weight.to(dev, torch.float32).permute(0, 2, 3, 4, 1)
        self.w = w.reshape(cout_pad, -1).to(torch.bfloat16).contiguous()
        # the convs that carry the decoder / encoder (3x3x3, 96 / 192 output channels) take the halo-staged kernel
        self.halo_head = self.use_halo and lead is None and stride_t == 1 and pad is None and \
            ops.conv3d_halo_supported(self.cin, cout, (kt, kh, kw), out_mode=2)            # Cout <= 16: the video head
        self.halo = (self.use_halo and lead is None and stride_t == 1 and pad is None and cout_pad == cout and
                     ops.conv3d_halo_supported(self.cin, cout, (kt, kh, kw)))
        self.w_halo = ops.pack_conv_weight_halo(w[:cout].to(torch.bfloat16)) if (self.halo or self.halo_head) else None
        self.bias = bias.to(dev, torch.float32).contiguous()
        self.buf = None

    def alloc(self, tmax, H, W, dev):
        lead = self.lead
        if self.buf is None or self.buf.shape != (lead + tmax, H, W, self.cin):
            self.buf = torch.zeros(lead + tmax, H, W, self.cin, device=dev, dtype=torch.bfloat16)
            _Conv.generation += 1
        else:
            self.buf[:lead].zero_()
        return self

    def slot(self, tc):
        """Where the producer writes this chunk's `tc` input frames."""
        lead = self.lead
        return self.buf[lead:lead + tc]

    def run(self, tc, out, res=None, keep_cache=True, **kw):
        lead = self.lead
        # (small frames leave the halo kernel's coarse work units — 2-4 output tiles x 96 / 192 channels — short of one wave)
        big = tc * ((self.buf.shape[1] + 15) // 16) * ((self.buf.shape[2] + 7) // 8) * max(1, self.cout // 192) >= 592
        if self.halo_head and big and kw.get("out_mode", 0) == 2 and res is None:
            ops.conv3d_halo_cl(self.buf[:lead + tc], self.w_halo, self.bias, cout=self.cout, out=out, kt=self.k[0], **kw)
        elif self.halo and big and kw.get("out_mode", 0) in (0, 1) and set(kw) <= {"out_mode"}:
            ops.conv3d_halo_cl(self.buf[:lead + tc], self.w_halo, self.bias, cout=self.cout, out=out, res=res, kt=self.k[0], **kw)
        else:
            ops.conv3d_cl(self.buf[:lead + tc], self.w, self.bias, cout=self.cout, k=self.k, out=out, res=res, pad=self.pad,
                          stride_t=self.stride_t, **kw)
        if keep_cache and lead:                      # new cache = last two frames of [old cache, x]  (wan_vae.py:208-220)
            for j in range(lead):
                self.buf[j].copy_(self.buf[tc + j])
        return out


class AutoencoderKLWan(nn.Module):
    def __init__(self, latent_channels=16, temporal_compression_ratio=4, spacial_compression_ratio=8):
        super().__init__()
        self.config = SimpleNamespace(latent_channels=latent_channels, temporal_compression_ratio=temporal_compression_ratio,
                                      spacial_compression_ratio=spacial_compression_ratio)
        self.dims, self.mods = vae_decoder_layout()
        _build_param_tree(self, vae_decoder_param_shapes(z_dim=latent_channels))
        self.enc_dims, self.enc_mods = vae_encoder_layout()
        _build_param_tree(self, vae_encoder_param_shapes(z_dim=latent_channels))
        self._has_encoder = False          # set by load_state_dict / init when encoder weights are actually present
        self.mean = torch.tensor(LATENT_MEAN, dtype=torch.float32)
        self.std = torch.tensor(LATENT_STD, dtype=torch.float32)
        self.scale = [self.mean, 1.0 / self.std]
        self._prep = self._prep_enc = None
        self._pp_group, self._pp_world, self._pp_rank = None, 1, 0
        # Opt-in (long-lived serving processes): steady-state decode chunks replayed from one captured graph. It saves
        # ~1.5 % of a decode (the decode is GPU-bound: 362 ms of kernels in 366 ms) but instantiating the graph costs
        # 0.3-0.8 s the first time a shape is seen twice, more than a one-off caller ever gets back.
        self.use_cuda_graph = False
        self._dec_graph = self._dec_seen = None

    @property
    def dtype(self):
        return torch.float32          # the boundary dtype of the reference VAE (inference.py:471-474)

    @property
    def device(self):
        return self.model.conv2.weight.device

    @staticmethod
    def _is_enc_key(k):
        return k.startswith("model.encoder.") or k.startswith("model.conv1.")

    def load_state_dict(self, state_dict, strict=True, assign=False):
        """Reference key names. A decode-only state dict (no `model.encoder.*` / `model.conv1.*`) is accepted even
        with strict=True; `encode` then raises instead of running on uninitialised weights."""
        self._prep = self._prep_enc = None
        own = set(self.state_dict().keys())
        kept = {k: v for k, v in state_dict.items() if k in own}
        extra = [k for k in state_dict if k not in own]
        if strict and extra:
            raise RuntimeError(f"unexpected keys: {extra[:5]}")
        has_enc = any(self._is_enc_key(k) for k in kept)
        if not has_enc:
            for k, v in self.state_dict().items():
                if self._is_enc_key(k):
                    kept[k] = v
        res = super().load_state_dict(kept, strict=strict, assign=assign)
        self._has_encoder = has_enc
        return res

    def _apply(self, fn, *a, **k):
        self._prep = self._prep_enc = None
        return super()._apply(fn, *a, **k)

    @classmethod
    def from_pretrained(cls, pretrained_model_path, additional_kwargs={}):
        """wan_vae.py:683-704: raw `.pth` / `.safetensors` state dict whose keys get the `model.` prefix."""
        import inspect
        keys = set(inspect.signature(cls.__init__).parameters) - {"self"}
        model = cls(**{k: v for k, v in additional_kwargs.items() if k in keys})
        if pretrained_model_path.endswith(".safetensors"):
            from safetensors.torch import load_file
            state_dict = load_file(pretrained_model_path)
        else:
            state_dict = torch.load(pretrained_model_path, map_location="cpu")
        m, u = model.load_state_dict({"model." + k: v for k, v in state_dict.items()}, strict=False)
        print(f"### missing keys: {len(m)}; \n### unexpected keys: {len(u)};")
        return model

    # ------------------------------------------------------------------ encode
    def _prepare_encoder(self):
        if self._prep_enc is not None:
            return self._prep_enc
        if not self._has_encoder:
            raise RuntimeError("AutoencoderKLWan (B200): no encoder weights were loaded (decode-only state dict); "
                               "load a state dict with model.encoder.* / model.conv1.* before calling encode")
        sd = {k: v for k, v in self.state_dict().items()}
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("AutoencoderKLWan (B200): parameters must live on a CUDA device; there is no CPU fallback")
        P = "model.encoder."
        conv = lambda n, **kw: _Conv(sd[n + ".weight"], sd[n + ".bias"], dev, **kw)  # noqa: E731
        f32 = lambda n: sd[n].to(dev, torch.float32).reshape(-1).contiguous()  # noqa: E731
        bf = lambda t: t.to(dev, torch.bfloat16).contiguous()  # noqa: E731

        def res(pre, cin, cout):
            d = dict(g0=f32(pre + "residual.0.gamma"), c0=conv(pre + "residual.2"), g1=f32(pre + "residual.3.gamma"),
                     c1=conv(pre + "residual.6"), cin=cin, cout=cout)
            if cin != cout:
                d["w_sc"] = bf(sd[pre + "shortcut.weight"].reshape(cout, cin))
                d["b_sc"] = bf(sd[pre + "shortcut.bias"])
            return d

        def down_conv(n):
            """3x3 stride-2 Conv2d behind ZeroPad2d(0,1,0,1) (:96-98) as a 2x2-tap conv over the space-to-depth input:
            input row 2y + ky = 2(y + ky//2) + ky%2, so tap (ty, tx) x sub-position (dy, dx) holds W[2ty+dy, 2tx+dx]."""
            w = sd[n + ".weight"].to(torch.float32)
            cout, c = w.shape[:2]
            w2 = torch.zeros(cout, 4 * c, 1, 2, 2)
            for ky in range(3):
                for kx in range(3):
                    d = (ky % 2) * 2 + (kx % 2)
                    w2[:, d * c:(d + 1) * c, 0, ky // 2, kx // 2] = w[:, :, ky, kx]
            return _Conv(w2, sd[n + ".bias"], dev, pad=(0, 0))

        downs = []
        for i, m in enumerate(self.enc_mods):
            q = f"{P}downsamples.{i}."
            if m[0] == "res":
                downs.append(("res", res(q, m[1], m[2])))
            else:
                d = dict(c=m[1], conv=down_conv(q + "resample.1"))
                if m[0] == "down3d":                         # (3,1,1) stride (2,1,1), one cached frame in front (:153-160)
                    d["tconv"] = conv(q + "time_conv", lead=1, stride_t=2)
                downs.append((m[0], d))
        c = self.enc_dims[-1]
        wqkv, bqkv = sd[P + "middle.1.to_qkv.weight"].reshape(3 * c, c), sd[P + "middle.1.to_qkv.bias"]
        attn = dict(g=f32(P + "middle.1.norm.gamma"), w_qk=bf(wqkv[:2 * c]), b_qk=bf(bqkv[:2 * c]), w_v=bf(wqkv[2 * c:]),
                    b_v=bf(bqkv[2 * c:]), w_o=bf(sd[P + "middle.1.proj.weight"].reshape(c, c)), b_o=bf(sd[P + "middle.1.proj.bias"]))
        z2 = 2 * self.config.latent_channels
        self._prep_enc = dict(conv1=conv(P + "conv1"), downs=downs, mid0=res(P + "middle.0.", c, c), attn=attn,
                              mid2=res(P + "middle.2.", c, c), head_g=f32(P + "head.0.gamma"), head=conv(P + "head.2"),
                              wc=sd["model.conv1.weight"].to(dev, torch.float32).reshape(z2, z2).contiguous(),
                              bc=sd["model.conv1.bias"].to(dev, torch.float32).contiguous(),
                              mean=self.mean.to(dev), std=self.std.to(dev))
        return self._prep_enc

    def _alloc_encoder(self, H, W, dev):
        """Ring buffers sized for a steady-state chunk: 4 frames until the first temporal downsample, then 2, then 1."""
        p = self._prepare_encoder()
        tc = 4
        p["conv1"].alloc(tc, H, W, dev)
        for kind, d in p["downs"]:
            if kind == "res":
                d["c0"].alloc(tc, H, W, dev), d["c1"].alloc(tc, H, W, dev)
            else:
                H, W = H // 2, W // 2
                d["conv"].alloc(tc, H, W, dev)
                if kind == "down3d":
                    d["tconv"].alloc(tc, H, W, dev)
                    tc //= 2
        for d in (p["mid0"], p["mid2"]):
            d["c0"].alloc(tc, H, W, dev), d["c1"].alloc(tc, H, W, dev)
        p["head"].alloc(tc, H, W, dev)

    def _encode_chunk(self, a, chunk, h_out):
        """Encoder3d.forward (:324-369) on one chunk a [tc, H, W, 32] bf16 (video frames, channels zero-padded);
        writes the head's fp32 output [tc', H/8, W/8, 2*z] into h_out."""
        p = self._prepare_encoder()
        dev, bf = a.device, torch.bfloat16
        tc, H, W, _ = a.shape
        p["conv1"].slot(tc).copy_(a)
        a = p["conv1"].run(tc, torch.empty(tc, H, W, self.enc_dims[0], device=dev, dtype=bf))
        for kind, d in p["downs"]:
            if kind == "res":
                a = self._res_block(d, a, tc)
                continue
            C = d["c"]
            if H % 2 or W % 2:
                raise ValueError(f"AutoencoderKLWan.encode: feature map {H}x{W} is not even; H and W must be multiples of 8")
            H, W = H // 2, W // 2
            ops.vae_space_to_depth(a, d["conv"].slot(tc))
            if kind == "down2d":
                a = d["conv"].run(tc, torch.empty(tc, H, W, C, device=dev, dtype=bf))
            elif chunk == 0:                                 # first chunk: the frame only seeds the time_conv cache (:148-151)
                a = d["conv"].run(tc, torch.empty(tc, H, W, C, device=dev, dtype=bf))
                d["tconv"].buf[0].copy_(a[tc - 1])
            else:
                d["conv"].run(tc, d["tconv"].slot(tc))
                a = d["tconv"].run(tc, torch.empty(tc // 2, H, W, C, device=dev, dtype=bf), keep_cache=False)
                d["tconv"].buf[0].copy_(d["tconv"].buf[tc])  # cache = last frame of this chunk (:154)
                tc //= 2
        a = self._res_block(p["mid0"], a, tc)
        a = self._attention(p["attn"], a)
        a = self._res_block(p["mid2"], a, tc)
        ops.vae_rmsnorm_silu(a, p["head_g"], p["head"].slot(tc))
        p["head"].run(tc, h_out, out_mode=3)
        return tc

    @torch.no_grad()
    def _encode_one(self, x, out):
        """x [3, T, H, W] fp32 -> out [2z, 1 + (T-1)//4, H/8, W/8] fp32: AutoencoderKLWan_.encode (:519-547)."""
        p = self._prepare_encoder()
        dev = x.device
        _, T, H, W = x.shape
        if H % 8 or W % 8:
            raise ValueError(f"AutoencoderKLWan.encode: H and W must be multiples of 8, got {H}x{W}")
        self._alloc_encoder(H, W, dev)
        n_chunks = 1 + (T - 1) // 4
        z2 = out.shape[0]
        h_all = torch.empty(n_chunks, H // 8, W // 8, z2, device=dev, dtype=torch.float32)
        for i in range(n_chunks):
            frames = x[:, :1] if i == 0 else x[:, 1 + 4 * (i - 1):1 + 4 * i]
            a = ops.vae_video_in(frames.contiguous(), 32)
            tcl = self._encode_chunk(a, i, h_all[i:i + 1])
            assert tcl == 1
        ops.vae_latent_out(h_all, p["wc"], p["bc"], p["mean"], p["std"], out)

    @torch.no_grad()
    def encode(self, x, return_dict=True):
        """wan_vae.py:649-664. x [B, 3, T, H, W] in [-1, 1]; frames beyond 1 + 4k are dropped like the reference does."""
        dev = self.device
        x = x.to(dev, torch.float32)
        B, _, T, H, W = x.shape
        z2 = 2 * self.config.latent_channels
        h = torch.empty(B, z2, 1 + (T - 1) // 4, H // 8, W // 8, device=dev, dtype=torch.float32)
        for b in range(B):
            self._encode_one(x[b], h[b])
        posterior = DiagonalGaussianDistribution(h)
        if not return_dict:
            return (posterior,)
        return AutoencoderKLOutput(latent_dist=posterior)

    # ------------------------------------------------------------------ one-time operand preparation
    def _prepare(self):
        if self._prep is not None:
            return self._prep
        sd = {k: v for k, v in self.state_dict().items()}
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("AutoencoderKLWan (B200): parameters must live on a CUDA device; there is no CPU fallback")
        P = "model.decoder."
        conv = lambda n: _Conv(sd[n + ".weight"], sd[n + ".bias"], dev)  # noqa: E731
        f32 = lambda n: sd[n].to(dev, torch.float32).reshape(-1).contiguous()  # noqa: E731
        bf = lambda t: t.to(dev, torch.bfloat16).contiguous()  # noqa: E731

        def res(pre, cin, cout):
            d = dict(g0=f32(pre + "residual.0.gamma"), c0=conv(pre + "residual.2"), g1=f32(pre + "residual.3.gamma"),
                     c1=conv(pre + "residual.6"), cin=cin, cout=cout)
            if cin != cout:
                d["w_sc"] = bf(sd[pre + "shortcut.weight"].reshape(cout, cin))
                d["b_sc"] = bf(sd[pre + "shortcut.bias"])
            return d

        c = self.dims[0]
        wqkv, bqkv = sd[P + "middle.1.to_qkv.weight"].reshape(3 * c, c), sd[P + "middle.1.to_qkv.bias"]
        attn = dict(g=f32(P + "middle.1.norm.gamma"), w_qk=bf(wqkv[:2 * c]), b_qk=bf(bqkv[:2 * c]), w_v=bf(wqkv[2 * c:]),
                    b_v=bf(bqkv[2 * c:]), w_o=bf(sd[P + "middle.1.proj.weight"].reshape(c, c)), b_o=bf(sd[P + "middle.1.proj.bias"]))
        ups = []
        for i, m in enumerate(self.mods):
            q = f"{P}upsamples.{i}."
            if m[0] == "res":
                ups.append(("res", res(q, m[1], m[2])))
            else:
                d = dict(c=m[1], conv=conv(q + "resample.1"))
                if m[0] == "up3d":
                    d["tconv"] = conv(q + "time_conv")
                ups.append((m[0], d))
        self._prep = dict(conv1=conv(P + "conv1"), mid0=res(P + "middle.0.", c, c), attn=attn, mid2=res(P + "middle.2.", c, c),
                          ups=ups, head_g=f32(P + "head.0.gamma"), head=conv(P + "head.2"),
                          wc=sd["model.conv2.weight"].to(dev, torch.float32).reshape(16, 16).contiguous(),
                          bc=sd["model.conv2.bias"].to(dev, torch.float32).contiguous(),
                          mean=self.mean.to(dev), std=self.std.to(dev))
        return self._prep

    # ------------------------------------------------------------------ decode
    def _res_block(self, d, a, tc):
        """ResidualBlock (wan_vae.py:205-223): x + conv(silu(norm(conv(silu(norm(x)))))) with an optional 1x1x1 shortcut."""
        T, H, W, _ = a.shape
        dev = a.device
        if "w_sc" in d:
            h = ops.gemm(a.view(-1, d["cin"]), d["w_sc"], d["b_sc"]).view(T, H, W, d["cout"])
        else:
            h = a
        ops.vae_rmsnorm_silu(a, d["g0"], d["c0"].slot(tc))
        b = d["c0"].run(tc, torch.empty(T, H, W, d["cout"], device=dev, dtype=torch.bfloat16))
        ops.vae_rmsnorm_silu(b, d["g1"], d["c1"].slot(tc))
        return d["c1"].run(tc, torch.empty(T, H, W, d["cout"], device=dev, dtype=torch.bfloat16), res=h)

    def _attention(self, d, a):
        """AttentionBlock (wan_vae.py:243-265): per frame, one head of width C over H*W positions."""
        T, H, W, C = a.shape
        P = H * W
        ldp = (P + 7) // 8 * 8
        out = torch.empty_like(a)
        for t in range(T):
            x = a[t].view(P, C)
            xn = ops.vae_rmsnorm_silu(x, d["g"], torch.empty_like(x), silu=False)
            qk = ops.gemm(xn, d["w_qk"], d["b_qk"])                                       # [P, 2C]
            vt = torch.zeros(C, ldp, device=a.device, dtype=torch.bfloat16)
            ops.gemm(d["w_v"], xn, out=vt[:, :P])                                         # V^T [C, P] (bias folded below)
            s = torch.empty(P, ldp, device=a.device, dtype=torch.float32)
            ops.gemm(qk[:, :C], qk[:, C:], out=s[:, :P], round_y=False)                   # q k^T in fp32
            p = torch.zeros(P, ldp, device=a.device, dtype=torch.bfloat16)
            ops.softmax_rows_into(s[:, :P], p[:, :P], C ** -0.5)
            o = ops.gemm(p, vt, d["b_v"])                                                 # softmax rows sum to 1: P(V + b) = PV + b
            ops.gemm(o, d["w_o"], d["b_o"], res=x, out=out[t].view(P, C))
        return out

    # ------------------------------------------------------------------ decoder as a list of units
    # One latent frame (chunk) flows through 20 units: conv1, middle (res, attention, res), the 15 upsample-stage modules,
    # head. A unit owns its causal ring buffers, so a contiguous range of units can run on its own GPU: pipeline
    # parallelism over the decoder depth keeps every op and every cache exactly as on one GPU (the caches chain through
    # time, so the decode cannot be split along T; a W split would need a halo exchange in each of the 33 convs).
    def _units(self):
        p = self._prepare()
        units = [("conv1", None), ("res", p["mid0"]), ("attn", p["attn"]), ("res", p["mid2"])]
        units += list(p["ups"])
        units.append(("head", None))
        return units

    def _unit_shapes(self, h, w, chunk):
        """(frames, H, W, C) entering every unit and leaving the last one, for chunk index `chunk`."""
        shapes, tc, H, W, C = [], 1, h, w, 32
        for kind, d in self._units():
            shapes.append((tc, H, W, C))
            if kind == "conv1":
                C = self.dims[0]
            elif kind == "res":
                C = d["cout"]
            elif kind in ("up3d", "up2d"):
                if kind == "up3d" and chunk > 0:
                    tc *= 2
                H, W, C = 2 * H, 2 * W, d["c"] // 2
            elif kind == "head":
                C = 3
        shapes.append((tc, H, W, C))
        return shapes

    def _unit_costs(self, h, w):
        """Relative time of every unit for a steady-state chunk: conv FLOPs divided by a per-channel-width efficiency
        taken from tools/conv_bench.py on the halo-staged kernel (profiles/r02_conv_bench.log)."""
        costs = []
        eff = lambda c: 1.05 if c <= 96 else (1.35 if c <= 192 else 1.2)  # noqa: E731  PFLOP/s
        for (kind, d), (tc, H, W, C) in zip(self._units(), self._unit_shapes(h, w, 1)):
            pos = tc * H * W
            if kind == "res":
                f = 2 * pos * 27 * (d["cin"] * d["cout"] + d["cout"] * d["cout"]) + (2 * pos * d["cin"] * d["cout"] if "w_sc" in d else 0)
                costs.append(f / eff(d["cout"]) + 8 * pos * d["cout"] * 40)
            elif kind == "attn":
                costs.append(4 * pos * pos * C + 8 * pos * C * C)
            elif kind in ("up3d", "up2d"):
                t2 = tc * 2 if kind == "up3d" else tc
                f = 2 * t2 * 4 * H * W * 9 * C * (C // 2) + (2 * pos * 3 * C * 2 * C if kind == "up3d" else 0)
                costs.append(f / eff(C // 2) + 8 * t2 * 4 * H * W * C * 40)
            elif kind == "head":
                costs.append(2 * pos * 27 * C * 16 / 0.8 + 8 * pos * C * 40)
            else:
                costs.append(2 * pos * 27 * 32 * self.dims[0])
        return costs

    @staticmethod
    def partition_units(costs, stages):
        """Contiguous partition of `costs` into `stages` ranges minimising the largest range sum (binary search + greedy).
        Returns the list of (lo, hi) unit ranges, one per pipeline rank; trailing ranks may be empty."""
        n = len(costs)
        lo_c, hi_c = max(costs), sum(costs)
        for _ in range(60):
            mid = (lo_c + hi_c) / 2
            parts, acc = 1, 0.0
            for c in costs:
                if acc + c > mid:
                    parts, acc = parts + 1, c
                else:
                    acc += c
            if parts <= stages:
                hi_c = mid
            else:
                lo_c = mid
        ranges, start, acc = [], 0, 0.0
        for i, c in enumerate(costs):
            if acc + c > hi_c * (1 + 1e-9) and i > start:
                ranges.append((start, i))
                start, acc = i, c
            else:
                acc += c
        ranges.append((start, n))
        while len(ranges) < stages:
            ranges.append((n, n))
        return ranges

    def _alloc_units(self, lo, hi, h, w, dev):
        p = self._prepare()
        units = self._units()
        shapes = self._unit_shapes(h, w, 1)
        for idx in range(lo, hi):
            kind, d = units[idx]
            tc, H, W, _ = shapes[idx]
            if kind == "conv1":
                p["conv1"].alloc(1, H, W, dev)
            elif kind == "res":
                d["c0"].alloc(tc, H, W, dev), d["c1"].alloc(tc, H, W, dev)
            elif kind in ("up3d", "up2d"):
                if kind == "up3d":
                    d["tconv"].alloc(tc, H, W, dev)
                d["conv"].alloc(shapes[idx + 1][0], 2 * H, 2 * W, dev)
            elif kind == "head":
                p["head"].alloc(tc, H, W, dev)

    def _run_unit(self, idx, a, chunk, video, t_out):
        """a: [tc, H, W, C] bf16 entering unit idx for chunk `chunk` -> activation leaving it (None after the head)."""
        p = self._prepare()
        kind, d = self._units()[idx]
        dev, bf = a.device, torch.bfloat16
        tc, Hc, Wc, C = a.shape
        if kind == "conv1":
            p["conv1"].slot(1).copy_(a)
            return p["conv1"].run(1, torch.empty(1, Hc, Wc, self.dims[0], device=dev, dtype=bf))
        if kind == "res":
            return self._res_block(d, a, tc)
        if kind == "attn":
            return self._attention(d, a)
        if kind in ("up3d", "up2d"):
            if kind == "up3d" and chunk > 0:                  # chunk 0: 'Rep' — no temporal upsampling (wan_vae.py:108-112)
                d["tconv"].slot(tc).copy_(a)
                a = d["tconv"].run(tc, torch.empty(2 * tc, Hc, Wc, C, device=dev, dtype=bf), out_mode=1)
                tc *= 2
            ops.vae_upsample2x(a, d["conv"].slot(tc))
            return d["conv"].run(tc, torch.empty(tc, 2 * Hc, 2 * Wc, C // 2, device=dev, dtype=bf))
        ops.vae_rmsnorm_silu(a, p["head_g"], p["head"].slot(tc))     # head
        p["head"].run(tc, video, out_mode=2, out_T_total=video.shape[1], out_t0=t_out)
        return None

    @torch.no_grad()
    def _decode_one(self, z, video):
        """z [16, T, h, w] fp32 (one sample) -> video [3, 1 + 4 (T-1), 8h, 8w] fp32, AutoencoderKLWan_.decode :549-574.
        With `enable_multi_gpus_decode` each rank runs a contiguous range of units for every chunk and hands the
        activation to the next rank (NCCL send/recv); the last rank owns the frames and broadcasts them."""
        import torch.distributed as dist
        p = self._prepare()
        dev = z.device
        _, T, h, w = z.shape
        n_units = len(self._units())
        world, rank, group = self._pp_world, self._pp_rank, self._pp_group
        ranges = self.partition_units(self._unit_costs(h, w), world) if world > 1 else [(0, n_units)]
        lo, hi = ranges[rank]
        self._alloc_units(lo, hi, h, w, dev)
        peer = (lambda r: dist.get_global_rank(group, r) if group is not None else r)
        last = max(r for r in range(world) if ranges[r][1] > ranges[r][0])
        x_all = ops.vae_latent_in(z.contiguous(), p["wc"], p["bc"], p["mean"], p["std"], 32) if lo == 0 and hi > 0 else None
        t_out = 0
        # Every chunk after the first has the same shapes and touches the same ring buffers: chunk 1 runs eagerly (it is also
        # the first use of the temporal-upsampling kernels), chunk 2 is captured as ONE CUDA graph and replayed for the rest
        # — ~80 kernel launches and ~60 cache copies per chunk leave the host, which was the decode's bottleneck once the
        # convolutions got faster (wall 478 ms for 362 ms of kernels). The graph reads its latent frame from / writes its
        # 4 video frames to static buffers.
        # The graph is kept across decode() calls for as long as the ring buffers it addresses stay where they are
        # (instantiating ~140 nodes costs 0.3-0.5 s, more than the decode).
        # A shape is captured on its SECOND decode only: a one-off decode would pay more for the capture than it saves.
        gkey = (h, w, str(dev), _Conv.generation, id(p))
        use_graph = world == 1 and self.use_cuda_graph and T >= 4 and self._dec_seen == gkey
        self._dec_seen = gkey
        if use_graph and (self._dec_graph is None or self._dec_graph[0] != gkey):
            self._dec_graph = None
        for i in range(T):
            shapes = self._unit_shapes(h, w, i)
            if use_graph and i >= 2:
                nf = shapes[n_units - 1][0]
                if self._dec_graph is None:
                    gin = x_all[i:i + 1].clone()
                    gvid = torch.empty(3, nf, video.shape[2], video.shape[3], device=dev, dtype=torch.float32)
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        a = gin
                        for idx in range(n_units):
                            a = self._run_unit(idx, a, i, gvid, 0)
                    self._dec_graph = (gkey, graph, gin, gvid)
                else:
                    self._dec_graph[2].copy_(x_all[i:i + 1])
                _, graph, gin, gvid = self._dec_graph
                graph.replay()
                video[:, t_out:t_out + nf].copy_(gvid)
            elif hi > lo:
                if lo == 0:
                    a = x_all[i:i + 1]
                else:
                    a = torch.empty(shapes[lo], device=dev, dtype=torch.bfloat16)
                    dist.recv(a, peer(rank - 1), group=group)
                for idx in range(lo, hi):
                    a = self._run_unit(idx, a, i, video, t_out)
                if hi < n_units:
                    dist.send(a.contiguous(), peer(rank + 1), group=group)
            t_out += shapes[n_units - 1][0]
        assert t_out == video.shape[1], (t_out, video.shape)
        if world > 1:
            dist.broadcast(video, peer(last), group=group)

    def enable_multi_gpus_decode(self, group=None):
        """Pipeline-parallel decode over the ranks of `group` (default: the whole torch.distributed world). The
        reference has no multi-GPU VAE (every rank decodes the whole clip, inference.py:574-577)."""
        import torch.distributed as dist
        self._pp_group, self._pp_world, self._pp_rank = group, dist.get_world_size(group), dist.get_rank(group)

    @torch.no_grad()
    def decode(self, z, return_dict=True):
        dev = self.device
        z = z.to(dev, torch.float32)
        B, _, T, h, w = z.shape
        video = torch.empty(B, 3, 1 + 4 * (T - 1), 8 * h, 8 * w, device=dev, dtype=torch.float32)
        for b in range(B):
            self._decode_one(z[b], video[b])
        if not return_dict:
            return (video,)
        return DecoderOutput(sample=video)
