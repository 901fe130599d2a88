// norm_kernels.cu — HBM-bound row kernels of the DiT block: LayerNorm (+affine) (+AdaLN modulate) (+gated self
// residual), RMSNorm (+3-D RoPE), and the modulation-table broadcast add.
//
// One warp owns one row: each lane keeps its 16-byte chunks of the row in registers (C/256 chunks, 6 for C = 1536),
// statistics are reduced with shuffles, so a row is read once and written once (algorithmic bytes = 2 x row bytes),
// there is no shared memory and no block barrier, and every load/store is a coalesced 128-bit access.
//
// Rounding points follow the reference's bf16 autocast flow:
//   WanLayerNorm   (1B.py:345-355): stats in fp32, result .type_as(x)
//   modulation     (1B.py:675-676): norm(x) * (1 + e1) + e0, every op rounded to bf16 when x is bf16
//   WanRMSNorm     (1B.py:326-342): (x * rsqrt(mean(x^2) + eps)).type_as(x) * weight
//   rope_apply     (1B.py:296-323): adjacent pairs rotated by exp(i * pos * theta_j), 22/21/21 pairs for (f, h, w)
#include "../../include/stableavatar_b200.h"
#include "sa_host.h"
#include "sa_ptx.cuh"

namespace sa {
namespace norm {

constexpr int WARPS_PER_BLOCK = 8;
constexpr int MAX_CHUNKS = 20;  // C <= 20 * 256 = 5120 per warp-row (the 14B width)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void load_chunk(const void* base, int dtype, long long idx, float (&v)[8]) {
  if (dtype == SA_BF16) {
    uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  } else {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
    float4 a = p[0], b = p[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
}
__device__ __forceinline__ void store_chunk(void* base, int dtype, long long idx, const float (&v)[8]) {
  if (dtype == SA_BF16) {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]);
    u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]);
    u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + idx) = u;
  } else {
    float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx);
    p[0] = make_float4(v[0], v[1], v[2], v[3]);
    p[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm
struct LnParams {
  const void* x; void* out;
  const void* weight; const void* bias;
  const __nv_bfloat16* shift; const __nv_bfloat16* scale; const __nv_bfloat16* gate;
  const void* res;
  long long ldx, ldo, ldr, mod_bs;
  int rows, C, rows_per_batch, x_dtype, out_dtype, w_dtype, round_bf16;
  float eps;
};

// The row is kept in registers in its storage format (packed bf16: 4 words per 8 elements) so that 6 chunks cost 24
// registers instead of 48 and four CTAs stay resident per SM; dtypes are compile-time so no chunk array is ever
// indexed through a run-time branch (which ptxas turns into local-memory traffic).
template <bool XBF>
struct RowChunk;
template <>
struct RowChunk<true> {
  uint4 u;
  __device__ __forceinline__ void load(const void* base, long long idx) {
    u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx);
  }
  __device__ __forceinline__ void get(float (&v)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
};
template <>
struct RowChunk<false> {
  float4 a, b;
  __device__ __forceinline__ void load(const void* base, long long idx) {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
    a = p[0];
    b = p[1];
  }
  __device__ __forceinline__ void get(float (&v)[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};
__device__ __forceinline__ void load8_bf16(const __nv_bfloat16* p, float (&v)[8]) {
  RowChunk<true> c;
  c.load(p, 0);
  c.get(v);
}

template <int NCH, bool XBF, bool OBF, bool WBF>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, NCH <= 6 ? (XBF ? 3 : 2) : (NCH <= 8 ? 2 : 1)) layernorm_kernel(const LnParams p) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const int nchunks = p.C >> 3;
  RowChunk<XBF> v[NCH];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int ch = lane + c * 32;
    if (ch < nchunks) v[c].load(p.x, (long long)row * p.ldx + ch * 8);
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (lane + c * 32 < nchunks) {
      float t[8];
      v[c].get(t);
#pragma unroll
      for (int i = 0; i < 8; ++i) s += t[i];
    }
  }
  const float mean = warp_sum(s) / p.C;
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (lane + c * 32 < nchunks) {
      float t[8];
      v[c].get(t);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = t[i] - mean;
        ss += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / p.C + p.eps);
  const long long mrow = (long long)(row / p.rows_per_batch) * p.mod_bs;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int ch = lane + c * 32;
    if (ch >= nchunks) continue;
    const int col = ch * 8;
    float y[8];
    v[c].get(y);
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = (y[i] - mean) * rstd;
    if (p.weight) {
      float w[8], b[8];
      RowChunk<WBF> cw, cb;
      cw.load(p.weight, col);
      cb.load(p.bias, col);
      cw.get(w);
      cb.get(b);
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = fmaf(y[i], w[i], b[i]);
    }
    if (p.round_bf16) {
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = bf16_round(y[i]);
    }
    if (p.scale) {
      float sc[8], sh[8];
      load8_bf16(p.scale + mrow + col, sc);
      load8_bf16(p.shift + mrow + col, sh);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float one_plus = bf16_round(1.0f + sc[i]);  // (1 + e1) is a bf16 op in both streams
        float t = y[i] * one_plus;
        if (p.round_bf16) t = bf16_round(t);
        t += sh[i];
        if (p.round_bf16) t = bf16_round(t);
        y[i] = t;
      }
    }
    if (p.gate) {  // adapter "pseudo self-attention" (vp1B.py:345-347): out = res + y * e2
      float g[8], r[8];
      load8_bf16(p.gate + mrow + col, g);
      RowChunk<OBF> cr;
      cr.load(p.res, (long long)row * p.ldr + col);
      cr.get(r);
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = r[i] + y[i] * g[i];
    }
    store_chunk(p.out, OBF ? SA_BF16 : SA_F32, (long long)row * p.ldo + col, y);
  }
}

// The two LayerNorm shapes of a DiT block in bf16 mode, specialised at compile time: MODE 1 = norm1 / norm2 / head with
// AdaLN modulation, norm(x) * (1 + scale) + shift (1B.py:675, 687, 721), MODE 2 = norm3 with bf16 affine parameters
// (1B.py:638-640, 682). The generic kernel above emulates every bf16 tensor op of the modulation chain in scalar fp32
// (cvt + shift per rounding, ~25 issue slots per element: 73-83 % issue-active at 3.1 TB/s). Here the row is unpacked to
// fp32 registers once, statistics and the normalisation run in fp32, and the chain runs in PACKED bf16 arithmetic —
// cvt.rn.bf16x2.f32 for norm(x).type_as(x), then HADD2 / HMUL2 / HADD2 on bf16x2 with the packed scale / shift words as
// loaded (each a single IEEE rounding to bf16, i.e. exactly the reference's bf16 tensor ops) — ~8 issue slots per element.
template <int NCH, int MODE>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, NCH <= 6 ? 3 : 2) layernorm_bf16_kernel(const LnParams p) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const int nchunks = p.C >> 3;
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(p.x) + (long long)row * p.ldx;
  uint4 raw[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int ch = lane + c * 32;
    if (ch < nchunks) raw[c] = *reinterpret_cast<const uint4*>(x + ch * 8);
  }
  float v[NCH][8];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const uint32_t w4[4] = {raw[c].x, raw[c].y, raw[c].z, raw[c].w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[c][2 * i] = __uint_as_float(w4[i] << 16);
      v[c][2 * i + 1] = __uint_as_float(w4[i] & 0xffff0000u);
    }
    if (lane + c * 32 < nchunks) {
#pragma unroll
      for (int i = 0; i < 8; ++i) s += v[c][i];
    }
  }
  const float mean = warp_sum(s) / p.C;
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (lane + c * 32 < nchunks) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = v[c][i] - mean;
        ss = fmaf(d, d, ss);
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) / p.C + p.eps);
  const long long mrow = MODE == 1 ? (long long)(row / p.rows_per_batch) * p.mod_bs : 0;
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)row * p.ldo;
  const __nv_bfloat162 one2 = __floats2bfloat162_rn(1.0f, 1.0f);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int ch = lane + c * 32;
    if (ch >= nchunks) continue;
    const int col = ch * 8;
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
    if constexpr (MODE == 1) {
      const uint4 sc4 = *reinterpret_cast<const uint4*>(p.scale + mrow + col);
      const uint4 sh4 = *reinterpret_cast<const uint4*>(p.shift + mrow + col);
      const __nv_bfloat162* sc = reinterpret_cast<const __nv_bfloat162*>(&sc4);
      const __nv_bfloat162* sh = reinterpret_cast<const __nv_bfloat162*>(&sh4);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 y = __floats2bfloat162_rn((v[c][2 * i] - mean) * rstd, (v[c][2 * i + 1] - mean) * rstd);
        const __nv_bfloat162 t = __hadd2(__hmul2(y, __hadd2(one2, sc[i])), sh[i]);
        ow[i] = *reinterpret_cast<const uint32_t*>(&t);
      }
    } else {
      const uint4 w4 = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.weight) + col);
      const uint4 b4 = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.bias) + col);
      const uint32_t ww[4] = {w4.x, w4.y, w4.z, w4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float y0 = fmaf((v[c][2 * i] - mean) * rstd, __uint_as_float(ww[i] << 16), __uint_as_float(bb[i] << 16));
        const float y1 = fmaf((v[c][2 * i + 1] - mean) * rstd, __uint_as_float(ww[i] & 0xffff0000u), __uint_as_float(bb[i] & 0xffff0000u));
        ow[i] = pack_bf16x2(y0, y1);
      }
    }
    *reinterpret_cast<uint4*>(out + col) = o;
  }
}

template <int NCH>
static void launch_ln(const LnParams& p, int grid, cudaStream_t stream) {
  const bool xb = p.x_dtype == SA_BF16, ob = p.out_dtype == SA_BF16, wb = p.w_dtype == SA_BF16;
  const int thr = WARPS_PER_BLOCK * 32;
#define SA_LN_CASE(X, O, W) layernorm_kernel<NCH, X, O, W><<<grid, thr, 0, stream>>>(p)
  if (xb && ob && wb) SA_LN_CASE(true, true, true);
  else if (xb && ob) SA_LN_CASE(true, true, false);
  else if (xb && wb) SA_LN_CASE(true, false, true);
  else if (xb) SA_LN_CASE(true, false, false);
  else if (ob && wb) SA_LN_CASE(false, true, true);
  else if (ob) SA_LN_CASE(false, true, false);
  else if (wb) SA_LN_CASE(false, false, true);
  else SA_LN_CASE(false, false, false);
#undef SA_LN_CASE
}

// ------------------------------------------------------------------------------------------------ RMSNorm + RoPE
struct RmsParams {
  __nv_bfloat16* x[2];
  const __nv_bfloat16* weight[2];
  const float2* freqs;  // [1024][64] (cos, sin), NULL = no rotation
  long long ld;
  int rows, C, rows_per_batch, F, H, W, tok_offset;
  float eps;
};

template <int NCH>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, NCH <= 6 ? 4 : (NCH <= 8 ? 3 : 1)) rmsnorm_rope_kernel(const RmsParams p) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  __nv_bfloat16* x = p.x[blockIdx.y];
  const __nv_bfloat16* wgt = p.weight[blockIdx.y];
  const int nchunks = p.C >> 3;
  RowChunk<true> v[NCH];
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int ch = lane + c * 32;
    if (ch < nchunks) v[c].load(x, (long long)row * p.ld + ch * 8);
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (lane + c * 32 < nchunks) {
      float t[8];
      v[c].get(t);
#pragma unroll
      for (int i = 0; i < 8; ++i) ss += t[i] * t[i];
    }
  }
  const float rinv = rsqrtf(warp_sum(ss) / p.C + p.eps);
  const int tok = row % p.rows_per_batch + p.tok_offset;
  const bool rotate = p.freqs != nullptr && tok < p.F * p.H * p.W;
  const int pf = tok / (p.H * p.W), ph = (tok / p.W) % p.H, pw = tok % p.W;
  // chunk ch = lane + 32 c sits at column (ch & 15) * 8 of its head: a lane meets the SAME four complex pairs in every
  // chunk it owns, so their (cos, sin) are fetched once per row instead of once per chunk
  float2 cs4[4];
  if (rotate) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = (lane & 15) * 4 + i;
      const int pos = j < 22 ? pf : (j < 43 ? ph : pw);
      cs4[i] = __ldg(&p.freqs[pos * 64 + j]);
    }
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int ch = lane + c * 32;
    if (ch >= nchunks) continue;
    const int col = ch * 8;
    float w[8], y[8];
    load8_bf16(wgt + col, w);
    v[c].get(y);
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = bf16_round(bf16_round(y[i] * rinv) * w[i]);
    if (rotate) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 cs = cs4[i];
        const float a = y[2 * i], b = y[2 * i + 1];
        y[2 * i] = a * cs.x - b * cs.y;
        y[2 * i + 1] = a * cs.y + b * cs.x;
      }
    }
    store_chunk(x, SA_BF16, (long long)row * p.ld + col, y);
  }
}

// ------------------------------------------------------------------------------------------------ RMSNorm + RoPE + scatter
// Sequence-parallel producer fusion: the same arithmetic as rmsnorm_rope_kernel on the q and k parts of a fused QKV row
// (v is copied), but every 16-byte chunk is stored straight into its destination rank's receive buffer over NVLink
// (layouts and destination arithmetic of sp_exchange.cu's scatter_qkv_kernel) instead of back into the local row —
// the normalised q / k never make a local round trip and the separate scatter launch disappears.
struct RmsScatterParams {
  const __nv_bfloat16* qkv;        // [B*Ll, ld] rows: q | k | v, each C = heads * 128 wide
  const __nv_bfloat16* weight[2];  // norm_q, norm_k
  const float2* freqs;
  uint4* kv_dst[8];
  uint4* q_dst[8];
  long long ld;
  int rows, C, Ll, B, F, H, W, tok_offset;
  int nh, P, rank, qs, hp;
  int b_first;               // first CFG sample of this launch (rows = b_count * Ll)
  float eps;
};

template <int NCH>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, NCH <= 6 ? 4 : (NCH <= 8 ? 3 : 1)) rmsnorm_rope_scatter_kernel(const RmsScatterParams p) {
  const int lane = threadIdx.x & 31;
  const int row_l = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (row_l >= p.rows) return;
  const int row = p.b_first * p.Ll + row_l;
  const int which = blockIdx.y;  // 0 q, 1 k, 2 v
  const __nv_bfloat16* x = p.qkv + which * p.C;
  const int nchunks = p.C >> 3;
  RowChunk<true> v[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int ch = lane + c * 32;
    if (ch < nchunks) v[c].load(x, (long long)row * p.ld + ch * 8);
  }
  float rinv = 0.f;
  if (which < 2) {
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      if (lane + c * 32 < nchunks) {
        float t[8];
        v[c].get(t);
#pragma unroll
        for (int i = 0; i < 8; ++i) ss += t[i] * t[i];
      }
    }
    rinv = rsqrtf(warp_sum(ss) / p.C + p.eps);
  }
  const int t_loc = row % p.Ll, b = row / p.Ll;
  const int tok = t_loc + p.tok_offset;
  const bool rotate = which < 2 && p.freqs != nullptr && tok < p.F * p.H * p.W;
  const int pf = tok / (p.H * p.W), ph = (tok / p.W) % p.H, pw = tok % p.W;
  const int s_me = p.rank % p.qs;
  const long long q_row = ((long long)b * (p.P / p.qs) + p.rank / p.qs) * p.Ll + t_loc;   // row index inside q_recv
  const long long kv_row = ((long long)b * p.P + p.rank) * p.Ll + t_loc;                  // row index inside kv_recv
  float2 cs4[4];   // the four complex pairs this lane meets in every chunk it owns (see rmsnorm_rope_kernel)
  if (rotate) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = (lane & 15) * 4 + i;
      const int pos = j < 22 ? pf : (j < 43 ? ph : pw);
      cs4[i] = __ldg(&p.freqs[pos * 64 + j]);
    }
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int ch = lane + c * 32;
    if (ch >= nchunks) continue;
    const int col = ch * 8;
    uint4 outv;
    if (which < 2) {
      float w[8], y[8];
      load8_bf16(p.weight[which] + col, w);
      v[c].get(y);
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = bf16_round(bf16_round(y[i] * rinv) * w[i]);
      if (rotate) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 cs = cs4[i];
          const float a = y[2 * i], bb = y[2 * i + 1];
          y[2 * i] = a * cs.x - bb * cs.y;
          y[2 * i + 1] = a * cs.y + bb * cs.x;
        }
      }
      outv.x = pack_bf16x2(y[0], y[1]);
      outv.y = pack_bf16x2(y[2], y[3]);
      outv.z = pack_bf16x2(y[4], y[5]);
      outv.w = pack_bf16x2(y[6], y[7]);
    } else {
      outv = v[c].u;
    }
    const int h = ch >> 4, c16 = ch & 15;   // 16 chunks of 16 bytes per 128-wide head
    const int g = h / p.hp, hl = h % p.hp;
    if (which == 0) {
      p.q_dst[g * p.qs + s_me][(q_row * p.hp + hl) * 16 + c16] = outv;
    } else {
      const long long off = ((kv_row * 2 + (which - 1)) * p.hp + hl) * 16 + c16;
      for (int s = 0; s < p.qs; ++s) p.kv_dst[g * p.qs + s][off] = outv;
    }
  }
}

// ------------------------------------------------------------------------------------------------ broadcast add
// out[i, j, :] = bf16(a[i, :] + b[j, :])   — e = modulation + e0 for every block at once (1B.py:672)
__global__ void add_bcast_kernel(const __nv_bfloat16* a, const __nv_bfloat16* b, __nv_bfloat16* out, int na, int nb,
                                 int n) {
  const long long total = (long long)na * nb * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int col = idx % n;
    const long long r = idx / n;
    const int j = r % nb, i = r / nb;
    out[idx] = __float2bfloat16_rn(__bfloat162float(a[(long long)i * n + col]) + __bfloat162float(b[(long long)j * n + col]));
  }
}

}  // namespace norm
}  // namespace sa

extern "C" int sa_layernorm_modulate(const sa_ln_args* a, sa_stream_t stream_) {
  using namespace sa;
  using namespace sa::norm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !a->x || !a->out) { set_error("sa_layernorm_modulate: null pointer"); return SA_ERR_BAD_ARG; }
  if (a->rows <= 0) return SA_OK;
  if (a->C <= 0 || a->C % 8 || a->C > MAX_CHUNKS * 256 || a->ldx % 8 || a->ldo % 8) {
    set_error("sa_layernorm_modulate: C must be a multiple of 8 and <= %d, strides multiples of 8 (C=%d)",
              MAX_CHUNKS * 256, a->C);
    return SA_ERR_BAD_ARG;
  }
  if ((a->weight == nullptr) != (a->bias == nullptr) || (a->shift == nullptr) != (a->scale == nullptr) ||
      (a->gate && !a->res)) {
    set_error("sa_layernorm_modulate: weight/bias, shift/scale must come in pairs; gate needs res");
    return SA_ERR_BAD_ARG;
  }
  LnParams p;
  p.x = a->x; p.out = a->out; p.weight = a->weight; p.bias = a->bias;
  p.shift = reinterpret_cast<const __nv_bfloat16*>(a->shift);
  p.scale = reinterpret_cast<const __nv_bfloat16*>(a->scale);
  p.gate = reinterpret_cast<const __nv_bfloat16*>(a->gate);
  p.res = a->res;
  p.ldx = a->ldx; p.ldo = a->ldo; p.ldr = a->ldr; p.mod_bs = a->mod_bs;
  p.rows = a->rows; p.C = a->C; p.rows_per_batch = a->rows_per_batch > 0 ? a->rows_per_batch : a->rows;
  p.x_dtype = a->x_dtype; p.out_dtype = a->out_dtype; p.w_dtype = a->w_dtype; p.round_bf16 = a->round_bf16;
  p.eps = a->eps;
  const int grid = (a->rows + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
  const int nch = (a->C / 8 + 31) / 32;
  // bf16 DiT fast paths (packed bf16 modulation chain / bf16 affine), rows of up to 2048 channels held in fp32 registers
  const bool bf_io = a->x_dtype == SA_BF16 && a->out_dtype == SA_BF16 && a->round_bf16 && !a->gate && !a->res && nch <= 8;
  const int thr = WARPS_PER_BLOCK * 32;
  if (bf_io && a->scale && !a->weight) {
    if (nch <= 2) layernorm_bf16_kernel<2, 1><<<grid, thr, 0, stream>>>(p);
    else if (nch <= 6) layernorm_bf16_kernel<6, 1><<<grid, thr, 0, stream>>>(p);
    else layernorm_bf16_kernel<8, 1><<<grid, thr, 0, stream>>>(p);
  } else if (bf_io && a->weight && a->w_dtype == SA_BF16 && !a->scale) {
    if (nch <= 2) layernorm_bf16_kernel<2, 2><<<grid, thr, 0, stream>>>(p);
    else if (nch <= 6) layernorm_bf16_kernel<6, 2><<<grid, thr, 0, stream>>>(p);
    else layernorm_bf16_kernel<8, 2><<<grid, thr, 0, stream>>>(p);
  } else if (nch <= 2) launch_ln<2>(p, grid, stream);
  else if (nch <= 6) launch_ln<6>(p, grid, stream);
  else if (nch <= 8) launch_ln<8>(p, grid, stream);
  else launch_ln<20>(p, grid, stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "layernorm_kernel launch");
  return SA_OK;
}

extern "C" int sa_rmsnorm_rope(const sa_rms_args* a, sa_stream_t stream_) {
  using namespace sa;
  using namespace sa::norm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !a->x || !a->weight) { set_error("sa_rmsnorm_rope: null pointer"); return SA_ERR_BAD_ARG; }
  if (a->rows <= 0) return SA_OK;
  if (a->C <= 0 || a->C % 8 || a->C > MAX_CHUNKS * 256 || a->ld % 8) {
    set_error("sa_rmsnorm_rope: C must be a multiple of 8 and <= %d (C=%d)", MAX_CHUNKS * 256, a->C);
    return SA_ERR_BAD_ARG;
  }
  if (a->freqs && (a->C % 128 || a->F > 1024 || a->H > 1024 || a->W > 1024 || a->F <= 0 || a->H <= 0 || a->W <= 0)) {
    set_error("sa_rmsnorm_rope: RoPE needs head_dim 128 and 0 < F,H,W <= 1024");
    return SA_ERR_BAD_ARG;
  }
  if ((a->x2 == nullptr) != (a->weight2 == nullptr)) { set_error("sa_rmsnorm_rope: x2/weight2 pair"); return SA_ERR_BAD_ARG; }
  RmsParams p;
  p.x[0] = reinterpret_cast<__nv_bfloat16*>(a->x);
  p.x[1] = reinterpret_cast<__nv_bfloat16*>(a->x2);
  p.weight[0] = reinterpret_cast<const __nv_bfloat16*>(a->weight);
  p.weight[1] = reinterpret_cast<const __nv_bfloat16*>(a->weight2);
  p.freqs = reinterpret_cast<const float2*>(a->freqs);
  p.ld = a->ld; p.rows = a->rows; p.C = a->C;
  p.rows_per_batch = a->rows_per_batch > 0 ? a->rows_per_batch : a->rows;
  p.F = a->F; p.H = a->H; p.W = a->W; p.tok_offset = a->tok_offset; p.eps = a->eps;
  dim3 grid((a->rows + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, a->x2 ? 2 : 1);
  const int nch = (a->C / 8 + 31) / 32;
  if (nch <= 2) rmsnorm_rope_kernel<2><<<grid, WARPS_PER_BLOCK * 32, 0, stream>>>(p);
  else if (nch <= 6) rmsnorm_rope_kernel<6><<<grid, WARPS_PER_BLOCK * 32, 0, stream>>>(p);
  else if (nch <= 8) rmsnorm_rope_kernel<8><<<grid, WARPS_PER_BLOCK * 32, 0, stream>>>(p);
  else rmsnorm_rope_kernel<20><<<grid, WARPS_PER_BLOCK * 32, 0, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "rmsnorm_rope_kernel launch");
  return SA_OK;
}

extern "C" int sa_add_bcast_bf16(const void* a, const void* b, void* out, int32_t na, int32_t nb, int32_t n,
                                 sa_stream_t stream_) {
  using namespace sa;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !b || !out || na <= 0 || nb <= 0 || n <= 0) { set_error("sa_add_bcast_bf16: bad argument"); return SA_ERR_BAD_ARG; }
  const long long total = (long long)na * nb * n;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  norm::add_bcast_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(a),
                                                   reinterpret_cast<const __nv_bfloat16*>(b),
                                                   reinterpret_cast<__nv_bfloat16*>(out), na, nb, n);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "add_bcast_kernel launch");
  return SA_OK;
}

extern "C" int sa_sp_norm_rope_scatter(const sa_sp_args* a, const void* weight_q, const void* weight_k, const void* freqs,
                                       int32_t F, int32_t H, int32_t W, int32_t tok_offset, float eps, sa_stream_t stream_) {
  using namespace sa;
  using namespace sa::norm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !a->src || !weight_q || !weight_k || a->P < 1 || a->P > 8 || a->rank < 0 || a->rank >= a->P || a->B <= 0 || a->Ll <= 0 ||
      a->heads <= 0 || a->hg <= 0 || a->P % a->hg || a->heads % a->hg || a->head_dim != 128) {
    set_error("sa_sp_norm_rope_scatter: bad argument (1 <= P <= 8, hg | P, hg | heads, head_dim 128)");
    return SA_ERR_BAD_ARG;
  }
  const int C = a->heads * 128;
  if (C > MAX_CHUNKS * 256 || a->ld % 8 || a->ld < 3LL * C) {
    set_error("sa_sp_norm_rope_scatter: heads * 128 <= %d, ld %% 8 == 0, ld >= 3 * heads * 128", MAX_CHUNKS * 256);
    return SA_ERR_BAD_ARG;
  }
  if (freqs && (F <= 0 || H <= 0 || W <= 0 || F > 1024 || H > 1024 || W > 1024)) { set_error("sa_sp_norm_rope_scatter: bad RoPE grid"); return SA_ERR_BAD_ARG; }
  RmsScatterParams p;
  p.qkv = reinterpret_cast<const __nv_bfloat16*>(a->src);
  p.weight[0] = reinterpret_cast<const __nv_bfloat16*>(weight_q);
  p.weight[1] = reinterpret_cast<const __nv_bfloat16*>(weight_k);
  p.freqs = reinterpret_cast<const float2*>(freqs);
  for (int r = 0; r < 8; ++r) {
    p.kv_dst[r] = r < a->P ? reinterpret_cast<uint4*>(a->dst_a[r]) : nullptr;
    p.q_dst[r] = r < a->P ? reinterpret_cast<uint4*>(a->dst_b[r]) : nullptr;
    if (r < a->P && (!p.kv_dst[r] || !p.q_dst[r])) { set_error("sa_sp_norm_rope_scatter: null destination for rank %d", r); return SA_ERR_BAD_ARG; }
  }
  const int b_count = a->b_count > 0 ? a->b_count : a->B - a->b_first;
  if (a->b_first < 0 || a->b_first + b_count > a->B) { set_error("sa_sp_norm_rope_scatter: sample range outside the batch"); return SA_ERR_BAD_ARG; }
  p.b_first = a->b_first;
  p.ld = a->ld; p.rows = b_count * a->Ll; p.C = C; p.Ll = a->Ll; p.B = a->B; p.F = F; p.H = H; p.W = W; p.tok_offset = tok_offset;
  p.nh = a->heads; p.P = a->P; p.rank = a->rank; p.qs = a->P / a->hg; p.hp = a->heads / a->hg; p.eps = eps;
  dim3 grid((p.rows + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, 3);
  const int nch = (C / 8 + 31) / 32;
  if (nch <= 2) rmsnorm_rope_scatter_kernel<2><<<grid, WARPS_PER_BLOCK * 32, 0, stream>>>(p);
  else if (nch <= 6) rmsnorm_rope_scatter_kernel<6><<<grid, WARPS_PER_BLOCK * 32, 0, stream>>>(p);
  else if (nch <= 8) rmsnorm_rope_scatter_kernel<8><<<grid, WARPS_PER_BLOCK * 32, 0, stream>>>(p);
  else rmsnorm_rope_scatter_kernel<20><<<grid, WARPS_PER_BLOCK * 32, 0, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "rmsnorm_rope_scatter_kernel launch");
  return SA_OK;
}
