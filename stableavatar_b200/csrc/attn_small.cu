// attn_small.cu — attention with a handful of queries per (batch, head): the audio adapter's cross-attention, where the
// 15 audio tokens of one latent frame attend to that frame's 1560 video tokens with 8 heads of 192
// (wan/models/vocal_projector_fantasy_1B.py:259-277, SDPA branch :178-203). 3 GFLOP per denoise step in total, so this is
// a CUDA-core kernel bounded by the single HBM pass over K and V: one CTA per (batch, head); scores for all queries
// live in shared memory (q_len x key-tile fp32); K rows are read once (one thread per key), V rows once (one thread per
// output column, coalesced across the head dim). The 14B adapter (vocal_projector_fantasy_14B.py, 8 heads of 640,
// 3600 keys per frame at 720x1280) walks the keys in tiles with a running softmax.
#include "../../include/stableavatar_b200.h"
#include "sa_host.h"
#include "sa_ptx.cuh"

namespace sa {
namespace attn_small {

constexpr int MAXQ_LIMIT = 32;   // queries per (batch, head): 15 for 81-frame windows fed 161 wav2vec tokens, 17-19 when the
                                 // pipeline feeds the window's 84 / 12 frames of audio (pipe.py:722-724)
constexpr int THREADS = 256;
constexpr int MAXCPT = 3;  // output columns per thread: head_dim <= 768 (192 for the 1.3B adapter, 640 for the 14B one)

struct Params {
  const __nv_bfloat16* q; const __nv_bfloat16* k; const __nv_bfloat16* v; __nv_bfloat16* out;
  long long q_bs, q_ls, k_bs, k_ls, v_bs, v_ls, o_bs, o_ls;
  int heads, q_len, kv_len, d, tk, ts;  // tk: keys per tile (scores of one tile live in shared memory), ts: row stride (tk rounded up to 4)
  float scale;
};

// Keys are walked in tiles of `tk` with the usual running (max, sum) rescale, so kv_len is unbounded; when the whole
// row of scores fits (the 1.3B shapes) there is a single tile and the rescale factors are all 1.
template <int MAXQ>
__global__ void __launch_bounds__(THREADS) attn_small_kernel(const Params p) {
  extern __shared__ float smem[];
  float* sq = smem;                      // [q_len][d]
  float* ss = sq + p.q_len * p.d;        // [q_len][ts]
  float* s_m = ss + p.q_len * p.ts;      // [MAXQ] running max
  float* s_l = s_m + MAXQ;               // [MAXQ] running sum
  float* s_f = s_l + MAXQ;               // [MAXQ] rescale factor of the current tile
  const int h = blockIdx.x % p.heads, b = blockIdx.x / p.heads;
  const __nv_bfloat16* qb = p.q + (long long)b * p.q_bs + h * p.d;
  const __nv_bfloat16* kb = p.k + (long long)b * p.k_bs + h * p.d;
  const __nv_bfloat16* vb = p.v + (long long)b * p.v_bs + h * p.d;
  for (int i = threadIdx.x; i < p.q_len * p.d; i += THREADS)
    sq[i] = __bfloat162float(qb[(long long)(i / p.d) * p.q_ls + (i % p.d)]) * p.scale;
  if (threadIdx.x < MAXQ) {
    s_m[threadIdx.x] = -INFINITY;
    s_l[threadIdx.x] = 0.f;
  }
  float acc[MAXCPT][MAXQ];
#pragma unroll
  for (int cc = 0; cc < MAXCPT; ++cc)
#pragma unroll
    for (int i = 0; i < MAXQ; ++i) acc[cc][i] = 0.f;
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t0 = 0; t0 < p.kv_len; t0 += p.tk) {
    const int tn = min(p.tk, p.kv_len - t0);
    // S = (q * scale) K^T : one thread per key
    for (int j = threadIdx.x; j < tn; j += THREADS) {
      float sc[MAXQ];
#pragma unroll
      for (int i = 0; i < MAXQ; ++i) sc[i] = 0.f;
      const uint4* kr = reinterpret_cast<const uint4*>(kb + (long long)(t0 + j) * p.k_ls);
      for (int c = 0; c < p.d / 8; ++c) {
        const uint4 u = kr[c];
        const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&u);
        float kv[8];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = __bfloat1622float2(hh[t]);
          kv[2 * t] = f.x;
          kv[2 * t + 1] = f.y;
        }
#pragma unroll
        for (int i = 0; i < MAXQ; ++i) {
          if (i < p.q_len) {
            // two 16-byte broadcast reads (d % 8 == 0 keeps every chunk 32-byte aligned) instead of eight scalar ones
            const float4 qa = *reinterpret_cast<const float4*>(sq + i * p.d + c * 8);
            const float4 qb4 = *reinterpret_cast<const float4*>(sq + i * p.d + c * 8 + 4);
            sc[i] = fmaf(qa.x, kv[0], fmaf(qa.y, kv[1], fmaf(qa.z, kv[2], fmaf(qa.w, kv[3], sc[i]))));
            sc[i] = fmaf(qb4.x, kv[4], fmaf(qb4.y, kv[5], fmaf(qb4.z, kv[6], fmaf(qb4.w, kv[7], sc[i]))));
          }
        }
      }
#pragma unroll
      for (int i = 0; i < MAXQ; ++i)
        if (i < p.q_len) ss[i * p.ts + j] = sc[i];
    }
    __syncthreads();

    // running softmax per query row: one warp per row
    for (int i = warp; i < p.q_len; i += THREADS / 32) {
      float* row = ss + i * p.ts;
      float mx = -INFINITY;
      for (int j = lane; j < tn; j += 32) mx = fmaxf(mx, row[j]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float m_old = s_m[i];
      const float m_new = fmaxf(m_old, mx);
      float sum = 0.f;
      for (int j = lane; j < tn; j += 32) {
        const float e = __expf(row[j] - m_new);
        row[j] = e;
        sum += e;
      }
      if (lane < 3 && tn + lane < ((tn + 3) & ~3)) row[tn + lane] = 0.f;   // the P V loop reads rows four keys at a time
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0) {
        const float f = __expf(m_old - m_new);  // 0 on the first tile (m_old = -inf)
        s_f[i] = f;
        s_l[i] = s_l[i] * f + sum;
        s_m[i] = m_new;
      }
    }
    __syncthreads();

    // O = O * f + P V : one thread per output column (up to MAXCPT columns per thread), V rows read coalesced
#pragma unroll
    for (int cc = 0; cc < MAXCPT; ++cc) {
      const int c = threadIdx.x + cc * THREADS;
      if (c < p.d) {
#pragma unroll
        for (int i = 0; i < MAXQ; ++i)
          if (i < p.q_len) acc[cc][i] *= s_f[i];
        // four keys per iteration: four independent V loads in flight and one 16-byte (broadcast) read of each query's
        // probabilities instead of four scalar ones — the scalar form took 1.85 ms per launch for 0.6 GB of K / V
        const int tn4 = (tn + 3) & ~3;
#pragma unroll 2
        for (int j = 0; j < tn4; j += 4) {
          float vv[4];
#pragma unroll
          for (int t = 0; t < 4; ++t)
            vv[t] = (j + t < tn) ? __bfloat162float(vb[(long long)(t0 + j + t) * p.v_ls + c]) : 0.f;
#pragma unroll
          for (int i = 0; i < MAXQ; ++i) {
            if (i < p.q_len) {
              const float4 pr = *reinterpret_cast<const float4*>(&ss[i * p.ts + j]);
              acc[cc][i] = fmaf(pr.x, vv[0], fmaf(pr.y, vv[1], fmaf(pr.z, vv[2], fmaf(pr.w, vv[3], acc[cc][i]))));
            }
          }
        }
      }
    }
    __syncthreads();
  }

#pragma unroll
  for (int cc = 0; cc < MAXCPT; ++cc) {
    const int c = threadIdx.x + cc * THREADS;
    if (c < p.d) {
      __nv_bfloat16* ob = p.out + (long long)b * p.o_bs + h * p.d + c;
#pragma unroll
      for (int i = 0; i < MAXQ; ++i)
        if (i < p.q_len) ob[(long long)i * p.o_ls] = __float2bfloat16_rn(acc[cc][i] / s_l[i]);
    }
  }
}

}  // namespace attn_small
}  // namespace sa

extern "C" int sa_attn_small_q(const sa_attn_args* a, int32_t head_dim, sa_stream_t stream_) {
  using namespace sa;
  using namespace sa::attn_small;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !a->q || !a->k || !a->v || !a->out) { set_error("sa_attn_small_q: null pointer"); return SA_ERR_BAD_ARG; }
  if (a->batch <= 0 || a->heads <= 0 || a->q_len <= 0 || a->kv_len <= 0 || head_dim <= 0 || head_dim % 8) {
    set_error("sa_attn_small_q: bad dims");
    return SA_ERR_BAD_ARG;
  }
  if (a->k_ls % 8 || a->k_bs % 8) { set_error("sa_attn_small_q: k strides must be multiples of 8"); return SA_ERR_BAD_ARG; }
  if (a->q_len > MAXQ_LIMIT || head_dim > MAXCPT * THREADS) {
    set_error("sa_attn_small_q: unsupported shape (q_len %d > %d or head_dim %d > %d)", a->q_len, MAXQ_LIMIT, head_dim, MAXCPT * THREADS);
    return SA_ERR_UNSUPPORTED;
  }
  // keys per tile: whatever of 200 KB the queries leave, in multiples of 32
  const int MAXQ = a->q_len <= 16 ? 16 : 32;
  const size_t fixed = ((size_t)a->q_len * head_dim + 3 * MAXQ) * sizeof(float);
  int tk = (int)((200 * 1024 - fixed) / (a->q_len * sizeof(float))) / 32 * 32 - 32;
  if (tk > a->kv_len) tk = a->kv_len;
  const int ts = (tk + 3) & ~3;
  const size_t smem = fixed + (size_t)a->q_len * ts * sizeof(float);
  if (a->accumulate) { set_error("sa_attn_small_q: accumulate not supported"); return SA_ERR_UNSUPPORTED; }
  Params p;
  p.q = reinterpret_cast<const __nv_bfloat16*>(a->q);
  p.k = reinterpret_cast<const __nv_bfloat16*>(a->k);
  p.v = reinterpret_cast<const __nv_bfloat16*>(a->v);
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  p.q_bs = a->q_bs; p.q_ls = a->q_ls; p.k_bs = a->k_bs; p.k_ls = a->k_ls;
  p.v_bs = a->v_bs; p.v_ls = a->v_ls; p.o_bs = a->o_bs; p.o_ls = a->o_ls;
  p.heads = a->heads; p.q_len = a->q_len; p.kv_len = a->kv_len; p.d = head_dim; p.tk = tk; p.ts = ts; p.scale = a->scale;
  if (MAXQ == 16) {
    if (int rc = ensure_dyn_smem(attn_small_kernel<16>, 200 * 1024, "attn_small_kernel")) return rc;
    attn_small_kernel<16><<<a->batch * a->heads, THREADS, smem, stream>>>(p);
  } else {
    if (int rc = ensure_dyn_smem(attn_small_kernel<32>, 200 * 1024, "attn_small_kernel")) return rc;
    attn_small_kernel<32><<<a->batch * a->heads, THREADS, smem, stream>>>(p);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "attn_small_kernel launch");
  return SA_OK;
}
