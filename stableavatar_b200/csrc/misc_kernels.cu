// misc_kernels.cu — the small memory-bound ops around the block stack: patchify gather (patch-embedding as a GEMM),
// unpatchify scatter, the fp32 time-embedding MLP rows, audio-window row gather, and the fused CFG + Euler step.
#include "../../include/stableavatar_b200.h"
#include "sa_host.h"
#include "sa_ptx.cuh"

namespace sa {
namespace misc {

// ------------------------------------------------------------------------------------------------ patchify
// A[b, tok, c*4 + q*2 + r] = cat(x, y)[b, c, f, 2h+q, 2w+r], tok = (f*Hp + h)*Wp + w. With this K order the Conv3d
// weight [dim, C, 1, 2, 2] flattened to [dim, C*4] is the GEMM's W operand unchanged (1B.py:830-831, 972-978).
// Rows tok >= F*Hp*Wp (zero padding up to seq_len, 1B.py:983) are written as zeros.
struct PatchParams {
  const __nv_bfloat16* x; const __nv_bfloat16* y; __nv_bfloat16* out;
  int B, Cx, Cy, F, H, W, seq_len, K_pad;
};
__global__ void patchify_kernel(const PatchParams p) {
  const int Hp = p.H / 2, Wp = p.W / 2;
  const int C = p.Cx + p.Cy;
  const long long total = (long long)p.B * p.seq_len * p.K_pad;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k = idx % p.K_pad;
    const long long rt = idx / p.K_pad;
    const int tok = rt % p.seq_len, b = rt / p.seq_len;
    __nv_bfloat16 v = __float2bfloat16_rn(0.f);
    if (tok < p.F * Hp * Wp && k < C * 4) {
      const int c = k >> 2, q = (k >> 1) & 1, r = k & 1;
      const int w = tok % Wp, h = (tok / Wp) % Hp, f = tok / (Wp * Hp);
      const long long sp = ((long long)f * p.H + (2 * h + q)) * p.W + (2 * w + r);
      const long long fhw = (long long)p.F * p.H * p.W;
      v = c < p.Cx ? p.x[((long long)b * p.Cx + c) * fhw + sp] : p.y[((long long)b * p.Cy + (c - p.Cx)) * fhw + sp];
    }
    p.out[idx] = v;
  }
}

// ------------------------------------------------------------------------------------------------ unpatchify
// out[b, c, f, 2h+q, 2w+r] = u[b, tok, (q*2 + r)*Cout + c]   (1B.py:1177-1183, einsum 'fhwpqrc->cfphqwr')
struct UnpatchParams {
  const __nv_bfloat16* u; __nv_bfloat16* out;
  long long u_bs, u_ls;
  int B, Cout, F, H, W;
};
__global__ void unpatchify_kernel(const UnpatchParams p) {
  const int Hp = p.H / 2, Wp = p.W / 2;
  const long long total = (long long)p.B * p.Cout * p.F * p.H * p.W;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x = idx % p.W;
    long long r = idx / p.W;
    const int y = r % p.H; r /= p.H;
    const int f = r % p.F; r /= p.F;
    const int c = r % p.Cout;
    const int b = r / p.Cout;
    const int tok = (f * Hp + (y >> 1)) * Wp + (x >> 1);
    const int j = ((y & 1) * 2 + (x & 1)) * p.Cout + c;
    p.out[idx] = p.u[(long long)b * p.u_bs + (long long)tok * p.u_ls + j];
  }
}

// ------------------------------------------------------------------------------------------------ small fp32 linear
// out[m, n] = sum_k pre(x[m, k]) * W[n, k] + bias[n], fp32 math, bf16 weights: the time-embedding island
// (1B.py:986-990: autocast(dtype=float32)). pre: 0 none, 1 SiLU, 2 sinusoid of t[m] (1B.py:210-220, fp64).
// One warp per output feature n; M <= 8 rows kept in registers.
struct SmallLinParams {
  const float* x; const void* w; const void* bias; float* out; __nv_bfloat16* out_bf16;
  int M, N, K, pre, w_dtype;
};
__global__ void small_linear_kernel(const SmallLinParams p) {
  extern __shared__ float xs[];  // [M][K] pre-activated input
  for (int i = threadIdx.x; i < p.M * p.K; i += blockDim.x) {
    const int m = i / p.K, k = i % p.K;
    float v;
    if (p.pre == 2) {
      const int half = p.K / 2;
      const int kk = k < half ? k : k - half;
      const double ang = (double)p.x[m] * pow(10000.0, -(double)kk / (double)half);
      v = (float)(k < half ? cos(ang) : sin(ang));
    } else {
      v = p.x[i];
      if (p.pre == 1) v = v / (1.0f + expf(-v));
    }
    xs[i] = v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= p.N) return;
  float acc[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) acc[m] = 0.f;
  for (int k = lane; k < p.K; k += 32) {
    const float w = p.w_dtype == SA_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.w)[(long long)n * p.K + k])
                                         : reinterpret_cast<const float*>(p.w)[(long long)n * p.K + k];
#pragma unroll
    for (int m = 0; m < 8; ++m)
      if (m < p.M) acc[m] = fmaf(xs[m * p.K + k], w, acc[m]);
  }
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    if (m >= p.M) break;
    float v = acc[m];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) {
      if (p.bias)
        v += p.w_dtype == SA_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.bias)[n])
                                  : reinterpret_cast<const float*>(p.bias)[n];
      if (p.out) p.out[(long long)m * p.N + n] = v;
      if (p.out_bf16) p.out_bf16[(long long)m * p.N + n] = __float2bfloat16_rn(v);
    }
  }
}

// ------------------------------------------------------------------------------------------------ row gather
// out[r, :] = idx[r] >= 0 ? src[idx[r], :] : 0    (audio windows, vp.py:81-131: right-zero-padded gathers)
__global__ void gather_rows_kernel(const float* src, const int* idx, float* out, int rows, int C) {
  const long long total = (long long)rows * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = i / C, c = i % C;
    const int s = idx[r];
    out[i] = s >= 0 ? src[(long long)s * C + c] : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------ CFG + Euler
// noise = uncond + a*(drop_audio - uncond) + t*(cond - drop_audio)   every op a bf16 tensor op (pipe.py:752-753)
// latents = bf16(float(latents) + bf16(dsigma * noise))             (diffusers FlowMatchEulerDiscreteScheduler.step)
template <typename LatT>
__global__ void cfg_euler_kernel(const __nv_bfloat16* pred, const LatT* lat, __nv_bfloat16* out,
                                 __nv_bfloat16* noise_out, long long n, float audio_scale, float text_scale,
                                 float dsigma, const float* dsigma_dev, int cfg) {
  if (dsigma_dev) dsigma = *dsigma_dev;   // schedule value kept on the device so that a captured CUDA graph can be replayed
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float np;
    if (cfg) {
      const float u = __bfloat162float(pred[i]), d = __bfloat162float(pred[n + i]), c = __bfloat162float(pred[2 * n + i]);
      const float r2 = bf16_round(audio_scale * bf16_round(d - u));
      const float r3 = bf16_round(u + r2);
      const float r5 = bf16_round(text_scale * bf16_round(c - d));
      np = bf16_round(r3 + r5);
    } else {
      np = __bfloat162float(pred[i]);
    }
    if (noise_out) noise_out[i] = __float2bfloat16_rn(np);
    // (sigma_next - sigma) is a 0-dim fp32 tensor: its product with the bf16 prediction is a bf16 tensor (type
    // promotion ignores 0-dim operands of the same category); the add with the fp32 sample is fp32; no FMA contraction.
    const float step = bf16_round(__fmul_rn(dsigma, np));
    out[i] = __float2bfloat16_rn(__fadd_rn(static_cast<float>(lat[i]), step));
  }
}

// ------------------------------------------------------------------------------------------------ window blend (K16)
// All windows of one denoise step written back into pred_latents with the overlap blend, in window order, in ONE
// launch (pipe.py:756-779). Everything is elementwise over (channel, h, w) and the frame bookkeeping is static, so a
// thread owns one (channel, pixel) column and walks windows and frames sequentially, reproducing statement by statement:
//     latents[:, :, s_j] = latents[:, :, s_j] * w_j + pred_latents[:, :, e_j] * (1 - w_j)     s_j = j % f, e_j = (prev_end - n + j) % N
//     latents = latents.to(bf16); pred_latents[:, :, (ws + i) % N] = latents[:, :, i]
// with torch's roundings: latents and the weights are bf16, so every product / sum of bf16 tensors is rounded to bf16;
// with an fp32 pred_latents (caller-supplied fp32 noise) the second product and the sum are fp32.
constexpr int BLEND_MAX_WINDOWS = 64, BLEND_MAX_OVERLAP = 64;
struct WindowBlendParams {
  const __nv_bfloat16* new_lat;   // [W, C, f_max, HW]: window k's f[k] frames first
  void* pred;                     // [C, N, HW] bf16 or f32, zero-initialised by the caller
  int W, C, N, HW, f_max, overlap, pred_f32;
  int ws[BLEND_MAX_WINDOWS], f[BLEND_MAX_WINDOWS], prev_end[BLEND_MAX_WINDOWS], blend[BLEND_MAX_WINDOWS];
  float w[BLEND_MAX_OVERLAP], omw[BLEND_MAX_OVERLAP];   // bf16 values of w_j and of the bf16 tensor (1 - w)_j
};

template <typename PredT>
__global__ void window_blend_kernel(const __grid_constant__ WindowBlendParams p) {
  PredT* pred = reinterpret_cast<PredT*>(p.pred);
  const long long total = (long long)p.C * p.HW;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = idx / p.HW, x = idx % p.HW;
    PredT* pc = pred + (long long)c * p.N * p.HW + x;
    for (int k = 0; k < p.W; ++k) {
      const __nv_bfloat16* nw = p.new_lat + (((long long)k * p.C + c) * p.f_max) * p.HW + x;
      const int f = p.f[k];
      float tmp[BLEND_MAX_OVERLAP];
      if (p.blend[k]) {                                   // right-hand sides from the values before any assignment
        for (int j = 0; j < p.overlap; ++j) {
          const int e = ((p.prev_end[k] - p.overlap + j) % p.N + p.N) % p.N;
          const float a = bf16_round(__fmul_rn(__bfloat162float(nw[(long long)(j % f) * p.HW]), p.w[j]));
          const float pe = static_cast<float>(pc[(long long)e * p.HW]);
          const float b = sizeof(PredT) == 4 ? __fmul_rn(pe, p.omw[j]) : bf16_round(__fmul_rn(pe, p.omw[j]));
          tmp[j] = bf16_round(__fadd_rn(a, b));
        }
      }
      for (int i = 0; i < f; ++i) {
        float v = __bfloat162float(nw[(long long)i * p.HW]);
        if (p.blend[k] && i < p.overlap) v = tmp[i + f * ((p.overlap - 1 - i) / f)];   // the last j with j % f == i wins
        pc[(long long)((p.ws[k] + i) % p.N) * p.HW] = static_cast<PredT>(v);
      }
    }
  }
}

static inline int grid_for(long long total, int block = 256) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace misc
}  // namespace sa

#define SA_LAUNCH_CHECK(name)                                   \
  do {                                                          \
    cudaError_t e__ = cudaGetLastError();                       \
    if (e__ != cudaSuccess) return sa::cuda_fail(e__, name);    \
  } while (0)

extern "C" int sa_patchify(const void* x, const void* y, void* out, int32_t B, int32_t Cx, int32_t Cy, int32_t F,
                           int32_t H, int32_t W, int32_t seq_len, int32_t K_pad, sa_stream_t stream) {
  using namespace sa;
  if (!x || !out || (Cy > 0 && !y) || B <= 0 || F <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1) ||
      K_pad < (Cx + Cy) * 4 || seq_len < F * (H / 2) * (W / 2)) {
    set_error("sa_patchify: bad argument (H,W even; K_pad >= 4*(Cx+Cy); seq_len >= F*H/2*W/2)");
    return SA_ERR_BAD_ARG;
  }
  misc::PatchParams p{reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(y),
                      reinterpret_cast<__nv_bfloat16*>(out), B, Cx, Cy, F, H, W, seq_len, K_pad};
  misc::patchify_kernel<<<misc::grid_for((long long)B * seq_len * K_pad), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  SA_LAUNCH_CHECK("patchify_kernel launch");
  return SA_OK;
}

extern "C" int sa_unpatchify(const void* u, void* out, int64_t u_bs, int64_t u_ls, int32_t B, int32_t Cout, int32_t F,
                             int32_t H, int32_t W, sa_stream_t stream) {
  using namespace sa;
  if (!u || !out || B <= 0 || Cout <= 0 || F <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1)) {
    set_error("sa_unpatchify: bad argument");
    return SA_ERR_BAD_ARG;
  }
  misc::UnpatchParams p{reinterpret_cast<const __nv_bfloat16*>(u), reinterpret_cast<__nv_bfloat16*>(out), u_bs, u_ls, B, Cout, F, H, W};
  misc::unpatchify_kernel<<<misc::grid_for((long long)B * Cout * F * H * W), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  SA_LAUNCH_CHECK("unpatchify_kernel launch");
  return SA_OK;
}

extern "C" int sa_small_linear_f32(const void* x, const void* w, const void* bias, void* out_f32, void* out_bf16,
                                   int32_t M, int32_t N, int32_t K, int32_t pre, int32_t w_dtype, sa_stream_t stream) {
  using namespace sa;
  if (!x || !w || (!out_f32 && !out_bf16) || M <= 0 || M > 8 || N <= 0 || K <= 0 || pre < 0 || pre > 2 ||
      (size_t)M * K * 4 > 96 * 1024) {
    set_error("sa_small_linear_f32: bad argument (1 <= M <= 8, M*K*4 <= 96 KB)");
    return SA_ERR_BAD_ARG;
  }
  misc::SmallLinParams p{reinterpret_cast<const float*>(x), w, bias, reinterpret_cast<float*>(out_f32),
                         reinterpret_cast<__nv_bfloat16*>(out_bf16), M, N, K, pre, w_dtype};
  const size_t smem = (size_t)M * K * 4;
  if (int rc = ensure_dyn_smem(misc::small_linear_kernel, 96 * 1024, "small_linear_kernel")) return rc;
  misc::small_linear_kernel<<<(N + 7) / 8, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  SA_LAUNCH_CHECK("small_linear_kernel launch");
  return SA_OK;
}

extern "C" int sa_gather_rows_f32(const void* src, const void* idx, void* out, int32_t rows, int32_t C, sa_stream_t stream) {
  using namespace sa;
  if (!src || !idx || !out || rows <= 0 || C <= 0) { set_error("sa_gather_rows_f32: bad argument"); return SA_ERR_BAD_ARG; }
  misc::gather_rows_kernel<<<misc::grid_for((long long)rows * C), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float*>(src), reinterpret_cast<const int*>(idx), reinterpret_cast<float*>(out), rows, C);
  SA_LAUNCH_CHECK("gather_rows_kernel launch");
  return SA_OK;
}

extern "C" int sa_cfg_euler_step(const void* pred, const void* latents, void* out, void* noise_out, int64_t n,
                                 float audio_scale, float text_scale, float dsigma, const void* dsigma_dev, int32_t cfg,
                                 int32_t latents_dtype, sa_stream_t stream) {
  using namespace sa;
  if (!pred || !latents || !out || n <= 0 || (latents_dtype != SA_BF16 && latents_dtype != SA_F32)) {
    set_error("sa_cfg_euler_step: bad argument");
    return SA_ERR_BAD_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  auto* pr = reinterpret_cast<const __nv_bfloat16*>(pred);
  auto* o = reinterpret_cast<__nv_bfloat16*>(out);
  auto* no = reinterpret_cast<__nv_bfloat16*>(noise_out);
  auto* ds = reinterpret_cast<const float*>(dsigma_dev);
  if (latents_dtype == SA_F32)
    misc::cfg_euler_kernel<float><<<misc::grid_for(n), 256, 0, st>>>(pr, reinterpret_cast<const float*>(latents), o, no, n,
                                                                     audio_scale, text_scale, dsigma, ds, cfg);
  else
    misc::cfg_euler_kernel<__nv_bfloat16><<<misc::grid_for(n), 256, 0, st>>>(
        pr, reinterpret_cast<const __nv_bfloat16*>(latents), o, no, n, audio_scale, text_scale, dsigma, ds, cfg);
  SA_LAUNCH_CHECK("cfg_euler_kernel launch");
  return SA_OK;
}

extern "C" int sa_window_blend(const sa_window_blend_args* a, sa_stream_t stream) {
  using namespace sa;
  if (!a || !a->new_latents || !a->pred_latents || a->n_windows <= 0 || a->n_windows > misc::BLEND_MAX_WINDOWS || a->C <= 0 ||
      a->N <= 0 || a->HW <= 0 || a->f_max <= 0 || a->overlap < 0 || a->overlap > misc::BLEND_MAX_OVERLAP ||
      (a->pred_dtype != SA_BF16 && a->pred_dtype != SA_F32)) {
    set_error("sa_window_blend: bad argument (<= %d windows, overlap <= %d)", misc::BLEND_MAX_WINDOWS, misc::BLEND_MAX_OVERLAP);
    return SA_ERR_BAD_ARG;
  }
  misc::WindowBlendParams p;
  p.new_lat = reinterpret_cast<const __nv_bfloat16*>(a->new_latents);
  p.pred = a->pred_latents;
  p.W = a->n_windows; p.C = a->C; p.N = a->N; p.HW = a->HW; p.f_max = a->f_max; p.overlap = a->overlap;
  p.pred_f32 = a->pred_dtype == SA_F32;
  for (int k = 0; k < p.W; ++k) {
    if (a->frames[k] <= 0 || a->frames[k] > a->f_max) { set_error("sa_window_blend: window %d has %d frames", k, a->frames[k]); return SA_ERR_BAD_ARG; }
    p.ws[k] = a->start[k]; p.f[k] = a->frames[k]; p.prev_end[k] = a->prev_end[k]; p.blend[k] = a->blend[k];
  }
  for (int j = 0; j < p.overlap; ++j) { p.w[j] = a->weight[j]; p.omw[j] = a->one_minus_weight[j]; }
  const long long total = (long long)p.C * p.HW;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (p.pred_f32) misc::window_blend_kernel<float><<<misc::grid_for(total), 256, 0, st>>>(p);
  else misc::window_blend_kernel<__nv_bfloat16><<<misc::grid_for(total), 256, 0, st>>>(p);
  SA_LAUNCH_CHECK("window_blend_kernel launch");
  return SA_OK;
}
