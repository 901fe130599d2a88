// fp32_kernels.cu — the fp32 mode of the DiT (BASELINE config 1: the reference's fp32 path, per-block tolerance 1e-4).
//
// Tensor cores have no fp32 operand type, so a fp32 Linear is run as ONE bf16 tcgen05 GEMM over a six-fold K:
// x = x0 + x1 + x2 and w = w0 + w1 + w2 are split into three bf16 terms each (24 mantissa bits) and
//     x w^T ~= x0 w0 + x0 w1 + x1 w0 + x1 w1 + x0 w2 + x2 w0        (dropped terms are < 2^-24 relative)
// is the K-concatenation [x0|x0|x1|x1|x0|x2] . [w0|w1|w0|w1|w2|w0]^T accumulated in the fp32 TMEM accumulator of the
// ordinary GEMM kernel (sa_gemm_bf16 with fp32 output, round_y = 0). sa_f32_split3 builds either operand. The other
// kernels are the fp32 versions of the elementwise steps whose bf16 versions are fused elsewhere; fp32 mode is a parity
// mode, not a throughput mode. References: 1B.py:296-342 (RMSNorm, RoPE), 1B.py:675-691 (modulation, gated residual),
// 1B.py:158-207 (softmax of the attention), 1B.py:972-983 / 1161-1184 (patchify / unpatchify), pipe.py:752-754 (CFG + Euler).
#include "../../include/stableavatar_b200.h"
#include "sa_host.h"
#include "sa_ptx.cuh"

namespace sa {
namespace f32k {

static inline int grid_for(long long total, int block = 256) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// out[m, s*K + k] = term[sel(s)](x[m, k]); pattern 0 (activation side): 0,0,1,1,0,2; pattern 1 (weight side): 0,1,0,1,2,0.
// K_pad >= K rounds the GEMM's K up to its 8-element granularity; the padding columns are zero.
__global__ void split3_kernel(const float* x, long long ld, long long M, int Kx, int K, __nv_bfloat16* out, int pattern) {
  const long long total = M * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / K;
    const int k = (int)(i % K);
    const float v = k < Kx ? x[m * ld + k] : 0.f;
    const __nv_bfloat16 t0 = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(t0);
    const __nv_bfloat16 t1 = __float2bfloat16_rn(r1);
    const __nv_bfloat16 t2 = __float2bfloat16_rn(r1 - __bfloat162float(t1));
    __nv_bfloat16* o = out + m * 6LL * K + k;
    if (pattern == 0) {
      o[0] = t0; o[K] = t0; o[2LL * K] = t1; o[3LL * K] = t1; o[4LL * K] = t0; o[5LL * K] = t2;
    } else {
      o[0] = t0; o[K] = t1; o[2LL * K] = t0; o[3LL * K] = t1; o[4LL * K] = t2; o[5LL * K] = t0;
    }
  }
}

struct PatchParams {
  const float* x; const float* y; float* out;
  int B, Cx, Cy, F, H, W, seq_len, K_pad;
};
__global__ void patchify_kernel(const PatchParams p) {  // fp32 twin of misc::patchify_kernel
  const int Hp = p.H / 2, Wp = p.W / 2;
  const int C = p.Cx + p.Cy;
  const long long total = (long long)p.B * p.seq_len * p.K_pad;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int k = idx % p.K_pad;
    const long long rt = idx / p.K_pad;
    const int tok = rt % p.seq_len, b = rt / p.seq_len;
    float v = 0.f;
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

__global__ void unpatchify_kernel(const float* u, float* out, long long u_bs, long long u_ls, int B, int Cout, int F, int H, int W) {
  const int Hp = H / 2, Wp = W / 2;
  const long long total = (long long)B * Cout * F * H * W;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int x = idx % W;
    long long r = idx / W;
    const int y = r % H; r /= H;
    const int f = r % F; r /= F;
    const int c = r % Cout;
    const int b = r / Cout;
    const int tok = (f * Hp + (y >> 1)) * Wp + (x >> 1);
    const int j = ((y & 1) * 2 + (x & 1)) * Cout + c;
    out[idx] = u[(long long)b * u_bs + (long long)tok * u_ls + j];
  }
}

// out = x * (1 + scale[b]) + shift[b]   (1B.py:675-676, 688-689, 721)
__global__ void modulate_kernel(const float* x, const float* shift, const float* scale, float* out, long long rows, int C,
                                int rows_per_batch, long long mod_bs) {
  const long long total = rows * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i % C);
    const long long mb = (r / rows_per_batch) * mod_bs + c;
    out[i] = x[i] * (1.0f + scale[mb]) + shift[mb];
  }
}

// h[r, :] += y[r, :] * gate[b, :] (gate == NULL: h += y)   (1B.py:679, 684, 691; vp1B.py:345-362)
__global__ void gated_add_kernel(float* h, const float* y, long long ldy, const float* gate, long long rows, int C,
                                 int rows_per_batch, long long gate_bs) {
  const long long total = rows * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i % C);
    const float v = y[r * ldy + c];
    h[i] += gate ? v * gate[(r / rows_per_batch) * gate_bs + c] : v;
  }
}

// One warp per row, in place: x = x * rsqrt(mean(x^2) + eps) * w, then the 3-D RoPE of 1B.py:296-323 (fp32 rotation from
// the fp32 (cos, sin) table; the reference rotates in complex128 and casts back to fp32).
__global__ void rmsnorm_rope_kernel(float* x, long long ld, const float* w, const float2* freqs, int rows, int C,
                                    int rows_per_batch, int F, int H, int W, float eps) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float* xr = x + (long long)row * ld;
  float ss = 0.f;
  for (int c = lane; c < C; c += 32) ss += xr[c] * xr[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rinv = rsqrtf(ss / C + eps);
  const int tok = row % rows_per_batch;
  const bool rotate = freqs != nullptr && tok < F * H * W;
  const int pf = tok / (H * W), ph = (tok / W) % H, pw = tok % W;
  for (int pr = lane; pr < C / 2; pr += 32) {       // one adjacent pair per iteration
    const int c = 2 * pr;
    float a = xr[c] * rinv * w[c], b = xr[c + 1] * rinv * w[c + 1];
    if (rotate) {
      const int j = (c & 127) >> 1;
      const int pos = j < 22 ? pf : (j < 43 ? ph : pw);
      const float2 cs = freqs[pos * 64 + j];
      const float ra = a * cs.x - b * cs.y, rb = a * cs.y + b * cs.x;
      a = ra; b = rb;
    }
    xr[c] = a; xr[c + 1] = b;
  }
}

// x[r, :] = softmax(x[r, :] * scale), in place; one block per row.
__global__ void softmax_rows_kernel(float* x, int n, long long ld, float scale) {
  float* row = x + (long long)blockIdx.x * ld;
  __shared__ float red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < n; j += 256) mx = fmaxf(mx, row[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int j = threadIdx.x; j < n; j += 256) {
    const float e = expf((row[j] - mx) * scale);
    row[j] = e;
    sum += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += red[i];
  const float inv = 1.0f / sum;
  for (int j = threadIdx.x; j < n; j += 256) row[j] *= inv;
}

// out[i, j, :] = a[i, :] + b[j, :]
__global__ void add_bcast_kernel(const float* a, const float* b, float* out, int na, int nb, int n) {
  const long long total = (long long)na * nb * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int col = idx % n;
    const long long r = idx / n;
    out[idx] = a[(r / nb) * n + col] + b[(r % nb) * n + col];
  }
}

// noise = u + a (d - u) + t (c - d); latents += dsigma * noise   (pipe.py:752-754 in fp32)
__global__ void cfg_euler_kernel(const float* pred, const float* lat, float* out, float* noise_out, long long n, float audio_scale,
                                 float text_scale, float dsigma, const float* dsigma_dev, int cfg) {
  if (dsigma_dev) dsigma = *dsigma_dev;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float np;
    if (cfg) {
      const float u = pred[i], d = pred[n + i], c = pred[2 * n + i];
      np = __fadd_rn(__fadd_rn(u, __fmul_rn(audio_scale, d - u)), __fmul_rn(text_scale, c - d));
    } else {
      np = pred[i];
    }
    if (noise_out) noise_out[i] = np;
    out[i] = __fadd_rn(lat[i], __fmul_rn(dsigma, np));
  }
}

}  // namespace f32k
}  // namespace sa

#define SA_F32_CHECK(name)                                      \
  do {                                                          \
    cudaError_t e__ = cudaGetLastError();                       \
    if (e__ != cudaSuccess) return sa::cuda_fail(e__, name);    \
  } while (0)
#define SA_ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int sa_f32_split3(const void* x, int64_t ld, int64_t M, int32_t K, int32_t K_pad, void* out, int32_t pattern,
                             sa_stream_t stream) {
  using namespace sa;
  if (!x || !out || M <= 0 || K <= 0 || ld < K || K_pad < K || pattern < 0 || pattern > 1) { set_error("sa_f32_split3: bad argument"); return SA_ERR_BAD_ARG; }
  f32k::split3_kernel<<<f32k::grid_for(M * K_pad), 256, 0, SA_ST(stream)>>>(reinterpret_cast<const float*>(x), ld, M, K, K_pad,
                                                                         reinterpret_cast<__nv_bfloat16*>(out), pattern);
  SA_F32_CHECK("split3_kernel launch");
  return SA_OK;
}

extern "C" int sa_f32_patchify(const void* x, const void* y, void* out, int32_t B, int32_t Cx, int32_t Cy, int32_t F, int32_t H,
                               int32_t W, int32_t seq_len, int32_t K_pad, sa_stream_t stream) {
  using namespace sa;
  if (!x || !out || (Cy > 0 && !y) || B <= 0 || F <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1) || K_pad < (Cx + Cy) * 4 ||
      seq_len < F * (H / 2) * (W / 2)) {
    set_error("sa_f32_patchify: bad argument");
    return SA_ERR_BAD_ARG;
  }
  f32k::PatchParams p{reinterpret_cast<const float*>(x), reinterpret_cast<const float*>(y), reinterpret_cast<float*>(out),
                      B, Cx, Cy, F, H, W, seq_len, K_pad};
  f32k::patchify_kernel<<<f32k::grid_for((long long)B * seq_len * K_pad), 256, 0, SA_ST(stream)>>>(p);
  SA_F32_CHECK("f32 patchify_kernel launch");
  return SA_OK;
}

extern "C" int sa_f32_unpatchify(const void* u, void* out, int64_t u_bs, int64_t u_ls, int32_t B, int32_t Cout, int32_t F, int32_t H,
                                 int32_t W, sa_stream_t stream) {
  using namespace sa;
  if (!u || !out || B <= 0 || Cout <= 0 || F <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1)) { set_error("sa_f32_unpatchify: bad argument"); return SA_ERR_BAD_ARG; }
  f32k::unpatchify_kernel<<<f32k::grid_for((long long)B * Cout * F * H * W), 256, 0, SA_ST(stream)>>>(
      reinterpret_cast<const float*>(u), reinterpret_cast<float*>(out), u_bs, u_ls, B, Cout, F, H, W);
  SA_F32_CHECK("f32 unpatchify_kernel launch");
  return SA_OK;
}

extern "C" int sa_f32_modulate(const void* x, const void* shift, const void* scale, void* out, int64_t rows, int32_t C,
                               int32_t rows_per_batch, int64_t mod_bs, sa_stream_t stream) {
  using namespace sa;
  if (!x || !shift || !scale || !out || rows <= 0 || C <= 0 || rows_per_batch <= 0) { set_error("sa_f32_modulate: bad argument"); return SA_ERR_BAD_ARG; }
  f32k::modulate_kernel<<<f32k::grid_for(rows * C), 256, 0, SA_ST(stream)>>>(
      reinterpret_cast<const float*>(x), reinterpret_cast<const float*>(shift), reinterpret_cast<const float*>(scale),
      reinterpret_cast<float*>(out), rows, C, rows_per_batch, mod_bs);
  SA_F32_CHECK("modulate_kernel launch");
  return SA_OK;
}

extern "C" int sa_f32_gated_add(void* h, const void* y, int64_t ldy, const void* gate, int64_t rows, int32_t C, int32_t rows_per_batch,
                                int64_t gate_bs, sa_stream_t stream) {
  using namespace sa;
  if (!h || !y || rows <= 0 || C <= 0 || ldy < C || (gate && rows_per_batch <= 0)) { set_error("sa_f32_gated_add: bad argument"); return SA_ERR_BAD_ARG; }
  f32k::gated_add_kernel<<<f32k::grid_for(rows * C), 256, 0, SA_ST(stream)>>>(
      reinterpret_cast<float*>(h), reinterpret_cast<const float*>(y), ldy, reinterpret_cast<const float*>(gate), rows, C,
      rows_per_batch > 0 ? rows_per_batch : 1, gate_bs);
  SA_F32_CHECK("gated_add_kernel launch");
  return SA_OK;
}

extern "C" int sa_f32_rmsnorm_rope(void* x, int64_t ld, const void* weight, const void* freqs, int32_t rows, int32_t C,
                                   int32_t rows_per_batch, int32_t F, int32_t H, int32_t W, float eps, sa_stream_t stream) {
  using namespace sa;
  if (!x || !weight || rows <= 0 || C <= 0 || (C & 1) || ld < C || (freqs && (C % 128 || F <= 0 || H <= 0 || W <= 0 || F > 1024 || H > 1024 || W > 1024))) {
    set_error("sa_f32_rmsnorm_rope: bad argument");
    return SA_ERR_BAD_ARG;
  }
  f32k::rmsnorm_rope_kernel<<<(rows + 7) / 8, 256, 0, SA_ST(stream)>>>(
      reinterpret_cast<float*>(x), ld, reinterpret_cast<const float*>(weight), reinterpret_cast<const float2*>(freqs), rows, C,
      rows_per_batch > 0 ? rows_per_batch : rows, F, H, W, eps);
  SA_F32_CHECK("f32 rmsnorm_rope_kernel launch");
  return SA_OK;
}

extern "C" int sa_f32_softmax_rows(void* x, int32_t rows, int32_t n, int64_t ld, float scale, sa_stream_t stream) {
  using namespace sa;
  if (!x || rows <= 0 || n <= 0 || ld < n) { set_error("sa_f32_softmax_rows: bad argument"); return SA_ERR_BAD_ARG; }
  f32k::softmax_rows_kernel<<<rows, 256, 0, SA_ST(stream)>>>(reinterpret_cast<float*>(x), n, ld, scale);
  SA_F32_CHECK("f32 softmax_rows_kernel launch");
  return SA_OK;
}

extern "C" int sa_f32_add_bcast(const void* a, const void* b, void* out, int32_t na, int32_t nb, int32_t n, sa_stream_t stream) {
  using namespace sa;
  if (!a || !b || !out || na <= 0 || nb <= 0 || n <= 0) { set_error("sa_f32_add_bcast: bad argument"); return SA_ERR_BAD_ARG; }
  f32k::add_bcast_kernel<<<f32k::grid_for((long long)na * nb * n), 256, 0, SA_ST(stream)>>>(
      reinterpret_cast<const float*>(a), reinterpret_cast<const float*>(b), reinterpret_cast<float*>(out), na, nb, n);
  SA_F32_CHECK("f32 add_bcast_kernel launch");
  return SA_OK;
}

extern "C" int sa_f32_cfg_euler_step(const void* pred, const void* latents, void* out, void* noise_out, int64_t n, float audio_scale,
                                     float text_scale, float dsigma, const void* dsigma_dev, int32_t cfg, sa_stream_t stream) {
  using namespace sa;
  if (!pred || !latents || !out || n <= 0) { set_error("sa_f32_cfg_euler_step: bad argument"); return SA_ERR_BAD_ARG; }
  f32k::cfg_euler_kernel<<<f32k::grid_for(n), 256, 0, SA_ST(stream)>>>(
      reinterpret_cast<const float*>(pred), reinterpret_cast<const float*>(latents), reinterpret_cast<float*>(out),
      reinterpret_cast<float*>(noise_out), n, audio_scale, text_scale, dsigma, reinterpret_cast<const float*>(dsigma_dev), cfg);
  SA_F32_CHECK("f32 cfg_euler_kernel launch");
  return SA_OK;
}
