// vae_kernels.cu — the HBM-bound kernels of the Wan VAE decoder, all on channels-last bf16 activations [P, C]:
// channel RMS-norm (+SiLU) (wan_vae.py:54-57 + nn.SiLU), nearest-exact 2x upsample (:60-66, 79-88), row softmax of the
// single-head middle attention (:243-265), and the latent de-normalisation + 1x1x1 conv2 (:552-559). Encode side:
// space-to-depth in front of the stride-2 Conv2d (:96-104), video-in layout change, 1x1x1 conv1 + latent normalisation
// (:539-545).
#include "../../include/stableavatar_b200.h"
#include "sa_host.h"
#include "sa_ptx.cuh"

namespace sa {
namespace vae {

// ------------------------------------------------------------------------------------------------ RMS-norm (+SiLU)
// y = x / max(||x||_2, 1e-12) * sqrt(C) * gamma ; optionally y * sigmoid(y). LPR lanes cooperate on one position.
template <int LPR, int NCH>
__global__ void __launch_bounds__(256) rmsnorm_silu_kernel(const __nv_bfloat16* x, const float* gamma, __nv_bfloat16* out,
                                                            long long P, int C, int silu) {
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const long long row = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * (32 / LPR) + lane / LPR;
  const bool row_ok = row < P;
  const int nchunks = C >> 3;
  float v[NCH][8];
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int ch = sub + c * LPR;
    if (row_ok && ch < nchunks) {
      const uint4 u = *reinterpret_cast<const uint4*>(x + row * C + ch * 8);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        v[c][2 * i] = f.x;
        v[c][2 * i + 1] = f.y;
        ss += f.x * f.x + f.y * f.y;
      }
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = sqrtf((float)C) / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int ch = sub + c * LPR;
    if (!(row_ok && ch < nchunks)) continue;
    float y[8];
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + ch * 8), g1 = *reinterpret_cast<const float4*>(gamma + ch * 8 + 4);
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float t = v[c][i] * inv * g[i];
      if (silu) t = __fdividef(t, 1.0f + __expf(-t));
      y[i] = t;
    }
    uint4 u;
    u.x = pack_bf16x2(y[0], y[1]);
    u.y = pack_bf16x2(y[2], y[3]);
    u.z = pack_bf16x2(y[4], y[5]);
    u.w = pack_bf16x2(y[6], y[7]);
    *reinterpret_cast<uint4*>(out + row * C + ch * 8) = u;
  }
}

// ------------------------------------------------------------------------------------------------ nearest 2x upsample
// out[t, y, x, :] = in[t, y/2, x/2, :]   (nearest-exact with scale 2 == floor((i + 0.5) / 2) == i / 2)
__global__ void upsample2x_kernel(const uint4* in, uint4* out, int T, int H, int W, int C8) {
  const long long total = (long long)T * 2 * H * 2 * W * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = i % C8;
    long long r = i / C8;
    const int x = r % (2 * W); r /= 2 * W;
    const int y = r % (2 * H);
    const int t = r / (2 * H);
    out[i] = in[(((long long)t * H + (y >> 1)) * W + (x >> 1)) * C8 + c];
  }
}

// ------------------------------------------------------------------------------------------------ row softmax
// out[r, :] = bf16(softmax(in[r, :] * scale)), in fp32 [R, n] (row stride ld_in), one CTA per row.
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* in, __nv_bfloat16* out, int n, long long ld_in,
                                                            long long ld_out, float scale_log2) {
  __shared__ float red[8];
  const float* row = in + (long long)blockIdx.x * ld_in;
  __nv_bfloat16* orow = out + (long long)blockIdx.x * ld_out;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
  for (int j = threadIdx.x; j < n; j += 256) sum += exp2f((row[j] - mx) * scale_log2);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += red[i];
  const float inv = 1.0f / sum;
  for (int j = threadIdx.x; j < n; j += 256) orow[j] = __float2bfloat16_rn(exp2f((row[j] - mx) * scale_log2) * inv);
}

// ------------------------------------------------------------------------------------------------ latent input
// x[t, h, w, co] = b[co] + sum_ci Wc[co, ci] * (z[ci, t, h, w] * std[ci] + mean[ci]), channels-last bf16 padded to Cpad.
__global__ void latent_in_kernel(const float* z, const float* wc, const float* bc, const float* mean, const float* stdv,
                                 __nv_bfloat16* out, int Cz, long long P, int Cpad) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < P * Cpad; i += (long long)gridDim.x * blockDim.x) {
    const int co = i % Cpad;
    const long long pos = i / Cpad;
    float acc = 0.f;
    if (co < Cz) {
      acc = bc[co];
      for (int ci = 0; ci < Cz; ++ci) acc = fmaf(wc[co * Cz + ci], z[(long long)ci * P + pos] * stdv[ci] + mean[ci], acc);
    }
    out[i] = __float2bfloat16_rn(acc);
  }
}

// ------------------------------------------------------------------------------------------------ encode helpers
// out[t, y, x, (dy*2+dx)*C + c] = in[t, 2y+dy, 2x+dx, c]; one 16-byte chunk (8 channels) per thread, writes coalesced.
__global__ void space_to_depth_kernel(const uint4* in, uint4* out, int T, int H, int W, int C8) {
  const int Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)T * Ho * Wo * 4 * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = i % C8;
    long long r = i / C8;
    const int d = r % 4;
    r /= 4;
    const int x = r % Wo;
    r /= Wo;
    const int y = r % Ho;
    const int t = r / Ho;
    out[i] = in[(((long long)t * H + 2 * y + (d >> 1)) * W + 2 * x + (d & 1)) * C8 + c];
  }
}

// x f32 planar [Cx, P] -> bf16 channels-last [P, Cpad] (zero padded channels).
__global__ void video_in_kernel(const float* x, __nv_bfloat16* out, int Cx, long long P, int Cpad) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < P * Cpad; i += (long long)gridDim.x * blockDim.x) {
    const int c = i % Cpad;
    const long long pos = i / Cpad;
    out[i] = __float2bfloat16_rn(c < Cx ? x[(long long)c * P + pos] : 0.f);
  }
}

// out[co, pos] = b[co] + sum_ci Wc[co, ci] * h[pos, ci]; co < Cz additionally (. - mean[co]) / std[co].
__global__ void latent_out_kernel(const float* h, const float* wc, const float* bc, const float* mean, const float* stdv,
                                  float* out, int Cz, long long P) {
  const int C2 = 2 * Cz;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < P * C2; i += (long long)gridDim.x * blockDim.x) {
    const long long pos = i % P;
    const int co = i / P;
    float acc = bc[co];
    for (int ci = 0; ci < C2; ++ci) acc = fmaf(wc[co * C2 + ci], h[pos * C2 + ci], acc);
    if (co < Cz) acc = (acc - mean[co]) * (1.0f / stdv[co]);
    out[i] = acc;
  }
}

static inline int grid_for(long long total, int block = 256) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace vae
}  // namespace sa

#define SA_LAUNCH_CHECK(name)                                   \
  do {                                                          \
    cudaError_t e__ = cudaGetLastError();                       \
    if (e__ != cudaSuccess) return sa::cuda_fail(e__, name);    \
  } while (0)

extern "C" int sa_vae_rmsnorm_silu(const void* x, const void* gamma, void* out, int64_t P, int32_t C, int32_t silu,
                                   sa_stream_t stream_) {
  using namespace sa;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!x || !gamma || !out || P <= 0 || C <= 0 || C % 8 || C > 512) {
    set_error("sa_vae_rmsnorm_silu: bad argument (C %% 8 == 0, C <= 512)");
    return SA_ERR_BAD_ARG;
  }
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  const float* gp = reinterpret_cast<const float*>(gamma);
  __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(out);
  if (C == 96 || C == 192 || C == 384) {
    // the decoder's widths: 3 chunks per lane, every lane busy (C / 24 lanes per position), three 16-byte loads in flight per
    // thread — the one-chunk-per-lane form below keeps 4 of 16 lanes idle at C = 96 and reached 57 % of the HBM copy rate
    const int lpr = C / 24;
    const long long rpb = 8 * (32 / lpr);
    const unsigned grid = (unsigned)((P + rpb - 1) / rpb);
    if (lpr == 4) vae::rmsnorm_silu_kernel<4, 3><<<grid, 256, 0, stream>>>(xp, gp, op, P, C, silu);
    else if (lpr == 8) vae::rmsnorm_silu_kernel<8, 3><<<grid, 256, 0, stream>>>(xp, gp, op, P, C, silu);
    else vae::rmsnorm_silu_kernel<16, 3><<<grid, 256, 0, stream>>>(xp, gp, op, P, C, silu);
  } else if (C <= 128) {
    const long long rows_per_block = 8 * 2;
    vae::rmsnorm_silu_kernel<16, 1><<<(unsigned)((P + rows_per_block - 1) / rows_per_block), 256, 0, stream>>>(xp, gp, op, P, C, silu);
  } else if (C <= 256) {
    vae::rmsnorm_silu_kernel<32, 1><<<(unsigned)((P + 7) / 8), 256, 0, stream>>>(xp, gp, op, P, C, silu);
  } else {
    vae::rmsnorm_silu_kernel<32, 2><<<(unsigned)((P + 7) / 8), 256, 0, stream>>>(xp, gp, op, P, C, silu);
  }
  SA_LAUNCH_CHECK("rmsnorm_silu_kernel launch");
  return SA_OK;
}

extern "C" int sa_vae_upsample2x(const void* in, void* out, int32_t T, int32_t H, int32_t W, int32_t C, sa_stream_t stream) {
  using namespace sa;
  if (!in || !out || T <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8) { set_error("sa_vae_upsample2x: bad argument"); return SA_ERR_BAD_ARG; }
  vae::upsample2x_kernel<<<vae::grid_for((long long)T * 4 * H * W * (C / 8)), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), T, H, W, C / 8);
  SA_LAUNCH_CHECK("upsample2x_kernel launch");
  return SA_OK;
}

extern "C" int sa_softmax_rows(const void* in, void* out, int32_t rows, int32_t n, int64_t ld_in, int64_t ld_out, float scale,
                               sa_stream_t stream) {
  using namespace sa;
  if (!in || !out || rows <= 0 || n <= 0) { set_error("sa_softmax_rows: bad argument"); return SA_ERR_BAD_ARG; }
  vae::softmax_rows_kernel<<<rows, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float*>(in), reinterpret_cast<__nv_bfloat16*>(out), n, ld_in, ld_out, scale * 1.4426950408889634f);
  SA_LAUNCH_CHECK("softmax_rows_kernel launch");
  return SA_OK;
}

extern "C" int sa_vae_latent_in(const void* z, const void* wc, const void* bc, const void* mean, const void* stdv, void* out,
                                int32_t Cz, int64_t P, int32_t Cpad, sa_stream_t stream) {
  using namespace sa;
  if (!z || !wc || !bc || !mean || !stdv || !out || Cz <= 0 || P <= 0 || Cpad < Cz) { set_error("sa_vae_latent_in: bad argument"); return SA_ERR_BAD_ARG; }
  vae::latent_in_kernel<<<vae::grid_for(P * Cpad), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float*>(z), reinterpret_cast<const float*>(wc), reinterpret_cast<const float*>(bc),
      reinterpret_cast<const float*>(mean), reinterpret_cast<const float*>(stdv), reinterpret_cast<__nv_bfloat16*>(out), Cz, P, Cpad);
  SA_LAUNCH_CHECK("latent_in_kernel launch");
  return SA_OK;
}

extern "C" int sa_vae_space_to_depth(const void* in, void* out, int32_t T, int32_t H, int32_t W, int32_t C, sa_stream_t stream) {
  using namespace sa;
  if (!in || !out || T <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8 || (H & 1) || (W & 1)) {
    set_error("sa_vae_space_to_depth: bad argument (C %% 8 == 0, H and W even)");
    return SA_ERR_BAD_ARG;
  }
  vae::space_to_depth_kernel<<<vae::grid_for((long long)T * H * W * (C / 8)), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), T, H, W, C / 8);
  SA_LAUNCH_CHECK("space_to_depth_kernel launch");
  return SA_OK;
}

extern "C" int sa_vae_video_in(const void* x, void* out, int32_t Cx, int64_t P, int32_t Cpad, sa_stream_t stream) {
  using namespace sa;
  if (!x || !out || Cx <= 0 || P <= 0 || Cpad < Cx) { set_error("sa_vae_video_in: bad argument"); return SA_ERR_BAD_ARG; }
  vae::video_in_kernel<<<vae::grid_for(P * Cpad), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float*>(x), reinterpret_cast<__nv_bfloat16*>(out), Cx, P, Cpad);
  SA_LAUNCH_CHECK("video_in_kernel launch");
  return SA_OK;
}

extern "C" int sa_vae_latent_out(const void* h, const void* wc, const void* bc, const void* mean, const void* stdv, void* out,
                                 int32_t Cz, int64_t P, sa_stream_t stream) {
  using namespace sa;
  if (!h || !wc || !bc || !mean || !stdv || !out || Cz <= 0 || P <= 0) { set_error("sa_vae_latent_out: bad argument"); return SA_ERR_BAD_ARG; }
  vae::latent_out_kernel<<<vae::grid_for(P * 2 * Cz), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float*>(h), reinterpret_cast<const float*>(wc), reinterpret_cast<const float*>(bc),
      reinterpret_cast<const float*>(mean), reinterpret_cast<const float*>(stdv), reinterpret_cast<float*>(out), Cz, P);
  SA_LAUNCH_CHECK("latent_out_kernel launch");
  return SA_OK;
}
