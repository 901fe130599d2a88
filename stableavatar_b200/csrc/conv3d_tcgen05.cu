// conv3d_tcgen05.cu — causal 3-D convolution of the Wan VAE decoder as an implicit GEMM on tcgen05 tensor cores.
//
//   out[t, h, w, n] = bias[n] + sum_{kt,kh,kw,c} in[t*stride_t + kt, h + kh - pad_h, w + kw - pad_w, c] * Wt[n, (kt,kh,kw,c)]
//
// (pad = K/2 'same' for the decoder; the encoder's stride-2 3x3 Conv2d runs as a 2x2-tap conv with pad 0 over a
// space-to-depth input, and its (3,1,1) stride-2 time conv uses stride_t = 2.)
//
// Activations are channels-last bf16 [T, H, W, C]; `in` already carries the KT-1 causal cache frames in front
// (ring buffer kept by the host, replacing the reference's per-chunk clone + cat: wan/models/wan_vae.py:31-39,
// 208-220), so temporal causality is a plain offset and the spatial 'same' padding is TMA out-of-bounds zero fill —
// no im2col and no padded copy is ever materialised.
//
// GEMM view: M = output positions (tile = 16 w x 8 h of one frame = 128 rows), N = Cout (tile bn <= 192),
// K = taps x Cin walked as (tap, 32-channel sub-tiles). Per pipeline stage the producer issues, for up to three
// 32-channel sub-tiles, one 4-D TMA box {32 c, 16 w, 8 h, 1 t} of the input shifted by the tap offset and one 2-D box
// {32 k, bn} of the weights, both SWIZZLE_64B; the MMA warp issues 2 x tcgen05.mma (K = 16) per sub-tile into a
// double-buffered TMEM accumulator; four epilogue warps add bias (+ residual), and store bf16 channels-last, the
// frame-interleaved split of the temporal upsampler (wan_vae.py:137-140), or the clamped fp32 planar video.
//
// Replaces CausalConv3d / Conv2d of wan/models/wan_vae.py (Decoder3d :372-475, Resample :69-143, ResidualBlock :189-223).
#include "../../include/stableavatar_b200.h"
#include "sa_host.h"
#include "sa_ptx.cuh"

namespace sa {
namespace conv {

constexpr int TW = 16, TH = 8, BM = TW * TH;  // 128 output positions per tile
constexpr int SUBK = 32;                      // channels per sub-tile (64-byte rows, SWIZZLE_64B)
constexpr int A_SUB_BYTES = BM * SUBK * 2;    // 8 KB
constexpr int MAX_BN = 192;
constexpr int MAX_KSUB = 3;
constexpr int MAX_STAGES = 6;
constexpr int NUM_THREADS = 192;              // TMA, MMA, 4 epilogue warps
constexpr int TMEM_COLS = 512;
constexpr int SMEM_DATA = 200 * 1024;
constexpr int SMEM_BYTES = SMEM_DATA + 256 + 1024;

struct Params {
  const float* bias;
  const __nv_bfloat16* res;
  void* out;
  int Tout, H, W, Cin, Cout, KT, KH, KW;
  int bn, ksub, groups_per_tap, stages, stage_bytes;
  int tiles_w, tiles_h, tiles_n;
  int out_mode, out_T_total, out_t0;
  int pad_h, pad_w, stride_t;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv3d_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_w, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SMEM_DATA);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* tfull = empty + MAX_STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int taps = p.KT * p.KH * p.KW;
  const int num_ks = taps * p.groups_per_tap;  // pipeline steps per tile
  const int tiles_per_frame = p.tiles_w * p.tiles_h;
  const int num_tiles = p.tiles_n * p.Tout * tiles_per_frame;
  const int b_sub_bytes = p.bn * SUBK * 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_in);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tfull[s], 1);
        mbar_init(&tempty[s], 4);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> (n tile, frame, h tile, w tile); w fastest so that neighbouring CTAs share input halos in L2
  auto decode = [&](int tile, int& nt, int& t, int& h0, int& w0) {
    nt = tile / (p.Tout * tiles_per_frame);
    int r = tile % (p.Tout * tiles_per_frame);
    t = r / tiles_per_frame;
    r %= tiles_per_frame;
    h0 = (r / p.tiles_w) * TH;
    w0 = (r % p.tiles_w) * TW;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {  // elect.sync: single active lane is known to ptxas -> no R2UR waterfall per UTCHMMA / UTMALDG
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int nt, t, h0, w0;
        decode(tile, nt, t, h0, w0);
        for (int ks = 0; ks < num_ks; ++ks, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          const int tap = ks / p.groups_per_tap, grp = ks % p.groups_per_tap;
          const int kw = tap % p.KW, kh = (tap / p.KW) % p.KH, kt = tap / (p.KW * p.KH);
          mbar_wait(&empty[s], ph ^ 1, 0x2100 | s);
          mbar_arrive_expect_tx(&full[s], p.stage_bytes);
          uint8_t* st = smem + s * p.stage_bytes;
          for (int j = 0; j < p.ksub; ++j) {
            const int c0 = (grp * p.ksub + j) * SUBK;
            tma_load_4d(st + j * A_SUB_BYTES, &tmap_in, &full[s], c0, w0 + kw - p.pad_w, h0 + kh - p.pad_h, t * p.stride_t + kt);
            tma_load_2d(st + p.ksub * A_SUB_BYTES + j * b_sub_bytes, &tmap_w, &full[s], tap * p.Cin + c0, nt * p.bn);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {  // elect.sync: single active lane is known to ptxas -> no R2UR waterfall per UTCHMMA / UTMALDG
      const uint32_t idesc = umma_idesc_bf16(BM, p.bn, 0, 0);
      uint32_t it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t as = tcount & 1;
        mbar_wait(&tempty[as], ((tcount >> 1) & 1) ^ 1, 0x2200 | as);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
        for (int ks = 0; ks < num_ks; ++ks, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(&full[s], ph, 0x2300 | s);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * p.stage_bytes);
          const uint32_t b_addr = a_addr + p.ksub * A_SUB_BYTES;
          for (int j = 0; j < p.ksub; ++j) {
            const uint64_t adesc = umma_smem_desc(a_addr + j * A_SUB_BYTES, 16, 512, kSwz64);
            const uint64_t bdesc = umma_smem_desc(b_addr + j * b_sub_bytes, 16, 512, kSwz64);
            umma_ss(d_tmem, adesc, bdesc, idesc, (ks | j) != 0);
            umma_ss(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);  // +32 bytes = second K=16 half of the 64-byte row
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&tfull[as]);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: warps 2..5, one output position per thread
    const int q = warp & 3;
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
      int nt, t, h0, w0;
      decode(tile, nt, t, h0, w0);
      const uint32_t as = tcount & 1;
      mbar_wait(&tfull[as], (tcount >> 1) & 1, 0x2400 | as);
      tc_fence_after();
      const int r = q * 32 + lane;
      const int h = h0 + r / TW, w = w0 + r % TW;
      const bool ok = h < p.H && w < p.W;
      const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + as * 256;
      const long long pos = ((long long)t * p.H + h) * p.W + w;
#pragma unroll 1
      for (int c = 0; c < p.bn / 16; ++c) {
        uint32_t rr[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]),
              "=r"(rr[8]), "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]), "=r"(rr[14]), "=r"(rr[15])
            : "r"(t_row + c * 16)
            : "memory");
        tmem_ld_wait();
        const int n0 = nt * p.bn + c * 16;
        if (!ok || n0 >= p.Cout) continue;
        float y[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) y[i] = __uint_as_float(rr[i]) + (n0 + i < p.Cout ? __ldg(p.bias + n0 + i) : 0.f);
        if (p.out_mode == 3) {  // fp32 channels-last [Tout, H, W, Cout], unclamped (encoder head, wan_vae.py:320-322)
          float* o = reinterpret_cast<float*>(p.out) + pos * p.Cout + n0;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (n0 + i < p.Cout) o[i] = y[i];
        } else if (p.out_mode == 2) {  // fp32 planar [Cout, T_total, H, W], clamped to [-1, 1] (wan_vae.py:668)
          float* o = reinterpret_cast<float*>(p.out);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (n0 + i < p.Cout)
              o[(((long long)(n0 + i) * p.out_T_total + p.out_t0 + t) * p.H + h) * p.W + w] = fminf(fmaxf(y[i], -1.f), 1.f);
        } else {
          long long off;
          if (p.out_mode == 1) {  // channels [0,C) -> frame 2t, [C,2C) -> frame 2t+1 (wan_vae.py:137-140)
            const int C = p.Cout >> 1;
            off = ((((long long)(2 * t + n0 / C)) * p.H + h) * p.W + w) * C + n0 % C;
          } else {
            off = pos * p.Cout + n0;
          }
          if (p.res) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.res + off);
            const uint4 u0 = rp[0], u1 = rp[1];
            const __nv_bfloat162* hh0 = reinterpret_cast<const __nv_bfloat162*>(&u0);
            const __nv_bfloat162* hh1 = reinterpret_cast<const __nv_bfloat162*>(&u1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f0 = __bfloat1622float2(hh0[i]), f1 = __bfloat1622float2(hh1[i]);
              y[2 * i] += f0.x; y[2 * i + 1] += f0.y;
              y[8 + 2 * i] += f1.x; y[8 + 2 * i + 1] += f1.y;
            }
          }
          uint4 v0, v1;
          v0.x = pack_bf16x2(y[0], y[1]);   v0.y = pack_bf16x2(y[2], y[3]);
          v0.z = pack_bf16x2(y[4], y[5]);   v0.w = pack_bf16x2(y[6], y[7]);
          v1.x = pack_bf16x2(y[8], y[9]);   v1.y = pack_bf16x2(y[10], y[11]);
          v1.z = pack_bf16x2(y[12], y[13]); v1.w = pack_bf16x2(y[14], y[15]);
          uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off);
          op[0] = v0;
          op[1] = v1;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace conv
}  // namespace sa

extern "C" int sa_conv3d_cl(const sa_conv_args* a, sa_stream_t stream_) {
  using namespace sa;
  using namespace sa::conv;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !a->in || !a->w || !a->bias || !a->out) { set_error("sa_conv3d_cl: null pointer"); return SA_ERR_BAD_ARG; }
  if (a->Tout <= 0 || a->H <= 0 || a->W <= 0 || a->Cin <= 0 || a->Cout <= 0 || a->KT <= 0 || a->KH <= 0 || a->KW <= 0 ||
      a->pad_h >= a->KH || a->pad_w >= a->KW || a->stride_t < 0) {
    set_error("sa_conv3d_cl: bad dims / padding / stride");
    return SA_ERR_BAD_ARG;
  }
  if (a->Cin % SUBK) { set_error("sa_conv3d_cl: Cin must be a multiple of 32 (pad on the host), got %d", a->Cin); return SA_ERR_BAD_ARG; }
  const int cout_pad = (a->Cout + 15) / 16 * 16;
  int bn;
  if (cout_pad % 192 == 0) bn = 192;
  else if (cout_pad <= MAX_BN) bn = cout_pad;
  else if (cout_pad % 128 == 0) bn = 128;
  else { set_error("sa_conv3d_cl: unsupported Cout %d", a->Cout); return SA_ERR_UNSUPPORTED; }
  if (a->out_mode < 0 || a->out_mode > 3 || (a->out_mode < 2 && a->Cout % 16) || (a->out_mode == 1 && (a->Cout / 2) % 16) ||
      (a->res && a->out_mode >= 2)) {
    set_error("sa_conv3d_cl: bad out_mode / Cout combination");
    return SA_ERR_BAD_ARG;
  }
  const int nsub = a->Cin / SUBK;
  const int ksub = nsub % 3 == 0 ? 3 : (nsub % 2 == 0 ? 2 : 1);
  Params p;
  p.bias = reinterpret_cast<const float*>(a->bias);
  p.res = reinterpret_cast<const __nv_bfloat16*>(a->res);
  p.out = a->out;
  p.Tout = a->Tout; p.H = a->H; p.W = a->W; p.Cin = a->Cin; p.Cout = a->Cout; p.KT = a->KT; p.KH = a->KH; p.KW = a->KW;
  p.bn = bn; p.ksub = ksub; p.groups_per_tap = nsub / ksub;
  p.stage_bytes = ksub * (A_SUB_BYTES + bn * SUBK * 2);
  p.stages = SMEM_DATA / p.stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  p.tiles_w = (a->W + TW - 1) / TW; p.tiles_h = (a->H + TH - 1) / TH; p.tiles_n = cout_pad / bn;
  p.out_mode = a->out_mode; p.out_T_total = a->out_T_total; p.out_t0 = a->out_t0;
  p.pad_h = a->pad_h < 0 ? a->KH / 2 : a->pad_h;
  p.pad_w = a->pad_w < 0 ? a->KW / 2 : a->pad_w;
  p.stride_t = a->stride_t > 0 ? a->stride_t : 1;

  CUtensorMap tin, tw;
  {
    const int Tin = (a->Tout - 1) * p.stride_t + a->KT;
    uint64_t dims[4] = {(uint64_t)a->Cin, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)Tin};
    uint64_t strides[3] = {(uint64_t)a->Cin * 2, (uint64_t)a->W * a->Cin * 2, (uint64_t)a->H * a->W * a->Cin * 2};
    uint32_t box[4] = {SUBK, TW, TH, 1};
    int rc = make_tmap_bf16(&tin, a->in, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }
  {
    const uint64_t Ktot = (uint64_t)a->KT * a->KH * a->KW * a->Cin;
    uint64_t dims[2] = {Ktot, (uint64_t)cout_pad};
    uint64_t strides[1] = {Ktot * 2};
    uint32_t box[2] = {SUBK, (uint32_t)bn};
    int rc = make_tmap_bf16(&tw, a->w, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }
  if (int rc = ensure_dyn_smem(conv3d_kernel, SMEM_BYTES, "conv3d_kernel")) return rc;
  const int tiles = p.tiles_n * p.Tout * p.tiles_w * p.tiles_h;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  conv3d_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tin, tw, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "conv3d_kernel launch");
  return SA_OK;
}
