// conv3d_halo_tcgen05.cu — the (1|3)x3x3 'same' stride-1 convolutions of the Wan VAE with 96 / 192 / 384 output channels and
// the 96 -> 3 video head (95 % of the decoder's FLOPs: every ResidualBlock conv, the 2-D convs of the upsamplers, the head)
// as an implicit GEMM whose input tile is staged ONCE.
//
// conv3d_tcgen05.cu re-loads the [128 positions x Cin] tile for each of the 27 taps and the tap's weights for every tile:
// 147 bytes per SM-clock of L2 -> shared-memory traffic at full tensor rate, three times what the L2 delivers — it runs at
// 0.44 (Cout 96) / 0.70 (Cout 192) of the sustained tensor rate. Here, per time tap kt and channel group:
//
//   * the (16+2) x (8+2) halo of a 16 h x 8 w output tile is copied to shared memory once (cp.async, 16-byte chunks, zero
//     fill outside the frame = the 'same' padding) in the tcgen05 NO-SWIZZLE K-major layout: plane c = channels 8c..8c+7,
//     inside a plane position (hh, ww) at (hh * 10 + ww) * 16 bytes. The A operand of spatial tap (kh, kw) is then the
//     SAME buffer at a start offset of (kh * 10 + kw) * 16 bytes with SBO = 160 bytes (one halo row: the next 8 output
//     positions) and LBO = plane size — 9 taps, no copy, no im2col;
//   * a weight stage (one tap x channel group, pre-packed [Cout/BN][tap][Cin/8][BN][8] on the host = the same layout with
//     SBO 128, LBO BN * 16) arrives as one cp.async.bulk and is used by G output tiles whose accumulators sit side by side
//     in TMEM (G = 4 x 96 or 2 x 192 columns; Cout 384 = two passes over 192-channel tiles), so weights cross L2 -> SM
//     once per G tiles: 64 / G bytes per SM-clock;
//   * the epilogue pulls an accumulator into registers and hands it back before bias / residual / stores; with nine weight
//     stages resident (BN 96 / 16) the last halo stage of a pass is issued tile by tile, one completion barrier per tile.
//
// L2 -> SM traffic at full tensor rate: 29 (Cout 96) / 39 (Cout 192) bytes per SM-clock instead of 147 / 120. Measured
// (tools/conv_bench.py): 96 -> 96 @4x480x832 608 -> 1070 TFLOP/s, 192 -> 192 @4x240x416 980 -> 1381.
// Warps: 0 = weight producer, 1 = MMA issuer, 2-5 = epilogue (TMEM lane quarter = warp % 4), 6-9 = halo producers.
//
// Replaces CausalConv3d / Conv2d of wan/models/wan_vae.py (:20-39, 69-143, 189-223, 438-441) for those shapes; everything
// else (3- / 16- / 32-channel inputs, (3,1,1) time convs, strided encoder convs, fp32 encoder head, frames too small to fill
// the coarse work units) stays on conv3d_tcgen05.cu.
#include "../../include/stableavatar_b200.h"
#include "sa_host.h"
#include "sa_ptx.cuh"

namespace sa {
namespace convh {

constexpr int TW = 8, TH = 16, BM = TW * TH;
constexpr int HW_ = TW + 2, HH_ = TH + 2, HPOS = HW_ * HH_;   // 10 x 18 halo positions
constexpr int PLANE = HPOS * 16 + 16;                          // + 16: consecutive planes start in different banks
constexpr int HALO_BUFS = 2;
constexpr int MAX_W_STAGES = 9;               // 9 = the taps of one halo stage: lets the last stage of a pass keep all of them resident
constexpr int NUM_THREADS = 320;
constexpr int PRODUCERS = 128;
constexpr int TMEM_COLS = 512;
constexpr int SMEM_DATA = 225 * 1024;
constexpr int BIAS_BYTES = 2048;               // up to 512 output channels, the tail of the data region
constexpr int SMEM_BYTES = SMEM_DATA + 256 + 1024;

struct Params {
  const __nv_bfloat16* in;
  const __nv_bfloat16* w;
  const float* bias;
  const __nv_bfloat16* res;
  void* out;
  int Tout, H, W, Cin, Cout, KT;
  int ncg;                 // channel groups per time tap
  int passes_per_nt;       // passes of one Cout tile (BN output channels); pass = nt * passes_per_nt + tile group
  int w_stage_bytes, w_stages;
  int tiles_w, tiles_h, num_tiles, num_passes;
  int out_mode, out_T_total, out_t0;
};

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void st_global_32B(void* ptr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f,
                                              uint32_t g, uint32_t h) {   // one full 32-byte sector per thread
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e),
               "r"(f), "r"(g), "r"(h)
               : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// G output tiles per pass, NCH 16-byte channel chunks (8 channels each) per halo stage, BN = Cout
template <int G, int NCH, int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1) conv3d_halo_kernel(const Params p) {
  constexpr int CGK = NCH * 8;                     // channels per halo stage
  constexpr int HALO_TILE = NCH * PLANE;
  constexpr int HALO_STAGE = G * HALO_TILE;
  static_assert(G * BN <= TMEM_COLS, "accumulators do not fit TMEM");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_halo = smem;
  uint8_t* s_w = smem + HALO_BUFS * HALO_STAGE;
  float* s_bias = reinterpret_cast<float*>(smem + SMEM_DATA - BIAS_BYTES);   // [Cout]
  uint64_t* hfull = reinterpret_cast<uint64_t*>(smem + SMEM_DATA);
  uint64_t* hempty = hfull + HALO_BUFS;
  uint64_t* wfull = hempty + HALO_BUFS;
  uint64_t* wempty = wfull + MAX_W_STAGES;
  uint64_t* tfull = wempty + MAX_W_STAGES;         // [G]  accumulator g of the pass is complete
  uint64_t* tempty = tfull + 4;                    // [G]  the epilogue has accumulator g in registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_per_frame = p.tiles_w * p.tiles_h;

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < HALO_BUFS; ++s) {
        mbar_init(&hfull[s], PRODUCERS);
        mbar_init(&hempty[s], 1);
      }
      for (int s = 0; s < p.w_stages; ++s) {
        mbar_init(&wfull[s], 1);
        mbar_init(&wempty[s], 1);
      }
      for (int g = 0; g < G; ++g) {
        mbar_init(&tfull[g], 1);
        mbar_init(&tempty[g], 4);
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

  // tile -> (frame, h0, w0); w fastest, so the G tiles of a pass are horizontal neighbours that share halo columns in L2
  auto decode = [&](int tile, int& t, int& h0, int& w0) {
    t = tile / tiles_per_frame;
    const int r = tile % tiles_per_frame;
    h0 = (r / p.tiles_w) * TH;
    w0 = (r % p.tiles_w) * TW;
  };
  const int stages_per_pass = p.KT * p.ncg;         // halo stages: (kt, channel group)

  if (warp == 0) {
    // ------------------------------------------------------------ weight producer: one bulk copy per (kt, cg, tap)
    if (elect_one()) {
      uint32_t wit = 0;
      for (int pass = blockIdx.x; pass < p.num_passes; pass += gridDim.x) {
        const int nt = pass / p.passes_per_nt;
        for (int hs = 0; hs < stages_per_pass; ++hs) {
          const int kt = hs / p.ncg, cg = hs % p.ncg;
          for (int tap = 0; tap < 9; ++tap, ++wit) {
            const int s = wit % p.w_stages;
            const uint32_t ph = (wit / p.w_stages) & 1;
            mbar_wait(&wempty[s], ph ^ 1, 0xb100 | s);
            mbar_arrive_expect_tx(&wfull[s], p.w_stage_bytes);
            const long long off = ((long long)(((nt * p.KT + kt) * 9 + tap) * (p.Cin / 8) + cg * NCH) * BN) * 8;   // elements
            bulk_load(smem_u32(s_w + s * p.w_stage_bytes), p.w + off, p.w_stage_bytes, &wfull[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
      uint32_t wit = 0, hit = 0, pcount = 0;
      // With 9 weight stages in the ring (BN 96 / 16) stage index == tap, and the LAST halo stage of a pass runs tile by
      // tile instead of tap by tap (all nine taps resident): tile g is complete 9 x CGK/16 MMAs after tile g - 1, so the
      // epilogue drains accumulator g while g + 1 ... are still being computed and the next pass finds its accumulators
      // free — the issuer sat 27 % of its time on tempty when all G tiles finished together (profiles/r02_conv_halo_ncu.txt).
      const bool seq_last = p.w_stages == 9;
      auto mma_tile_tap = [&](uint32_t h_addr, uint32_t w_addr, int g, int tap, int hs) {
        const uint32_t tap_off = ((tap / 3) * HW_ + (tap % 3)) * 16;
#pragma unroll
        for (int j = 0; j < CGK / 16; ++j) {
          const uint64_t adesc = umma_smem_desc(h_addr + g * HALO_TILE + 2 * j * PLANE + tap_off, PLANE, HW_ * 16, 0);
          const uint64_t bdesc = umma_smem_desc(w_addr + 2 * j * (BN * 16), BN * 16, 128, 0);
          umma_ss(tmem_base + g * BN, adesc, bdesc, idesc, (hs | tap | j) != 0);
        }
      };
      for (int pass = blockIdx.x; pass < p.num_passes; pass += gridDim.x, ++pcount) {
        for (int hs = 0; hs < stages_per_pass; ++hs, ++hit) {
          const int hb = hit & 1;
          mbar_wait(&hfull[hb], (hit >> 1) & 1, 0xb200 | hb);
          const uint32_t h_addr = smem_u32(s_halo + hb * HALO_STAGE);
          if (seq_last && hs == stages_per_pass - 1) {
            for (int g = 0; g < G; ++g) {
              if (hs == 0) {               // single-stage pass: this is also the first MMA on accumulator g
                mbar_wait(&tempty[g], (pcount & 1) ^ 1, 0xb400 | g);
                tc_fence_after();
              }
              for (int tap = 0; tap < 9; ++tap) {
                if (g == 0) {
                  mbar_wait(&wfull[tap], ((wit + tap) / 9) & 1, 0xb300 | tap);
                  tc_fence_after();
                }
                mma_tile_tap(h_addr, smem_u32(s_w + tap * p.w_stage_bytes), g, tap, hs);
              }
              umma_commit(&tfull[g]);
            }
            for (int tap = 0; tap < 9; ++tap) umma_commit(&wempty[tap]);
            wit += 9;
          } else {
            for (int tap = 0; tap < 9; ++tap, ++wit) {
              const int s = wit % p.w_stages;
              mbar_wait(&wfull[s], (wit / p.w_stages) & 1, 0xb300 | s);
              tc_fence_after();
              const uint32_t w_addr = smem_u32(s_w + s * p.w_stage_bytes);
#pragma unroll
              for (int g = 0; g < G; ++g) {
                if (hs == 0 && tap == 0) {   // first MMA of the pass on accumulator g: the epilogue must have drained it
                  mbar_wait(&tempty[g], (pcount & 1) ^ 1, 0xb400 | g);
                  tc_fence_after();
                }
                mma_tile_tap(h_addr, w_addr, g, tap, hs);
              }
              umma_commit(&wempty[s]);
            }
            if (hs == stages_per_pass - 1)
              for (int g = 0; g < G; ++g) umma_commit(&tfull[g]);
          }
          umma_commit(&hempty[hb]);
        }
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------ epilogue: one output position per thread. The
    // accumulators of a pass are single-buffered (4 x 96 or 2 x 192 columns), so the next pass's MMAs wait for this code:
    // a tile's accumulator is pulled into registers 96 columns at a time and handed back (tempty) BEFORE bias / residual /
    // packing / stores run — the MMA issuer sat 32 % of its time on tempty when the hand-back came after the stores.
    const int q = warp & 3;
    const int etid = threadIdx.x - 64;
    for (int i = etid; i < p.Cout; i += 128) s_bias[i] = p.bias[i];
    named_bar_sync(1, 128);
    uint32_t pcount = 0;
    for (int pass = blockIdx.x; pass < p.num_passes; pass += gridDim.x, ++pcount) {
      const int r = q * 32 + lane;
      const int nt = pass / p.passes_per_nt, grp = pass % p.passes_per_nt;
#pragma unroll 1
      for (int g = 0; g < G; ++g) {
        mbar_wait(&tfull[g], pcount & 1, 0xb500 | g);
        tc_fence_after();
        const int tile = grp * G + g;
        int t = 0, h0 = 0, w0 = 0;
        const bool tile_ok = tile < p.num_tiles;
        if (tile_ok) decode(tile, t, h0, w0);
        const int h = h0 + r / TW, w = w0 + r % TW;
        const bool ok = tile_ok && h < p.H && w < p.W;
        const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + g * BN;
        const long long pos = ((long long)t * p.H + h) * p.W + w;
        if constexpr (BN == 16) {
          // the decoder head (96 -> 3): fp32 planar [Cout, T_total, H, W], clamped to [-1, 1] (wan_vae.py:668)
          uint32_t acc[16];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]), "=r"(acc[7]),
                "=r"(acc[8]), "=r"(acc[9]), "=r"(acc[10]), "=r"(acc[11]), "=r"(acc[12]), "=r"(acc[13]), "=r"(acc[14]), "=r"(acc[15])
              : "r"(t_row)
              : "memory");
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[g]);
          if (ok) {
            float* o = reinterpret_cast<float*>(p.out);
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i < p.Cout)
                o[(((long long)i * p.out_T_total + p.out_t0 + t) * p.H + h) * p.W + w] =
                    fminf(fmaxf(__uint_as_float(acc[i]) + s_bias[i], -1.f), 1.f);
          }
        } else
#pragma unroll 1
        for (int half = 0; half < BN / 96; ++half) {
          uint32_t acc[96];
          tmem_ld_x32(t_row + half * 96, *reinterpret_cast<uint32_t(*)[32]>(&acc[0]));
          tmem_ld_x32(t_row + half * 96 + 32, *reinterpret_cast<uint32_t(*)[32]>(&acc[32]));
          tmem_ld_x32(t_row + half * 96 + 64, *reinterpret_cast<uint32_t(*)[32]>(&acc[64]));
          tmem_ld_wait();
          if (half == BN / 96 - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[g]);
          }
          if (!ok) continue;
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            const int n0 = nt * BN + half * 96 + c * 16;          // output channel
            float y[16];
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const float4 bb = *reinterpret_cast<const float4*>(&s_bias[n0 + i4 * 4]);
              y[i4 * 4 + 0] = __uint_as_float(acc[c * 16 + i4 * 4 + 0]) + bb.x;
              y[i4 * 4 + 1] = __uint_as_float(acc[c * 16 + i4 * 4 + 1]) + bb.y;
              y[i4 * 4 + 2] = __uint_as_float(acc[c * 16 + i4 * 4 + 2]) + bb.z;
              y[i4 * 4 + 3] = __uint_as_float(acc[c * 16 + i4 * 4 + 3]) + bb.w;
            }
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
            st_global_32B(reinterpret_cast<__nv_bfloat16*>(p.out) + off, pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]),
                          pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]), pack_bf16x2(y[8], y[9]), pack_bf16x2(y[10], y[11]),
                          pack_bf16x2(y[12], y[13]), pack_bf16x2(y[14], y[15]));
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ halo producers (128 threads)
    const int ptid = threadIdx.x - 6 * 32;
    uint32_t hit = 0;
    for (int pass = blockIdx.x; pass < p.num_passes; pass += gridDim.x) {
      int tt[G], th0[G], tw0[G];
      bool tok[G];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int tile = (pass % p.passes_per_nt) * G + g;
        tok[g] = tile < p.num_tiles;
        tt[g] = th0[g] = tw0[g] = 0;
        if (tok[g]) decode(tile, tt[g], th0[g], tw0[g]);
      }
      for (int hs = 0; hs < stages_per_pass; ++hs, ++hit) {
        const int kt = hs / p.ncg, cg = hs % p.ncg;
        const int hb = hit & 1;
        mbar_wait(&hempty[hb], ((hit >> 1) & 1) ^ 1, 0xb600 | hb);
        const uint32_t dst0 = smem_u32(s_halo + hb * HALO_STAGE);
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const __nv_bfloat16* src_t = p.in + (long long)(tt[g] + kt) * p.H * p.W * p.Cin + cg * CGK;
#pragma unroll 2
          for (int idx = ptid; idx < HPOS * NCH; idx += PRODUCERS) {
            const int pos = idx / NCH, c = idx % NCH;     // consecutive threads: consecutive 16-byte chunks of one position
            const int hh = pos / HW_, ww = pos % HW_;
            const int gh = th0[g] - 1 + hh, gw = tw0[g] - 1 + ww;
            const bool valid = tok[g] && gh >= 0 && gh < p.H && gw >= 0 && gw < p.W;
            const __nv_bfloat16* src = src_t + ((long long)(valid ? gh : 0) * p.W + (valid ? gw : 0)) * p.Cin + c * 8;
            cp_async16_zfill(dst0 + g * HALO_TILE + c * PLANE + pos * 16, src, valid);
          }
        }
        cp_async_wait_all();
        fence_proxy_async_smem();          // generic-proxy writes (cp.async) -> visible to the tensor core's async proxy
        mbar_arrive(&hfull[hb]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace convh
}  // namespace sa

extern "C" int sa_conv3d_halo_supported(int32_t Cin, int32_t Cout, int32_t KT, int32_t KH, int32_t KW, int32_t stride_t,
                                        int32_t out_mode) {
  if ((KT != 3 && KT != 1) || KH != 3 || KW != 3 || stride_t > 1 || out_mode < 0 || out_mode > 2) return 0;
  if (out_mode == 2) return Cout <= 16 && Cin % 48 == 0;       // the video head: weights padded to 16 output channels
  if (Cout == 96) return Cin % 48 == 0;
  if (Cout > 0 && Cout % 192 == 0 && Cout <= 384) return Cin % 96 == 0;
  return 0;
}

extern "C" int sa_conv3d_halo_cl(const sa_conv_args* a, sa_stream_t stream_) {
  using namespace sa;
  using namespace sa::convh;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !a->in || !a->w || !a->bias || !a->out) { set_error("sa_conv3d_halo_cl: null pointer"); return SA_ERR_BAD_ARG; }
  if (a->Tout <= 0 || a->H <= 0 || a->W <= 0) { set_error("sa_conv3d_halo_cl: bad dims"); return SA_ERR_BAD_ARG; }
  if (!sa_conv3d_halo_supported(a->Cin, a->Cout, a->KT, a->KH, a->KW, a->stride_t, a->out_mode) ||
      (a->pad_h >= 0 && a->pad_h != 1) || (a->pad_w >= 0 && a->pad_w != 1)) {
    set_error("sa_conv3d_halo_cl: only (1|3)x3x3 'same' stride-1 convs with Cout 96 (Cin %% 48 == 0) or 192 / 384 (Cin %% 96 == 0), "
              "out_mode 0 / 1 (or Cout <= 16 with out_mode 2); got Cin %d Cout %d K %dx%dx%d", a->Cin, a->Cout, a->KT, a->KH, a->KW);
    return SA_ERR_UNSUPPORTED;
  }
  Params p;
  p.in = reinterpret_cast<const __nv_bfloat16*>(a->in);
  p.w = reinterpret_cast<const __nv_bfloat16*>(a->w);
  p.bias = reinterpret_cast<const float*>(a->bias);
  p.res = reinterpret_cast<const __nv_bfloat16*>(a->res);
  p.out = a->out;
  p.Tout = a->Tout; p.H = a->H; p.W = a->W; p.Cin = a->Cin; p.Cout = a->Cout; p.KT = a->KT;
  p.tiles_w = (a->W + TW - 1) / TW; p.tiles_h = (a->H + TH - 1) / TH;
  p.num_tiles = a->Tout * p.tiles_w * p.tiles_h;
  p.out_mode = a->out_mode;
  const bool head = a->out_mode == 2;
  const int G = (head || a->Cout == 96) ? 4 : 2, nch = (head || a->Cout == 96) ? 6 : 12, bn = head ? 16 : (a->Cout == 96 ? 96 : 192);
  p.out_T_total = a->out_T_total; p.out_t0 = a->out_t0;
  p.ncg = a->Cin / (nch * 8);
  p.passes_per_nt = (p.num_tiles + G - 1) / G;
  p.num_passes = p.passes_per_nt * (head ? 1 : a->Cout / bn);
  p.w_stage_bytes = nch * 8 * bn * 2;
  const int halo_bytes = HALO_BUFS * G * nch * PLANE;
  p.w_stages = (SMEM_DATA - BIAS_BYTES - halo_bytes) / p.w_stage_bytes;
  if (p.w_stages > MAX_W_STAGES) p.w_stages = MAX_W_STAGES;
  if (p.w_stages < 2) { set_error("sa_conv3d_halo_cl: shared memory budget"); return SA_ERR_UNSUPPORTED; }
  const int grid = p.num_passes < sm_count() ? p.num_passes : sm_count();
  int rc;
  if (head) {
    if ((rc = ensure_dyn_smem(conv3d_halo_kernel<4, 6, 16>, SMEM_BYTES, "conv3d_halo_kernel"))) return rc;
    conv3d_halo_kernel<4, 6, 16><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(p);
  } else if (a->Cout == 96) {
    if ((rc = ensure_dyn_smem(conv3d_halo_kernel<4, 6, 96>, SMEM_BYTES, "conv3d_halo_kernel"))) return rc;
    conv3d_halo_kernel<4, 6, 96><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(p);
  } else {
    if ((rc = ensure_dyn_smem(conv3d_halo_kernel<2, 12, 192>, SMEM_BYTES, "conv3d_halo_kernel"))) return rc;
    conv3d_halo_kernel<2, 12, 192><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(p);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "conv3d_halo_kernel launch");
  return SA_OK;
}
