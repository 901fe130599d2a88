// attn_tcgen05.cu — flash attention forward, head_dim 128, non-causal, for sm_100a.
//
// One CTA = one (batch, head) x 256 query rows (two 128-row Q tiles that ping-pong on the tensor core).
// Per 128-key KV tile:  S_i = Q_i K^T (SS MMA, fp32 in TMEM)  ->  softmax warpgroup i reads S_i from TMEM, writes
// P_i = exp2(S_i*c - m*c) as bf16 back into the same TMEM columns  ->  O_i += P_i V (TS MMA, A operand from TMEM).
// Running max uses lazy rescaling: O_i/l_i are only rescaled when the row max grew by more than 2^8, so the
// accumulator stays in TMEM almost always (the rescale itself is done by the softmax warpgroup between two PV MMAs).
// The scale is a packed FFMA2 per column pair, the row max uses 3-input FMNMX3, and a compile-time fraction of the
// exponentials runs on the FMA/ALU pipes (Cody-Waite split + cubic) to take load off MUFU.EX2 (16/clk/SM: as many
// cycles as the tile's MMAs, tools/micro/mufu_bench.cu). A 64-key sub-step variant with double-buffered score tiles
// was measured slower (814 vs 1177 TFLOP/s: per-step synchronisation latency dominates), see profiles/README.md.
//
// Warps: 0-3 softmax for Q tile 0, 4-7 softmax for Q tile 1 (TMEM lane quarter = warp % 4), 8 = TMA producer,
// 9 = MMA issuer + TMEM owner. TMEM columns: S0/P0 @0, S1/P1 @128, O0 @256, O1 @384.
// smem: Q 2x32 KB, K 2x32 KB, V 2x32 KB; every 128x128 bf16 tile is two SWIZZLE_128B panels of [128 rows x 64 cols].
//
// Replaces attention() of the reference (wan/models/wan_fantasy_transformer3d_1B.py:158-207, SDPA branch).
#include <stdlib.h>

#include <type_traits>

#include "../../include/stableavatar_b200.h"
#include "sa_host.h"
#include "sa_ptx.cuh"

namespace sa {
namespace attn {

constexpr int BQ = 128, BKV = 128, D = 128;
constexpr int KV_STAGES = 2;
constexpr int PANEL_BYTES = 128 * 64 * 2;  // 16 KB
constexpr int TILE_BYTES = 2 * PANEL_BYTES;
constexpr int NUM_THREADS = 320;
constexpr int TMEM_COLS = 512;
constexpr int SMEM_BYTES = (2 + 2 * KV_STAGES) * TILE_BYTES + 256 + 1024;
constexpr float kRescaleThreshold = 8.0f;  // log2 units
constexpr int kDefaultPolyPairs = 0;       // of every 8 column pairs, how many use the FMA-pipe exp2 (SA_ATTN_POLY overrides)
constexpr int kDefaultImpl = 8;            // 2: this file; 4: decoupled kernel (attn_v4_tcgen05.cu); 8: decoupled + one
                                           // MMA-issuing warp per Q tile (attn_v8_tcgen05.cu). SA_ATTN_IMPL overrides.

struct Params {
  __nv_bfloat16* out;
  long long o_bs, o_ls;
  int q_len, kv_len;
  float scale_log2;
  int accumulate;
};

template <int kPolyPairs>
__global__ void __launch_bounds__(NUM_THREADS, 1)
flash_attn_d128_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                       const __grid_constant__ CUtensorMap tmap_v, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                  // [2][TILE]
  uint8_t* sK = smem + 2 * TILE_BYTES;                 // [KV_STAGES][TILE]
  uint8_t* sV = sK + KV_STAGES * TILE_BYTES;           // [KV_STAGES][TILE]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + KV_STAGES * TILE_BYTES);
  uint64_t* q_full = bars;              // [2]
  uint64_t* k_full = bars + 2;          // [2]
  uint64_t* k_empty = bars + 4;         // [2]
  uint64_t* v_full = bars + 6;          // [2]
  uint64_t* v_empty = bars + 8;         // [2]
  uint64_t* s_full = bars + 10;         // [2]  S_i ready in TMEM
  uint64_t* p_full = bars + 12;         // [2]  P_i written (and O_i rescaled if needed)
  uint64_t* o_final = bars + 14;        // [2]  last PV_i done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * BQ);
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int n_kv = (p.kv_len + BKV - 1) / BKV;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
  }
  if (warp == 9) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&q_full[i], 1);
        mbar_init(&k_full[i], 1);
        mbar_init(&k_empty[i], 1);
        mbar_init(&v_full[i], 1);
        mbar_init(&v_empty[i], 1);
        mbar_init(&s_full[i], 1);
        mbar_init(&p_full[i], 4);  // one arrive per softmax warp
        mbar_init(&o_final[i], 1);
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

  if (warp == 8) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {  // elect.sync: single active lane is known to ptxas -> no R2UR waterfall per UTCHMMA / UTMALDG
      auto load_tile = [&](uint8_t* dst, const CUtensorMap* m, uint64_t* bar, int row0) {
        mbar_arrive_expect_tx(bar, TILE_BYTES);
        tma_load_4d(dst, m, bar, 0, row0, head, b);
        tma_load_4d(dst + PANEL_BYTES, m, bar, 64, row0, head, b);
      };
      load_tile(sQ, &tmap_q, &q_full[0], q0);
      load_tile(sK, &tmap_k, &k_full[0], 0);
      load_tile(sQ + TILE_BYTES, &tmap_q, &q_full[1], q0 + BQ);
      load_tile(sV, &tmap_v, &v_full[0], 0);
      for (int j = 1; j < n_kv; ++j) {
        const int s = j % KV_STAGES;
        const uint32_t ph = (j / KV_STAGES) & 1;
        mbar_wait(&k_empty[s], ph ^ 1, 0x1100 | s);
        load_tile(sK + s * TILE_BYTES, &tmap_k, &k_full[s], j * BKV);
        mbar_wait(&v_empty[s], ph ^ 1, 0x1200 | s);
        load_tile(sV + s * TILE_BYTES, &tmap_v, &v_full[s], j * BKV);
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {  // elect.sync: single active lane is known to ptxas -> no R2UR waterfall per UTCHMMA / UTMALDG
      const uint32_t idesc_qk = umma_idesc_bf16(BQ, BKV, 0, 0);  // A = Q (K-major), B = K (K-major)
      const uint32_t idesc_pv = umma_idesc_bf16(BQ, D, 0, 1);    // A = P (TMEM),   B = V (MN-major: d contiguous)
      auto issue_S = [&](int i, int ks) {
        const uint32_t qa = smem_u32(sQ + i * TILE_BYTES);
        const uint32_t ka = smem_u32(sK + ks * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t off = (k >> 2) * PANEL_BYTES + (k & 3) * 32;
          umma_ss(tmem_base + i * 128, umma_smem_desc(qa + off, 16, 1024, kSwz128),
                  umma_smem_desc(ka + off, 16, 1024, kSwz128), idesc_qk, k != 0);
        }
      };
      auto issue_PV = [&](int i, int vs, bool acc) {
        const uint32_t va = smem_u32(sV + vs * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          // 16 keys = 16 rows of 128 B inside each panel; LBO = distance between the two 64-wide d panels.
          umma_ts(tmem_base + 256 + i * 128, tmem_base + i * 128 + k * 8,
                  umma_smem_desc(va + k * 2048, PANEL_BYTES, 1024, kSwz128), idesc_pv, (acc || k != 0) ? 1u : 0u);
        }
      };
      mbar_wait(&q_full[0], 0, 0x1300);
      mbar_wait(&k_full[0], 0, 0x1310);
      tc_fence_after();
      issue_S(0, 0);
      umma_commit(&s_full[0]);
      mbar_wait(&q_full[1], 0, 0x1301);
      tc_fence_after();
      issue_S(1, 0);
      umma_commit(&s_full[1]);
      umma_commit(&k_empty[0]);
      for (int j = 0; j < n_kv; ++j) {
        const int vs = j % KV_STAGES;
        const uint32_t vph = (j / KV_STAGES) & 1;
        const bool last = (j == n_kv - 1);
        const int ks = (j + 1) % KV_STAGES;
        mbar_wait(&v_full[vs], vph, 0x1320 | vs);
        mbar_wait(&p_full[0], j & 1, 0x1330);
        tc_fence_after();
        issue_PV(0, vs, j > 0);
        if (last) umma_commit(&o_final[0]);
        if (!last) {
          mbar_wait(&k_full[ks], ((j + 1) / KV_STAGES) & 1, 0x1340 | ks);
          tc_fence_after();
          issue_S(0, ks);
          umma_commit(&s_full[0]);
        }
        mbar_wait(&p_full[1], j & 1, 0x1331);
        tc_fence_after();
        issue_PV(1, vs, j > 0);
        umma_commit(&v_empty[vs]);
        if (last) umma_commit(&o_final[1]);
        if (!last) {
          issue_S(1, ks);
          umma_commit(&s_full[1]);
          umma_commit(&k_empty[ks]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax warpgroups (one thread per query row)
    const int i = warp >> 2;  // Q tile
    const int quarter = warp & 3;
    const uint32_t lane_sel = uint32_t(quarter * 32) << 16;
    const uint32_t tS = tmem_base + lane_sel + i * 128;
    const uint32_t tO = tmem_base + lane_sel + 256 + i * 128;
    const int row = q0 + i * BQ + quarter * 32 + lane;
    const float c = p.scale_log2;
    const uint64_t c2 = pack_f32x2(c, c);
    float m_ref = 0.f;
    uint64_t lsum2 = pack_f32x2(0.f, 0.f);

    auto step = [&](int j, auto mask_tag) {
      constexpr bool MASK = decltype(mask_tag)::value;
      mbar_wait(&s_full[i], j & 1, 0x1400 | i);
      tc_fence_after();
      uint32_t s[128];
      {
        uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
        uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
        uint32_t(&s2)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[64]);
        uint32_t(&s3)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[96]);
        tmem_ld_x32(tS, s0);
        tmem_ld_x32(tS + 32, s1);
        tmem_ld_x32(tS + 64, s2);
        tmem_ld_x32(tS + 96, s3);
        tmem_ld_wait();
      }
      if constexpr (MASK) {
        const int valid = p.kv_len - j * BKV;
#pragma unroll
        for (int t = 0; t < 128; ++t)
          if (t >= valid) s[t] = 0xff800000u;  // -inf
      }
      // row max: four independent chains of 3-input max (FMNMX3), 2 new values per instruction
      float mx0 = __uint_as_float(s[0]), mx1 = __uint_as_float(s[1]), mx2 = __uint_as_float(s[2]), mx3 = __uint_as_float(s[3]);
#pragma unroll
      for (int t = 4; t < 128; t += 8) {
        mx0 = max3(mx0, __uint_as_float(s[t]), __uint_as_float(s[t + 1]));
        mx1 = max3(mx1, __uint_as_float(s[t + 2]), __uint_as_float(s[t + 3]));
        if (t + 4 < 128) {
          mx2 = max3(mx2, __uint_as_float(s[t + 4]), __uint_as_float(s[t + 5]));
          mx3 = max3(mx3, __uint_as_float(s[t + 6]), __uint_as_float(s[t + 7]));
        }
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));

      float alpha = 1.0f;
      bool need = false;
      if (j == 0) {
        m_ref = mx;
      } else if ((mx - m_ref) * c > kRescaleThreshold) {
        alpha = ex2_approx((m_ref - mx) * c);
        m_ref = mx;
        need = true;
      }
      if (__any_sync(0xffffffffu, need)) {
        // PV_i(j-1) has completed (its commit precedes S_i(j)'s), PV_i(j) waits for p_full: O_i is ours to rescale.
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          uint32_t o[32];
          tmem_ld_x32(tO + cc * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 32; ++t) o[t] = __float_as_uint(__uint_as_float(o[t]) * alpha);
          tmem_st_x32(tO + cc * 32, o);
        }
        float l_lo, l_hi;
        unpack_f32x2(lsum2, l_lo, l_hi);
        lsum2 = pack_f32x2(l_lo * alpha, l_hi * alpha);
      }
      // p = 2^(s*c - m*c): one packed FFMA2 per column pair for the scale; kPolyPairs of every 8 pairs use the cubic
      const float nmc = -m_ref * c;
      const uint64_t nmc2 = pack_f32x2(nmc, nmc);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t pk[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const uint64_t x2 = fma_f32x2(pack_f32x2(__uint_as_float(s[cc * 32 + 2 * t]), __uint_as_float(s[cc * 32 + 2 * t + 1])), c2, nmc2);
          uint64_t p2;
          if ((t & 7) < kPolyPairs) {
            float x0, x1;
            unpack_f32x2(x2, x0, x1);
            x0 = fmaxf(x0, -126.0f);
            x1 = fmaxf(x1, -126.0f);
            const uint64_t xc = pack_f32x2(x0, x1);
            const uint64_t xi = add_f32x2(xc, pack_f32x2(12582912.0f, 12582912.0f));   // low mantissa bits = round(x)
            const uint64_t xr = add_f32x2(xi, pack_f32x2(-12582912.0f, -12582912.0f));
            const uint64_t f = fma_f32x2(xr, pack_f32x2(-1.0f, -1.0f), xc);             // x - round(x) in [-0.5, 0.5]
            uint64_t q = fma_f32x2(f, pack_f32x2(0.05550411f, 0.05550411f), pack_f32x2(0.24022651f, 0.24022651f));
            q = fma_f32x2(q, f, pack_f32x2(0.69314718f, 0.69314718f));
            q = fma_f32x2(q, f, pack_f32x2(1.0f, 1.0f));
            float q0, q1, i0, i1;
            unpack_f32x2(q, q0, q1);
            unpack_f32x2(xi, i0, i1);
            q0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(i0) << 23));
            q1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(i1) << 23));
            p2 = pack_f32x2(q0, q1);
            pk[t] = pack_bf16x2(q0, q1);
          } else {
            float x0, x1;
            unpack_f32x2(x2, x0, x1);
            const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
            p2 = pack_f32x2(p0, p1);
            pk[t] = pack_bf16x2(p0, p1);
          }
          lsum2 = add_f32x2(lsum2, p2);
        }
        tmem_st_x16(tS + cc * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[i]);
    };

    const bool ragged = (p.kv_len % BKV) != 0;
    for (int j = 0; j < n_kv - 1; ++j) step(j, std::false_type{});
    if (ragged) step(n_kv - 1, std::true_type{});
    else step(n_kv - 1, std::false_type{});

    // ---- epilogue: O_i / l -> bf16 -> global
    mbar_wait(&o_final[i], 0, 0x1500 | i);
    tc_fence_after();
    float l_lo, l_hi;
    unpack_f32x2(lsum2, l_lo, l_hi);
    const float inv_l = 1.0f / (l_lo + l_hi);
    const bool row_ok = row < p.q_len;
    __nv_bfloat16* orow = p.out + (long long)b * p.o_bs + (long long)row * p.o_ls + head * D;
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      uint32_t o[32];
      tmem_ld_x32(tO + cc * 32, o);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int v8 = 0; v8 < 4; ++v8) {
          float y[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) y[t] = __uint_as_float(o[v8 * 8 + t]) * inv_l;
          uint4* dst = reinterpret_cast<uint4*>(orow + cc * 32 + v8 * 8);
          if (p.accumulate) {
            uint4 old = *dst;
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&old);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              float2 f = __bfloat1622float2(h[t]);
              y[2 * t] = f.x + bf16_round(y[2 * t]);
              y[2 * t + 1] = f.y + bf16_round(y[2 * t + 1]);
            }
          }
          uint4 u;
          u.x = pack_bf16x2(y[0], y[1]);
          u.y = pack_bf16x2(y[2], y[3]);
          u.z = pack_bf16x2(y[4], y[5]);
          u.w = pack_bf16x2(y[6], y[7]);
          *dst = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace attn
}  // namespace sa

int sa_flash_attn_d128_v4(const sa_attn_args* a, int poly, cudaStream_t stream);  // attn_v4_tcgen05.cu
int sa_flash_attn_d128_v8(const sa_attn_args* a, int poly, cudaStream_t stream);  // attn_v8_tcgen05.cu

extern "C" int sa_flash_attn_d128(const sa_attn_args* a, sa_stream_t stream_) {
  using namespace sa;
  using namespace sa::attn;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !a->q || !a->k || !a->v || !a->out) { set_error("sa_flash_attn_d128: null pointer"); return SA_ERR_BAD_ARG; }
  if (a->batch <= 0 || a->heads <= 0 || a->q_len <= 0 || a->kv_len <= 0) {
    set_error("sa_flash_attn_d128: non-positive dims");
    return SA_ERR_BAD_ARG;
  }
  if (a->q_ls % 8 || a->k_ls % 8 || a->v_ls % 8 || a->o_ls % 8 || a->q_bs % 8 || a->k_bs % 8 || a->v_bs % 8 ||
      a->o_bs % 8) {
    set_error("sa_flash_attn_d128: strides must be multiples of 8 elements");
    return SA_ERR_BAD_ARG;
  }
  static int impl = -1, poly_env = -1;
  if (impl < 0) {
    const char* e1 = getenv("SA_ATTN_IMPL");
    impl = e1 ? atoi(e1) : kDefaultImpl;
    const char* e2 = getenv("SA_ATTN_POLY");
    poly_env = e2 ? atoi(e2) : kDefaultPolyPairs;
    if (poly_env < 0 || poly_env > 4) poly_env = kDefaultPolyPairs;
  }
  if (impl == 4) return sa_flash_attn_d128_v4(a, poly_env, stream);
  if (impl == 8) return sa_flash_attn_d128_v8(a, poly_env, stream);
  CUtensorMap tq, tk, tv;
  auto mk = [&](CUtensorMap* m, const void* base, int len, long long ls, long long bs) {
    uint64_t dims[4] = {(uint64_t)D, (uint64_t)len, (uint64_t)a->heads, (uint64_t)a->batch};
    uint64_t strides[3] = {(uint64_t)ls * 2, (uint64_t)D * 2, (uint64_t)bs * 2};
    uint32_t box[4] = {64, 128, 1, 1};
    return make_tmap_bf16(m, base, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  };
  int rc;
  if ((rc = mk(&tq, a->q, a->q_len, a->q_ls, a->q_bs))) return rc;
  if ((rc = mk(&tk, a->k, a->kv_len, a->k_ls, a->k_bs))) return rc;
  if ((rc = mk(&tv, a->v, a->kv_len, a->v_ls, a->v_bs))) return rc;
  Params p;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  p.o_bs = a->o_bs;
  p.o_ls = a->o_ls;
  p.q_len = a->q_len;
  p.kv_len = a->kv_len;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.accumulate = a->accumulate;
  static int poly = -1;
  if (poly < 0) {
    const char* env = getenv("SA_ATTN_POLY");
    poly = env ? atoi(env) : kDefaultPolyPairs;
    if (poly < 0 || poly > 4) poly = kDefaultPolyPairs;
    cudaError_t e = cudaSuccess;
    auto set = [&](auto* k) { if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES); };
    set(flash_attn_d128_kernel<0>); set(flash_attn_d128_kernel<1>); set(flash_attn_d128_kernel<2>);
    set(flash_attn_d128_kernel<3>); set(flash_attn_d128_kernel<4>);
    if (e != cudaSuccess) { poly = -1; return cuda_fail(e, "cudaFuncSetAttribute(flash_attn_d128_kernel)"); }
  }
  dim3 grid((a->q_len + 2 * BQ - 1) / (2 * BQ), a->heads, a->batch);
  switch (poly) {
    case 0: flash_attn_d128_kernel<0><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tq, tk, tv, p); break;
    case 1: flash_attn_d128_kernel<1><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tq, tk, tv, p); break;
    case 3: flash_attn_d128_kernel<3><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tq, tk, tv, p); break;
    case 4: flash_attn_d128_kernel<4><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tq, tk, tv, p); break;
    default: flash_attn_d128_kernel<2><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tq, tk, tv, p); break;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "flash_attn_d128_kernel launch");
  return SA_OK;
}
