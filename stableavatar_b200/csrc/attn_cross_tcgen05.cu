// attn_cross_tcgen05.cu — the three cross-attentions of a DiT block in ONE persistent launch (sm_100a, head_dim 128).
//
// WanI2VTalkingCrossAttention (wan/models/wan_fantasy_transformer3d_1B.py:534-605) sums three softmax attentions that
// share the query: text (512 keys), CLIP image (257 keys) and audio (15 keys of the latent frame's audio window, paired
// with the token group by q.view(b * G, -1, n, d), :575-586). With only 8 + 5 + 1 key steps per Q tile a short-lived CTA
// is a serial latency chain: the round-1 kernel (one 128-row tile per CTA, profiles/experiments/
// attn_cross_one_tile_per_cta.cu.txt) spent ~10 us per CTA on launch, TMEM allocation, the Q load, the first TMA and three
// read-modify-write epilogues — tensor pipe 18 % active, 1.2 ms per launch for 0.47 TFLOP. This kernel is PERSISTENT:
//
//   * one CTA per SM walks work items (batch, head, 256 query rows) with a grid stride; barriers, TMEM and tensor maps
//     are set up once, every mbarrier phase keeps running across key sets and items (global step counters);
//   * an item is TWO 128-row Q tiles on the decoupled pipeline of attn_self_tcgen05.cu (Q in TMEM, S = Q K^T as a TS MMA,
//     P through a double-buffered shared-memory panel, one MMA-issuing warp and one softmax warpgroup per tile), so each
//     K/V tile fetched from L2 serves 256 rows and one tile's fill / drain / epilogue overlaps the other's steady state;
//   * the TMA producer runs ahead across sets and items through a 3-stage ring, so the next item's first keys are in
//     shared memory before its Q is;
//   * the per-set results are summed in a shared-memory staging tile (bf16, every partial result rounded to bf16 before
//     the bf16 add, set order text, image, audio — the arithmetic of three separate launches; bit for bit for the plain
//     sets; the windowed step keeps all its exponentials on MUFU, see the step body) and leave the
//     SM once per item as a TMA store (rows beyond q_len are clipped by the tensor map).
//
// The audio set is one 64-key step: the keys of the (at most 64 / A) consecutive windows that the item's 256 rows can
// touch are loaded together and every row masks the step down to its own window [ (g - g0) A, (g - g0) A + A ), with
// g = (tok_offset + row) / rows_per_group — which also covers token shards of the sequence-parallel path.
#include <stdlib.h>

#include <type_traits>

#include "../../include/stableavatar_b200.h"
#include "sa_host.h"
#include "sa_ptx.cuh"

namespace sa {
namespace attnx {

constexpr int BQ = 128, SUB = 64, D = 128;
constexpr int STAGES = 3;
constexpr int KV_PANEL = SUB * 128;          // 64 rows x 128 B = 8 KB
constexpr int KV_TILE = 2 * KV_PANEL;        // [64 keys x 128 d] = 16 KB
constexpr int STAGE_BYTES = 2 * KV_TILE;     // K + V
constexpr int P_BYTES = BQ * 128;            // [128 rows x 64 keys] bf16 = 16 KB
constexpr int OUT_PANEL = BQ * 128;          // [128 rows x 64 cols] bf16 = 16 KB, SWIZZLE_128B (the TMA store's layout)
constexpr int OUT_BYTES = 2 * OUT_PANEL;     // one Q tile's [128 x 128] output
constexpr int NUM_THREADS = 352;   // 8 softmax warps (two Q tiles), TMA producer, one MMA-issuing warp per Q tile
constexpr int TMEM_COLS = 512;     // Q0 0-63, Q1 64-127, S0 128-191, S1 192-255, O0 256-383, O1 384-511
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 4 * P_BYTES + 2 * OUT_BYTES + 256 + 1024;
constexpr float kRescaleThreshold = 8.0f;

constexpr int MAX_SETS = 3;
struct Params {
  const __nv_bfloat16* q;
  __nv_bfloat16* out;
  long long q_bs, q_ls, o_bs, o_ls;
  int q_len, heads, n_pairs, n_items;
  float scale_log2;
  int accumulate;
  int n_sets;
  int kv_len[MAX_SETS];     // keys of the set (windowed set: keys per window)
  int windowed[MAX_SETS];   // 1: per-row window of kv_len keys inside one 64-key step
  int rows_per_group, tok_offset;
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
cross_attn_kernel(const __grid_constant__ CUtensorMap tk0, const __grid_constant__ CUtensorMap tv0,
                  const __grid_constant__ CUtensorMap tk1, const __grid_constant__ CUtensorMap tv1,
                  const __grid_constant__ CUtensorMap tk2, const __grid_constant__ CUtensorMap tv2,
                  const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ CUtensorMap tmap_q, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sKV = smem;                                   // [STAGES][K tile | V tile]
  uint8_t* sP = smem + STAGES * STAGE_BYTES;             // [2 q tiles][2 buffers][P_BYTES]
  uint8_t* sOut = sP + 4 * P_BYTES;                      // [2 q tiles][2 panels][OUT_PANEL]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + 2 * OUT_BYTES);
  uint64_t* kv_full = bars;                 // [STAGES]
  uint64_t* kv_empty = bars + STAGES;       // [STAGES]
  uint64_t* s_full = bars + 2 * STAGES;     // [2]  S_i(U) complete in TMEM
  uint64_t* s_cons = s_full + 2;            // [2]  softmax i has S_i(U) in registers
  uint64_t* p_full = s_cons + 2;            // [2 tiles][2 P buffers]  P_i(U) in shared memory (and O_i rescaled if needed)
  uint64_t* pv_done = p_full + 4;           // [2]  one completion per P_i(U) V
  uint64_t* o_final = pv_done + 2;          // [2]  one completion per key set: the set's accumulator is complete
  uint64_t* q_ready = o_final + 2;          // [2]  one completion per item: Q tile i stored in TMEM
  uint64_t* o_free = q_ready + 2;           // [2]  one completion per key set: its epilogue has O_i in registers
  uint64_t* q_full = o_free + 2;            // [2]  one completion per item: Q tile i has landed in tile i's P buffers (TMA)
  uint64_t* res_full = q_full + 2;          // [2]  one completion per item: the residual tile has landed in the staging tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  auto n_sub_of = [&](int si) { return p.windowed[si] ? 1 : (p.kv_len[si] + SUB - 1) / SUB; };

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tk0);
    tma_prefetch_desc(&tv0);
    tma_prefetch_desc(&tmap_o);
    tma_prefetch_desc(&tmap_q);
  }
  if (warp == 9) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&kv_full[s], 1);
        mbar_init(&kv_empty[s], 2);   // one tcgen05.commit per MMA warp
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1);
        mbar_init(&s_cons[i], 4);
        mbar_init(&pv_done[i], 1);
        mbar_init(&o_final[i], 1);
        mbar_init(&q_ready[i], 4);
        mbar_init(&o_free[i], 4);
        mbar_init(&q_full[i], 1);
        mbar_init(&res_full[i], 1);
      }
      for (int i = 0; i < 4; ++i) mbar_init(&p_full[i], 4);
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
    // ------------------------------------------------------------ TMA producer: {K, V} 64-key tiles, running ahead
    // across key sets and work items through the ring
    if (elect_one()) {  // elect.sync: single active lane is known to ptxas -> no R2UR waterfall per UTCHMMA / UTMALDG
      int U = 0;
      for (int w = blockIdx.x; w < p.n_items; w += gridDim.x) {
        const int pt = w % p.n_pairs, head = (w / p.n_pairs) % p.heads, b = w / (p.n_pairs * p.heads);
        const int g0 = (p.tok_offset + pt * 2 * BQ) / p.rows_per_group;   // first window any row of this item can belong to
        for (int si = 0; si < p.n_sets; ++si) {
          const CUtensorMap* tk = si == 0 ? &tk0 : (si == 1 ? &tk1 : &tk2);
          const CUtensorMap* tv = si == 0 ? &tv0 : (si == 1 ? &tv1 : &tv2);
          const int row0 = p.windowed[si] ? g0 * p.kv_len[si] : 0;
          const int ns = n_sub_of(si);
          for (int u = 0; u < ns; ++u, ++U) {
            const int s = U % STAGES;
            const uint32_t ph = (U / STAGES) & 1;
            mbar_wait(&kv_empty[s], ph ^ 1, 0x9100 | s);
            mbar_arrive_expect_tx(&kv_full[s], STAGE_BYTES);
            uint8_t* st = sKV + s * STAGE_BYTES;
            tma_load_4d(st, tk, &kv_full[s], 0, row0 + u * SUB, head, b);
            tma_load_4d(st + KV_PANEL, tk, &kv_full[s], 64, row0 + u * SUB, head, b);
            tma_load_4d(st + KV_TILE, tv, &kv_full[s], 0, row0 + u * SUB, head, b);
            tma_load_4d(st + KV_TILE + KV_PANEL, tv, &kv_full[s], 64, row0 + u * SUB, head, b);
          }
        }
      }
    }
  } else if (warp == 9 || warp == 10) {
    // ------------------------------------------------------------ MMA issuers, one warp per Q tile
    if (elect_one()) {
      const int i = warp - 9;
      constexpr uint32_t idesc_qk = umma_idesc_bf16(BQ, SUB, 0, 0);  // A = Q (TMEM), B = 64 keys of K (K-major)
      constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, D, 0, 1);    // A = P (smem, K-major), B = V (MN-major)
      // descriptors as in attn_self_tcgen05.cu: a 32-bit low word (address >> 4 | LBO) advanced by constants, one constant
      // high word (SBO 1024, version 1, SWIZZLE_128B) — with 8 + 5 + 1 key steps per item this thread's serial path is on
      // the critical chain at every item and set boundary
      constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (kSwz128 << 29);
      const uint32_t k_lo0 = ((smem_u32(sKV) & 0x3ffff) >> 4) | (1u << 16);
      const uint32_t v_lo0 = ((smem_u32(sKV + KV_TILE) & 0x3ffff) >> 4) | (uint32_t(KV_PANEL >> 4) << 16);
      const uint32_t p_lo0 = ((smem_u32(sP + i * 2 * P_BYTES) & 0x3ffff) >> 4) | (1u << 16);
      const uint32_t tSi = tmem_base + 128 + i * SUB, tQi = tmem_base + i * 64, tOi = tmem_base + 256 + i * 128;
      auto issue_S = [&](int U) {
        const uint32_t k_lo = k_lo0 + (U % STAGES) * (STAGE_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma_ts_lh(tSi, tQi + k * 8, k_lo + (k >> 2) * (KV_PANEL >> 4) + (k & 3) * 2, DESC_HI, idesc_qk, k != 0);
        umma_commit(&s_full[i]);
      };
      auto issue_PV = [&](int U, bool first) {
        const uint32_t v_lo = v_lo0 + (U % STAGES) * (STAGE_BYTES >> 4), p_lo = p_lo0 + (U & 1) * (P_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < SUB / 16; ++k)
          umma_ss_lh(tOi, p_lo + k * 2, DESC_HI, v_lo + k * (2048 >> 4), DESC_HI, idesc_pv, (!first || k != 0) ? 1u : 0u);
        umma_commit(&pv_done[i]);
      };
      int U = 0, N = 0, it = 0;   // global step, key-set and item counters: barrier phases keep running
      for (int w = blockIdx.x; w < p.n_items; w += gridDim.x, ++it) {
        // first scores of the item: Q_i is in TMEM (which implies every softmax warp consumed the previous item's last S_i)
        mbar_wait(&q_ready[i], it & 1, 0x9300 | i);
        if (U > 0) mbar_wait(&s_cons[i], (U - 1) & 1, 0x9360 | i);
        mbar_wait(&kv_full[U % STAGES], (U / STAGES) & 1, 0x9310 | (U % STAGES));
        tc_fence_after();
        issue_S(U);
        for (int si = 0; si < p.n_sets; ++si, ++N) {
          const int ns = n_sub_of(si);
          for (int u = 0; u < ns; ++u, ++U) {
            const bool last_of_item = si == p.n_sets - 1 && u == ns - 1;
            if (!last_of_item) {
              mbar_wait(&kv_full[(U + 1) % STAGES], ((U + 1) / STAGES) & 1, 0x9320 | ((U + 1) % STAGES));
              mbar_wait(&s_cons[i], U & 1, 0x9330 | i);   // S_i(U) is in the softmax registers: its columns are free
              tc_fence_after();
              issue_S(U + 1);
            }
            mbar_wait(&p_full[i * 2 + (U & 1)], (U >> 1) & 1, 0x9340 | (i * 2 + (U & 1)));
            if (u == 0 && N > 0) mbar_wait(&o_free[i], (N - 1) & 1, 0x9350 | i);   // the previous set's O_i has been read out
            tc_fence_after();
            issue_PV(U, u == 0);
            if (u == ns - 1) umma_commit(&o_final[i]);
            umma_commit(&kv_empty[U % STAGES]);   // this tile is done with K(U), V(U)
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax warpgroups (one thread per query row)
    const int i = warp >> 2;  // Q tile
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // row inside the tile
    const uint32_t lane_sel = uint32_t(quarter * 32) << 16;
    const uint32_t tQ = tmem_base + lane_sel + i * 64;
    const uint32_t tS = tmem_base + lane_sel + 128 + i * SUB;
    const uint32_t tO = tmem_base + lane_sel + 256 + i * 128;
    uint8_t* p_row0 = sP + i * 2 * P_BYTES + r * 128;
    uint8_t* out_tile = sOut + i * OUT_BYTES;
    uint8_t* out_row = out_tile + r * 128;
    const bool store_thread = quarter == 0 && lane == 0;   // issues and retires this tile's TMA stores

    const float c = p.scale_log2;
    const uint64_t c2 = pack_f32x2(c, c);
    float m_ref = 0.f;
    uint64_t lsum2 = pack_f32x2(0.f, 0.f), lsum2b = pack_f32x2(0.f, 0.f);  // two independent row-sum chains

    // One thread per tile feeds the tile's staging buffers by TMA: Q of item w into the P buffers, and (accumulate mode) the
    // rows the result is added to into the output staging tile, so that every epilogue reads `old` from shared memory.
    // Non-accumulate mode: the next item's first epilogue may only write the staging tile once this item's store has read it
    // — the store thread retires the store (bulk_wait_read0) before it next arrives anywhere the other threads wait on.
    auto load_q = [&](int w) {
      const int pt = w % p.n_pairs, head = (w / p.n_pairs) % p.heads, b = w / (p.n_pairs * p.heads);
      const int t0 = pt * 2 * BQ + i * BQ;
      mbar_arrive_expect_tx(&q_full[i], 2 * P_BYTES);
      tma_load_4d(sP + i * 2 * P_BYTES, &tmap_q, &q_full[i], 0, t0, head, b);
      tma_load_4d(sP + i * 2 * P_BYTES + P_BYTES, &tmap_q, &q_full[i], 64, t0, head, b);
    };
    auto load_res = [&](int w) {
      if (!p.accumulate) return;
      const int pt = w % p.n_pairs, head = (w / p.n_pairs) % p.heads, b = w / (p.n_pairs * p.heads);
      const int t0 = pt * 2 * BQ + i * BQ;
      mbar_arrive_expect_tx(&res_full[i], OUT_BYTES);
      tma_load_4d(out_tile, &tmap_o, &res_full[i], 0, t0, head, b);
      tma_load_4d(out_tile + OUT_PANEL, &tmap_o, &res_full[i], 64, t0, head, b);
    };
    if (store_thread && (int)blockIdx.x < p.n_items) {
      load_res(blockIdx.x);
      load_q(blockIdx.x);
    }

    int U = 0, N = 0, it = 0;
    for (int w = blockIdx.x; w < p.n_items; w += gridDim.x, ++it) {
      const int pt = w % p.n_pairs, head = (w / p.n_pairs) % p.heads, b = w / (p.n_pairs * p.heads);
      const int q0 = pt * 2 * BQ;
      const int g0 = (p.tok_offset + q0) / p.rows_per_group;
      const int row = q0 + i * BQ + r;
      const bool row_ok = row < p.q_len;

      // ---- Q row -> TMEM (A operand layout: lane = row, column c holds elements 2c, 2c+1). The tile arrived by TMA in
      // this Q tile's two (idle) P buffers — issued at the end of the previous item, see below — as two SWIZZLE_128B panels;
      // every thread reads only its own row, as it later writes only its own P row. (Round 1/2 loaded the row with sixteen
      // 16-byte global loads per thread: 32 different lines per warp instruction, ~3 us of L1 wavefronts per item, and the
      // same again for the residual row in the first epilogue.) The previous item's last score MMA has completed (this
      // thread waited for its s_full), so the TMEM columns are free.
      {
        mbar_wait(&q_full[i], it & 1, 0x9410 | i);
        const uint8_t* qrow = sP + i * 2 * P_BYTES + r * 128;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t wq[32];
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) {
            const uint4 v = *reinterpret_cast<const uint4*>(qrow + h * P_BYTES + ((cc ^ (r & 7)) << 4));
            wq[cc * 4] = v.x; wq[cc * 4 + 1] = v.y; wq[cc * 4 + 2] = v.z; wq[cc * 4 + 3] = v.w;
          }
          tmem_st_x32(tQ + h * 32, wq);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&q_ready[i]);
      }

      // U: global step (barrier phases, P buffer), u: step inside the current set, MASK: 0 none, 1 ragged tail, 2 window
      auto step = [&](int u, int kv_len, auto mask_tag) {
        constexpr int MASK = decltype(mask_tag)::value;
        mbar_wait(&s_full[i], U & 1, 0x9400 | i);
        tc_fence_after();
        uint32_t s[SUB];
        {
          uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
          uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
          tmem_ld_x32(tS, s0);
          tmem_ld_x32(tS + 32, s1);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_cons[i]);   // the MMA warp may overwrite S_i with the next step's scores
        if constexpr (MASK == 1) {
          const int valid = kv_len - u * SUB;
#pragma unroll
          for (int t = 0; t < SUB; ++t)
            if (t >= valid) s[t] = 0xff800000u;  // -inf
        } else if constexpr (MASK == 2) {
          const int lo = ((p.tok_offset + (row_ok ? row : q0)) / p.rows_per_group - g0) * kv_len, hi = lo + kv_len;
#pragma unroll
          for (int t = 0; t < SUB; ++t)
            if (t < lo || t >= hi) s[t] = 0xff800000u;  // keys of other windows
        }
        float mx0 = __uint_as_float(s[0]), mx1 = __uint_as_float(s[1]), mx2 = __uint_as_float(s[2]), mx3 = __uint_as_float(s[3]);
#pragma unroll
        for (int t = 4; t < SUB; t += 8) {
          mx0 = max3(mx0, __uint_as_float(s[t]), __uint_as_float(s[t + 1]));
          mx1 = max3(mx1, __uint_as_float(s[t + 2]), __uint_as_float(s[t + 3]));
          if (t + 4 < SUB) {
            mx2 = max3(mx2, __uint_as_float(s[t + 4]), __uint_as_float(s[t + 5]));
            mx3 = max3(mx3, __uint_as_float(s[t + 6]), __uint_as_float(s[t + 7]));
          }
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));

        float alpha = 1.0f;
        bool need = false;
        if (u == 0) {
          m_ref = mx;
        } else if ((mx - m_ref) * c > kRescaleThreshold) {
          alpha = ex2_approx((m_ref - mx) * c);
          m_ref = mx;
          need = true;
        }
        // The P panel is double-buffered: P_i(U-2) V completed before S_i(U) did (issue order), so buffer U&1 is free.
        if (__any_sync(0xffffffffu, need)) {
          // P_i(U-1) V may still be accumulating into O_i: wait for it before rescaling (u >= 1 here).
          mbar_wait(&pv_done[i], (U - 1) & 1, 0x9450 | i);
          tc_fence_after();
#pragma unroll 1
          for (int cc = 0; cc < 4; ++cc) {
            uint32_t o[32];
            tmem_ld_x32(tO + cc * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 32; ++t) o[t] = __float_as_uint(__uint_as_float(o[t]) * alpha);
            tmem_st_x32(tO + cc * 32, o);
          }
          tmem_st_wait();
          float l_lo, l_hi;
          unpack_f32x2(lsum2, l_lo, l_hi);
          lsum2 = pack_f32x2(l_lo * alpha, l_hi * alpha);
          unpack_f32x2(lsum2b, l_lo, l_hi);
          lsum2b = pack_f32x2(l_lo * alpha, l_hi * alpha);
        }
        const float nmc = -m_ref * c;
        const uint64_t nmc2 = pack_f32x2(nmc, nmc);
        uint8_t* p_row = p_row0 + (U & 1) * P_BYTES;
        // Staged so that no instruction waits on its predecessor: (A) 32 independent packed scales, (B) 64 MUFU.EX2 back
        // to back, (C) row sums on 4 chains + bf16 packing + stores.
        uint64_t x2[32];
#pragma unroll
        for (int t = 0; t < 32; ++t)
          x2[t] = fma_f32x2(pack_f32x2(__uint_as_float(s[2 * t]), __uint_as_float(s[2 * t + 1])), c2, nmc2);
        float pe[64];
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          // the self-attention kernel's mix of MUFU and FMA-pipe exp2 — except in the windowed step: where a row's window
          // sits inside the 64-key step depends on the item's first window, i.e. on how the tokens are sharded, and the
          // sequence-parallel forward must stay bit-identical to the single-GPU one
          if constexpr (MASK == 2) exp2_pair<false>(x2[t], pe[2 * t], pe[2 * t + 1]);
          else SA_EXP2_PAIR(t, x2[t], pe[2 * t], pe[2 * t + 1]);
        }
        uint64_t la = lsum2, lb = lsum2b, lc = pack_f32x2(0.f, 0.f), ld = pack_f32x2(0.f, 0.f);
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          uint32_t pk[4];
#pragma unroll
          for (int t4 = 0; t4 < 4; ++t4) {
            const int t = c8 * 4 + t4;
            const uint64_t p2 = pack_f32x2(pe[2 * t], pe[2 * t + 1]);
            if (t4 == 0) la = add_f32x2(la, p2);
            else if (t4 == 1) lb = add_f32x2(lb, p2);
            else if (t4 == 2) lc = add_f32x2(lc, p2);
            else ld = add_f32x2(ld, p2);
            pk[t4] = pack_bf16x2(pe[2 * t], pe[2 * t + 1]);
          }
          *reinterpret_cast<uint4*>(p_row + ((c8 ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        lsum2 = add_f32x2(la, lc);
        lsum2b = add_f32x2(lb, ld);
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[i * 2 + (U & 1)]);
        ++U;
      };

      for (int si = 0; si < p.n_sets; ++si, ++N) {
        const int kv_len = p.kv_len[si];
        lsum2 = pack_f32x2(0.f, 0.f);
        lsum2b = pack_f32x2(0.f, 0.f);
        if (p.windowed[si]) {
          step(0, kv_len, std::integral_constant<int, 2>{});
        } else {
          const int ns = (kv_len + SUB - 1) / SUB;
          for (int u = 0; u < ns - 1; ++u) step(u, kv_len, std::integral_constant<int, 0>{});
          if (kv_len % SUB) step(ns - 1, kv_len, std::integral_constant<int, 1>{});
          else step(ns - 1, kv_len, std::integral_constant<int, 0>{});
        }
        // ---- epilogue of the set: O / l -> bf16 -> (+=) the staging tile. The tile is free (the previous item's TMA store
        // was retired before this item's Q load was issued) and, in accumulate mode, already holds the rows to add to.
        if (si == 0 && p.accumulate) mbar_wait(&res_full[i], it & 1, 0x9510 | i);
        mbar_wait(&o_final[i], N & 1, 0x9500 | i);
        tc_fence_after();
        // the item's last P V has completed: this tile's P buffers are idle until the next item's first softmax step, so
        // the next item's Q starts travelling now, under this epilogue and the output store
        if (si == p.n_sets - 1 && store_thread && w + (int)gridDim.x < p.n_items) load_q(w + gridDim.x);
        float l_lo, l_hi;
        lsum2 = add_f32x2(lsum2, lsum2b);
        unpack_f32x2(lsum2, l_lo, l_hi);
        const float inv_l = 1.0f / (l_lo + l_hi);
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          uint32_t o[32];
          tmem_ld_x32(tO + cc * 32, o);
          tmem_ld_wait();
          if (cc == 3) {   // all of O_i is in registers: the next set's first P V may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_free[i]);
          }
#pragma unroll
          for (int v8 = 0; v8 < 4; ++v8) {
            const int chunk = cc * 4 + v8;                       // 16-byte chunk of the 256-byte output row
            uint4* dst = reinterpret_cast<uint4*>(out_row + (chunk >> 3) * OUT_PANEL + (((chunk & 7) ^ (r & 7)) << 4));
            float y[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) y[t] = __uint_as_float(o[v8 * 8 + t]) * inv_l;
            if (si > 0 || p.accumulate) {
              const uint4 old = *dst;
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&old);
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float2 f = __bfloat1622float2(h[t]);
                y[2 * t] = f.x + bf16_round(y[2 * t]);
                y[2 * t + 1] = f.y + bf16_round(y[2 * t + 1]);
              }
            }
            uint4 uu;
            uu.x = pack_bf16x2(y[0], y[1]);
            uu.y = pack_bf16x2(y[2], y[3]);
            uu.z = pack_bf16x2(y[4], y[5]);
            uu.w = pack_bf16x2(y[6], y[7]);
            *dst = uu;
          }
        }
      }
      // ---- the item's output leaves as one TMA store per 64-column panel (rows >= q_len clipped by the tensor map)
      fence_proxy_async_smem();
      named_bar_sync(1 + i, 128);
      if (store_thread) {
        if (q0 + i * BQ < p.q_len) {
          tma_store_4d(&tmap_o, out_tile, 0, q0 + i * BQ, head, b);
          tma_store_4d(&tmap_o, out_tile + OUT_PANEL, 64, q0 + i * BQ, head, b);
          bulk_commit();
        }
        if (w + (int)gridDim.x < p.n_items) {
          bulk_wait_read0();            // the staging tile has been read: the next item's residual rows may land in it
          load_res(w + gridDim.x);
        }
      }
    }
    if (store_thread) bulk_wait0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace attnx
}  // namespace sa

extern "C" int sa_cross_attn3_d128(const sa_cross_attn_args* a, sa_stream_t stream_) {
  using namespace sa;
  using namespace sa::attnx;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !a->q || !a->out || a->n_sets < 1 || a->n_sets > MAX_SETS || a->batch <= 0 || a->heads <= 0 || a->q_len <= 0) {
    set_error("sa_cross_attn3_d128: bad argument");
    return SA_ERR_BAD_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(a->q) & 15) != 0 || (reinterpret_cast<uintptr_t>(a->out) & 15) != 0) {
    set_error("sa_cross_attn3_d128: q / out must be 16-byte aligned");
    return SA_ERR_BAD_ARG;
  }
  if (a->q_ls % 8 || a->q_bs % 8 || a->o_ls % 8 || a->o_bs % 8) {
    set_error("sa_cross_attn3_d128: strides must be multiples of 8 elements");
    return SA_ERR_BAD_ARG;
  }
  Params p;
  p.q = reinterpret_cast<const __nv_bfloat16*>(a->q);
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  p.q_bs = a->q_bs; p.q_ls = a->q_ls; p.o_bs = a->o_bs; p.o_ls = a->o_ls;
  p.q_len = a->q_len;
  p.heads = a->heads;
  p.n_pairs = (a->q_len + 2 * BQ - 1) / (2 * BQ);
  const long long items = (long long)p.n_pairs * a->heads * a->batch;
  if (items >= (1LL << 31)) { set_error("sa_cross_attn3_d128: too many work items"); return SA_ERR_UNSUPPORTED; }
  p.n_items = (int)items;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.accumulate = a->accumulate;
  p.n_sets = a->n_sets;
  p.rows_per_group = a->rows_per_group > 0 ? a->rows_per_group : 1 << 30;
  p.tok_offset = a->tok_offset;
  CUtensorMap tk[MAX_SETS], tv[MAX_SETS], to;
  int rc;
  for (int s = 0; s < MAX_SETS; ++s) {
    const int ss = s < a->n_sets ? s : 0;   // unused slots repeat set 0 (never dereferenced)
    const sa_cross_attn_set* cs = &a->set[ss];
    if (!cs->k || !cs->v || cs->kv_len <= 0 || cs->kv_total < cs->kv_len) { set_error("sa_cross_attn3_d128: bad key set %d", ss); return SA_ERR_BAD_ARG; }
    if (cs->windowed) {
      // the 256 rows of a work item touch at most 255 / rows_per_group + 2 consecutive windows; all their keys share one step
      const int span = 255 / p.rows_per_group + 2;
      if (a->rows_per_group <= 0 || span * cs->kv_len > SUB) {
        set_error("sa_cross_attn3_d128: windowed set needs (255 / rows_per_group + 2) * kv_len <= %d", SUB);
        return SA_ERR_UNSUPPORTED;
      }
    }
    p.kv_len[s] = cs->kv_len;
    p.windowed[s] = cs->windowed;
    auto mk = [&](CUtensorMap* m, const void* base, long long ls, long long bs) {
      uint64_t dims[4] = {(uint64_t)D, (uint64_t)cs->kv_total, (uint64_t)a->heads, (uint64_t)a->batch};
      uint64_t strides[3] = {(uint64_t)ls * 2, (uint64_t)D * 2, (uint64_t)bs * 2};
      uint32_t box[4] = {64, SUB, 1, 1};
      return make_tmap_bf16(m, base, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    };
    if ((rc = mk(&tk[s], cs->k, cs->k_ls, cs->k_bs))) return rc;
    if ((rc = mk(&tv[s], cs->v, cs->v_ls, cs->v_bs))) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)D, (uint64_t)a->q_len, (uint64_t)a->heads, (uint64_t)a->batch};
    uint64_t strides[3] = {(uint64_t)a->o_ls * 2, (uint64_t)D * 2, (uint64_t)a->o_bs * 2};
    uint32_t box[4] = {64, BQ, 1, 1};
    if ((rc = make_tmap_bf16(&to, a->out, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  CUtensorMap tq;
  {
    uint64_t dims[4] = {(uint64_t)D, (uint64_t)a->q_len, (uint64_t)a->heads, (uint64_t)a->batch};
    uint64_t strides[3] = {(uint64_t)a->q_ls * 2, (uint64_t)D * 2, (uint64_t)a->q_bs * 2};
    uint32_t box[4] = {64, BQ, 1, 1};
    if ((rc = make_tmap_bf16(&tq, a->q, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  if ((rc = ensure_dyn_smem(cross_attn_kernel, SMEM_BYTES, "cross_attn_kernel"))) return rc;
  const int grid = p.n_items < sm_count() ? p.n_items : sm_count();
  cross_attn_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tk[0], tv[0], tk[1], tv[1], tk[2], tv[2], to, tq, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "cross_attn_kernel launch");
  return SA_OK;
}
