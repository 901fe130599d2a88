// attn_self_tcgen05.cu — flash attention forward, head_dim 128, non-causal, for sm_100a. The one kernel behind
// sa_flash_attn_d128 (earlier generations live as text under profiles/experiments/ with what each one taught).
//
// Softmax and tensor pipe are decoupled:
//
//   * Q lives in TMEM (64 columns per 128-row tile, stored once by the softmax threads) and S = Q K^T is a TS-mode MMA
//     (A from TMEM) — which also halves the shared-memory operand traffic of the score MMAs;
//   * P goes through shared memory (double-buffered SWIZZLE_128B panel of 128 rows x 64 keys per Q tile) and O += P V is
//     an SS MMA;
//   * therefore S_i(u+1) is issued as soon as softmax warpgroup i has pulled S_i(u) into registers, and is complete
//     long before that warpgroup finishes step u: the softmax warpgroups never wait for the tensor pipe in steady
//     state and the tensor pipe only waits for P.
//
// One CTA = one (batch, head) x 256 query rows (two 128-row Q tiles), 64 keys per step, 4-stage TMA ring of
// {K, V} 64-key tiles. TMEM columns: Q0 0-63, Q1 64-127, S0 128-191, S1 192-255, O0 256-383, O1 384-511.
// Warps: 0-3 softmax tile 0, 4-7 softmax tile 1 (TMEM lane quarter = warp % 4), 8 = TMA producer, 9 / 10 = MMA issuer
// of Q tile 0 / 1 (a single issuer serving both tiles in a fixed order parks on the other tile's P: +4 % from the split).
// Softmax arithmetic: lazy rescale of O / l (only when the row max grew by more than 2^8), FMNMX3 row max, FFMA2 scale,
// MUFU.EX2 exponentials staged back to back (16/clk/SM: the XU pipe is the co-limiter, tools/micro/mufu_bench.cu).
//
// Replaces attention() of the reference (wan/models/wan_fantasy_transformer3d_1B.py:158-207, SDPA branch).
#include <stdlib.h>

#include <mutex>
#include <type_traits>

#include "../../include/stableavatar_b200.h"
#include "sa_host.h"
#include "sa_ptx.cuh"

namespace sa {
namespace attn8 {

constexpr int BQ = 128, SUB = 64, D = 128;
constexpr int STAGES = 4;
constexpr int KV_PANEL = SUB * 128;          // 64 rows x 128 B = 8 KB
constexpr int KV_TILE = 2 * KV_PANEL;        // [64 keys x 128 d] = 16 KB
constexpr int STAGE_BYTES = 2 * KV_TILE;     // K + V
constexpr int P_BYTES = BQ * 128;            // [128 rows x 64 keys] bf16 = 16 KB
constexpr int NUM_THREADS = 352;   // 8 softmax warps, TMA producer, one MMA-issuing warp per Q tile
constexpr int TMEM_COLS = 512;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 4 * P_BYTES + 256 + 1024;
constexpr float kRescaleThreshold = 8.0f;
constexpr int MAX_DST = 8;

struct Params {
  const __nv_bfloat16* q;
  __nv_bfloat16* dst_ptr[MAX_DST];   // destination k holds query rows [k * rows_per_dst, (k + 1) * rows_per_dst)
  long long q_bs, q_ls, dst_bs, dst_ls;
  int q_len, kv_len;
  float scale_log2;
  int accumulate;
  // TMA-store epilogue (accumulate == 0): query row r belongs to destination r / rows_per_dst (one tensor map each,
  // [B, rows_per_dst, heads, 128] views). One destination = the caller's `out`; several = the sequence-parallel O exchange,
  // where every destination is the o_recv buffer of the rank that owns those tokens (a peer mapping over NVLink).
  int n_dst, rows_per_dst, peer_dst;
};
struct OutMaps {
  CUtensorMap m[MAX_DST];
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

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__global__ void __launch_bounds__(NUM_THREADS, 1)
flash_attn_v8_kernel(const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                     const __grid_constant__ OutMaps tmap_o, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sKV = smem;                                   // [STAGES][K tile | V tile]
  uint8_t* sP = smem + STAGES * STAGE_BYTES;             // [2 q tiles][2 buffers][P_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * P_BYTES);
  uint64_t* kv_full = bars;                 // [STAGES]
  uint64_t* kv_empty = bars + STAGES;       // [STAGES]
  uint64_t* s_full = bars + 2 * STAGES;     // [2]  S_i(u) complete in TMEM
  uint64_t* s_cons = s_full + 2;            // [2]  softmax i has S_i(u) in registers
  uint64_t* p_full = s_cons + 2;            // [2 tiles][2 P buffers]  P_i(u) in shared memory (and O_i rescaled if needed).
                                            // One barrier per P buffer: a softmax warpgroup may finish steps u and u+1
                                            // before the MMA warp (held up by the other tile) consumes P_i(u); with a
                                            // single barrier those two completions would alias in the phase parity.
  uint64_t* pv_done = p_full + 4;           // [2]  one completion per PV_i(u)
  uint64_t* o_final = pv_done + 2;          // [2]
  uint64_t* q_ready = o_final + 2;          // [1]  Q tiles stored in TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_ready + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * BQ);
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int n_sub = (p.kv_len + SUB - 1) / SUB;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
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
      }
      for (int i = 0; i < 4; ++i) mbar_init(&p_full[i], 4);
      mbar_init(q_ready, 8);
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

  // Register budget: the three single-thread roles give theirs to the softmax warpgroups (two score tiles in flight).
  // (setmaxnreg sits inside each role branch: ptxas allocates every region after a join for the smallest budget.)
  if (warp == 8) {
    // ------------------------------------------------------------ TMA producer: {K, V} 64-key tiles
    if (elect_one()) {  // elect.sync: single active lane is known to ptxas -> no R2UR waterfall per UTCHMMA / UTMALDG
      for (int u = 0; u < n_sub; ++u) {
        const int s = u % STAGES;
        const uint32_t ph = (u / STAGES) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1, 0x4100 | s);
        mbar_arrive_expect_tx(&kv_full[s], STAGE_BYTES);
        uint8_t* st = sKV + s * STAGE_BYTES;
        tma_load_4d(st, &tmap_k, &kv_full[s], 0, u * SUB, head, b);
        tma_load_4d(st + KV_PANEL, &tmap_k, &kv_full[s], 64, u * SUB, head, b);
        tma_load_4d(st + KV_TILE, &tmap_v, &kv_full[s], 0, u * SUB, head, b);
        tma_load_4d(st + KV_TILE + KV_PANEL, &tmap_v, &kv_full[s], 64, u * SUB, head, b);
      }
    }
  } else if (warp == 9 || warp == 10) {
    // ------------------------------------------------------------ MMA issuers: one warp per Q tile, so that the next
    // score tile of a Q tile is issued the moment its softmax warpgroup has consumed the current one, whatever the
    // other tile is doing (a single issuer serving both tiles in a fixed order parks on the other tile's P).
    if (elect_one()) {
      const int i = warp - 9;
      constexpr uint32_t idesc_qk = umma_idesc_bf16(BQ, SUB, 0, 0);  // A = Q (TMEM), B = 64 keys of K (K-major)
      constexpr uint32_t idesc_pv = umma_idesc_bf16(BQ, D, 0, 1);    // A = P (smem, K-major), B = V (MN-major)
      // This thread's serial path sits between "P_i(u) is ready" and "S_i(u+2) is issued": everything that does not depend
      // on a barrier is prepared before the wait. Shared-memory descriptors are kept as a 32-bit low word (address >> 4 |
      // LBO) that advances by constants and one constant high word (SBO 1024, version 1, SWIZZLE_128B).
      constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (kSwz128 << 29);
      const uint32_t k_lo0 = ((smem_u32(sKV) & 0x3ffff) >> 4) | (1u << 16);                                    // LBO 16
      const uint32_t v_lo0 = ((smem_u32(sKV + KV_TILE) & 0x3ffff) >> 4) | (uint32_t(KV_PANEL >> 4) << 16);     // LBO = panel
      const uint32_t p_lo0 = ((smem_u32(sP + i * 2 * P_BYTES) & 0x3ffff) >> 4) | (1u << 16);
      const uint32_t tS = tmem_base + 128 + i * SUB, tQ = tmem_base + i * 64, tO = tmem_base + 256 + i * 128;
      auto issue_S = [&](uint32_t k_lo) {
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma_ts_lh(tS, tQ + k * 8, k_lo + (k >> 2) * (KV_PANEL >> 4) + (k & 3) * 2, DESC_HI, idesc_qk, k != 0);
        umma_commit(&s_full[i]);
      };
      auto issue_PV = [&](uint32_t p_lo, uint32_t v_lo, uint32_t acc0) {
#pragma unroll
        for (int k = 0; k < SUB / 16; ++k)
          umma_ss_lh(tO, p_lo + k * 2, DESC_HI, v_lo + k * (2048 >> 4), DESC_HI, idesc_pv, k == 0 ? acc0 : 1u);
        umma_commit(&pv_done[i]);
      };
      mbar_wait(q_ready, 0, 0x8300);
      mbar_wait(&kv_full[0], 0, 0x8310);
      tc_fence_after();
      issue_S(k_lo0);
      int ws = 1, wph = 0;          // ring stage / phase of the next score tile to issue
      auto ahead = [&](int w) {   // issue S_i(w) once K(w) has landed and S_i(w - 1) is in the softmax registers
        uint32_t k_lo = k_lo0 + ws * (STAGE_BYTES >> 4);
        asm volatile("" : "+r"(k_lo));
        mbar_wait(&kv_full[ws], wph, 0x8320 | ws);
        mbar_wait(&s_cons[i], (w - 1) & 1, 0x8330 | i);
        tc_fence_after();
        issue_S(k_lo);
        if (++ws == STAGES) { ws = 0; wph ^= 1; }
      };
      int us = 0;
      for (int u = 0; u < n_sub; ++u) {
        if (u + 1 < n_sub) ahead(u + 1);
        uint32_t p_lo = p_lo0 + (u & 1) * (P_BYTES >> 4), v_lo = v_lo0 + us * (STAGE_BYTES >> 4);
        asm volatile("" : "+r"(p_lo), "+r"(v_lo));
        mbar_wait(&p_full[i * 2 + (u & 1)], (u >> 1) & 1, 0x8340 | (i * 2 + (u & 1)));
        tc_fence_after();
        issue_PV(p_lo, v_lo, u > 0 ? 1u : 0u);
        if (u == n_sub - 1) umma_commit(&o_final[i]);
        umma_commit(&kv_empty[us]);   // this tile is done with K(u), V(u)
        if (++us == STAGES) us = 0;
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
    const int row = q0 + i * BQ + r;
    const bool row_ok = row < p.q_len;
    uint8_t* p_row0 = sP + i * 2 * P_BYTES + r * 128;

    // ---- Q row -> TMEM (A operand layout: lane = row, column c holds elements 2c, 2c+1)
    {
      const uint4* qp = reinterpret_cast<const uint4*>(p.q + (long long)b * p.q_bs + (long long)(row_ok ? row : 0) * p.q_ls + head * D);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t w[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 v = row_ok ? qp[h * 8 + c] : make_uint4(0, 0, 0, 0);
          w[c * 4] = v.x; w[c * 4 + 1] = v.y; w[c * 4 + 2] = v.z; w[c * 4 + 3] = v.w;
        }
        tmem_st_x32(tQ + h * 32, w);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_ready);
    }

    const float c = p.scale_log2;
    const uint64_t c2 = pack_f32x2(c, c);
    float m_ref = 0.f;
    uint64_t lsum2 = pack_f32x2(0.f, 0.f), lsum2b = pack_f32x2(0.f, 0.f);  // two independent row-sum chains

    auto step = [&](int u, auto mask_tag) {
      constexpr bool MASK = decltype(mask_tag)::value;
      mbar_wait(&s_full[i], u & 1, 0x4400 | i);
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
      if constexpr (MASK) {
        const int valid = p.kv_len - u * SUB;
#pragma unroll
        for (int t = 0; t < SUB; ++t)
          if (t >= valid) s[t] = 0xff800000u;  // -inf
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
      // The P panel is double-buffered: P_i(u-2) V completed before S_i(u) did (issue order), so buffer u&1 is free.
      if (__any_sync(0xffffffffu, need)) {
        // P_i(u-1) V may still be accumulating into O_i: wait for it before rescaling (u >= 1 here).
        mbar_wait(&pv_done[i], (u - 1) & 1, 0x4450 | i);
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
      uint8_t* p_row = p_row0 + (u & 1) * P_BYTES;
      {
        // (A) 32 packed scales, (B) 64 MUFU.EX2, (C) row sums on 4 chains + bf16 packing + stores. ptxas turns (B) + (C)
        // into a steady "MUFU, MUFU, FADD2, F2FP" stream whose stall counts (8 + 1 + 1 + 6) make one warp alone issue
        // exactly at the XU rate (8 cycles per warp instruction); the two softmax warps of an SM sub-partition share that
        // pipe, so a step costs each of them >= 1024 cycles of MUFU time — as much as the step's MMAs take on the tensor
        // pipe — plus ~400 cycles of barrier waits, TMEM load, row max and fences (in-kernel event trace,
        // profiles/r02_attn_investigation.md). Three sixteenths of the exponentials therefore take the FMA / ALU pipes
        // (exp2_pair): +3.5 ... +5 % in 30-round randomised rotations at B = 3 (r02_attn_ab_emulation_30rounds*.log), while
        // a quarter or more loses it again (the extra ~10 instructions per pair cost issue slots and power). Round-2 variants that attacked
        // the rest and did NOT pay, measured the same way: TMEM load + row max of S(u+1) hidden under
        // the exponentials of step u (224 registers through setmaxnreg): -8 ... -17 %; a shared 128-column score buffer
        // with 128-key steps handed between the two Q tiles: +-2 %. Sources: profiles/experiments/.
        uint64_t x2[32];
#pragma unroll
        for (int t = 0; t < 32; ++t)
          x2[t] = fma_f32x2(pack_f32x2(__uint_as_float(s[2 * t]), __uint_as_float(s[2 * t + 1])), c2, nmc2);
        float pe[64];
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          SA_EXP2_PAIR(t, x2[t], pe[2 * t], pe[2 * t + 1]);
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
          sts128(smem_u32(p_row) + ((c8 ^ (r & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
        }
        lsum2 = add_f32x2(la, lc);
        lsum2b = add_f32x2(lb, ld);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[i * 2 + (u & 1)]);
    };

    const bool ragged = (p.kv_len % SUB) != 0;
    for (int u = 0; u < n_sub - 1; ++u) step(u, std::false_type{});
    if (ragged) step(n_sub - 1, std::true_type{});
    else step(n_sub - 1, std::false_type{});
  

    // ---- epilogue: O_i / l -> bf16 -> global
    mbar_wait(&o_final[i], 0, 0x4500 | i);
    tc_fence_after();
    float l_lo, l_hi;
    lsum2 = add_f32x2(lsum2, lsum2b);
    unpack_f32x2(lsum2, l_lo, l_hi);
    const float inv_l = 1.0f / (l_lo + l_hi);
    const int t0 = q0 + i * BQ;
    const int k0 = t0 / p.rows_per_dst;                        // destination of the tile's first row
    const bool straddle = k0 + 1 < p.n_dst && (k0 + 1) * p.rows_per_dst < t0 + BQ;
    if (!p.accumulate && !straddle) {
      // The tile's two P buffers are free (o_final covers the last P V): they become the staging tile — two
      // [128 rows x 64 cols] SWIZZLE_128B panels — and the tile leaves as TMA stores: coalesced 128-byte lines instead of
      // 16-byte fragments per thread, which is also what makes storing STRAIGHT INTO A PEER'S o_recv over NVLink efficient
      // (the sequence-parallel O exchange then has no kernel of its own). A tile that straddles two token owners takes the
      // per-row path below instead: a TMA store may run past the END of a tensor but must not start at a negative row
      // (cudaErrorIllegalInstruction on sm_100a, found on the first 2-GPU run), and a box cannot be shortened per launch.
      uint8_t* stage = sP + i * 2 * P_BYTES;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t o[32];
        tmem_ld_x32(tO + cc * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int v8 = 0; v8 < 4; ++v8) {
          const int chunk = cc * 4 + v8;                       // 16-byte chunk of the 256-byte output row
          uint4 uu;
          uu.x = pack_bf16x2(__uint_as_float(o[v8 * 8 + 0]) * inv_l, __uint_as_float(o[v8 * 8 + 1]) * inv_l);
          uu.y = pack_bf16x2(__uint_as_float(o[v8 * 8 + 2]) * inv_l, __uint_as_float(o[v8 * 8 + 3]) * inv_l);
          uu.z = pack_bf16x2(__uint_as_float(o[v8 * 8 + 4]) * inv_l, __uint_as_float(o[v8 * 8 + 5]) * inv_l);
          uu.w = pack_bf16x2(__uint_as_float(o[v8 * 8 + 6]) * inv_l, __uint_as_float(o[v8 * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(stage + (chunk >> 3) * P_BYTES + r * 128 + (((chunk & 7) ^ (r & 7)) << 4)) = uu;
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + i, 128);
      if (quarter == 0 && lane == 0 && t0 < p.q_len) {
        const int r0 = t0 - k0 * p.rows_per_dst;               // rows past the end of the destination are clipped
        tma_store_4d(&tmap_o.m[k0], stage, 0, r0, head, b);
        tma_store_4d(&tmap_o.m[k0], stage + P_BYTES, 64, r0, head, b);
        bulk_commit();
        if (p.peer_dst) bulk_wait0();        // peer writes are complete before the kernel (and the flag barrier after it) ends
        else bulk_wait_read0();              // the staging tile has been read; the writes are ordered by kernel completion
      }
    } else {
      // one row per thread, 16-byte stores: the read-modify-write (accumulate) form, and tiles that straddle two destinations
      const int kr = row_ok ? row / p.rows_per_dst : 0;
      __nv_bfloat16* orow = p.dst_ptr[kr] + (long long)b * p.dst_bs + (long long)(row - kr * p.rows_per_dst) * p.dst_ls + head * D;
      const bool rmw = p.accumulate != 0;
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
            if (rmw) {
              const uint4 old = *dst;
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&old);
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                float2 f = __bfloat1622float2(h[t]);
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
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace attn8
}  // namespace sa

// dst == nullptr: ordinary call, the output goes to a->out. Otherwise the sequence-parallel O exchange: query row r is
// stored to dst[r / rows_per_dst] at [b, r % rows_per_dst, head, :] (element strides dst_bs / dst_ls, heads 128 apart).
static int launch_flash_attn(const sa_attn_args* a, void* const* dst, int n_dst, int rows_per_dst, long long dst_bs,
                             long long dst_ls, cudaStream_t stream) {
  using namespace sa;
  using namespace sa::attn8;
  if (!a || !a->q || !a->k || !a->v || (!dst && !a->out)) { set_error("sa_flash_attn_d128: null pointer"); return SA_ERR_BAD_ARG; }
  if (a->batch <= 0 || a->heads <= 0 || a->q_len <= 0 || a->kv_len <= 0) {
    set_error("sa_flash_attn_d128: non-positive dims");
    return SA_ERR_BAD_ARG;
  }
  if (a->q_ls % 8 || a->k_ls % 8 || a->v_ls % 8 || a->q_bs % 8 || a->k_bs % 8 || a->v_bs % 8 ||
      (!dst && (a->o_ls % 8 || a->o_bs % 8)) || (dst && (dst_ls % 8 || dst_bs % 8))) {
    set_error("sa_flash_attn_d128: strides must be multiples of 8 elements");
    return SA_ERR_BAD_ARG;
  }
  if (dst && (a->accumulate || n_dst < 1 || n_dst > MAX_DST || rows_per_dst < 1 || (long long)n_dst * rows_per_dst < a->q_len)) {
    set_error("sa_flash_attn_d128_sp: 1 <= n_dst <= %d, n_dst * rows_per_dst >= q_len, no accumulate", MAX_DST);
    return SA_ERR_BAD_ARG;
  }
  CUtensorMap tk, tv;
  OutMaps to;
  auto mk = [&](CUtensorMap* m, const void* base, int len, long long ls, long long bs, int box_rows) {
    uint64_t dims[4] = {(uint64_t)D, (uint64_t)len, (uint64_t)a->heads, (uint64_t)a->batch};
    uint64_t strides[3] = {(uint64_t)ls * 2, (uint64_t)D * 2, (uint64_t)bs * 2};
    uint32_t box[4] = {64, (uint32_t)box_rows, 1, 1};
    return make_tmap_bf16(m, base, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  };
  int rc;
  if ((rc = mk(&tk, a->k, a->kv_len, a->k_ls, a->k_bs, SUB))) return rc;
  if ((rc = mk(&tv, a->v, a->kv_len, a->v_ls, a->v_bs, SUB))) return rc;
  if ((reinterpret_cast<uintptr_t>(a->q) & 15) != 0) { set_error("sa_flash_attn_d128: q must be 16-byte aligned"); return SA_ERR_BAD_ARG; }
  Params p;
  p.q = reinterpret_cast<const __nv_bfloat16*>(a->q);
  for (int k = 0; k < MAX_DST; ++k) p.dst_ptr[k] = reinterpret_cast<__nv_bfloat16*>(dst ? dst[k < n_dst ? k : 0] : a->out);
  p.q_bs = a->q_bs; p.q_ls = a->q_ls;
  p.dst_bs = dst ? dst_bs : a->o_bs; p.dst_ls = dst ? dst_ls : a->o_ls;
  p.q_len = a->q_len; p.kv_len = a->kv_len;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.accumulate = a->accumulate;
  p.n_dst = 1; p.rows_per_dst = a->q_len; p.peer_dst = 0;
  if (dst) {
    p.n_dst = n_dst; p.rows_per_dst = rows_per_dst; p.peer_dst = 1;
    for (int k = 0; k < n_dst; ++k) {
      if (!dst[k]) { set_error("sa_flash_attn_d128_sp: null destination %d", k); return SA_ERR_BAD_ARG; }
      if ((rc = mk(&to.m[k], dst[k], rows_per_dst, dst_ls, dst_bs, BQ))) return rc;
    }
    for (int k = n_dst; k < MAX_DST; ++k) to.m[k] = to.m[0];
  } else {
    if (!a->accumulate && (rc = mk(&to.m[0], a->out, a->q_len, a->o_ls, a->o_bs, BQ))) return rc;
    if (a->accumulate) to.m[0] = tk;           // never dereferenced on the read-modify-write path
    for (int k = 1; k < MAX_DST; ++k) to.m[k] = to.m[0];
  }
  dim3 grid((a->q_len + 2 * BQ - 1) / (2 * BQ), a->heads, a->batch);
  if ((rc = ensure_dyn_smem(flash_attn_v8_kernel, SMEM_BYTES, "flash_attn_v8_kernel"))) return rc;
  flash_attn_v8_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tk, tv, to, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "flash_attn_v8_kernel launch");
  return SA_OK;
}

extern "C" int sa_flash_attn_d128(const sa_attn_args* a, sa_stream_t stream) {
  return launch_flash_attn(a, nullptr, 0, 0, 0, 0, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int sa_flash_attn_d128_sp(const sa_attn_args* a, void* const* dst, int32_t n_dst, int32_t rows_per_dst,
                                     int64_t dst_bs, int64_t dst_ls, sa_stream_t stream) {
  if (!dst) { sa::set_error("sa_flash_attn_d128_sp: null destination table"); return sa::SA_ERR_BAD_ARG; }
  return launch_flash_attn(a, dst, n_dst, rows_per_dst, dst_bs, dst_ls, reinterpret_cast<cudaStream_t>(stream));
}
