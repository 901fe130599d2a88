// gemm_tcgen05.cu — persistent, warp-specialised bf16 GEMM for sm_100a with fused epilogues.
//
//   out[M,N] = epilogue(A[M,K] . W[N,K]^T)          (both operands K-major == nn.Linear layout)
//
// One CTA per SM, 320 threads: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (+TMEM owner), warps 2-9 = epilogue
// (two warps per TMEM lane quarter, each owning half of the tile's columns). Tile 128 x 256 x 64, 4 smem stages (48 KB
// each, SWIZZLE_128B), fp32 accumulators double-buffered in TMEM (2 x 256 columns) so tile i's epilogue overlaps tile
// i+1's MMAs. The epilogue is specialised at compile time on (activation, residual mode): bf16 results are staged in
// shared memory in the 128-byte swizzle pattern and written with TMA stores (coalesced, M/N tails clipped by the
// hardware); a generic run-time-switched epilogue with direct stores covers fp32 outputs and odd shapes.
// M/N/K tails on the load side are TMA zero-fill.
//
// Replaces the nn.Linear calls of the reference (see include/stableavatar_b200.h for the file:line list).
#include <stdio.h>

#include "../../include/stableavatar_b200.h"
#include "sa_host.h"
#include "sa_ptx.cuh"

namespace sa {
namespace gemm {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_BYTES = BN * BK * 2;   // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = (2 + EPI_WARPS) * 32;
constexpr int TMEM_COLS = 512;
constexpr int PANEL_BYTES = 32 * 128;  // one epilogue warp's staging panel: 32 rows x 64 bf16 columns, SWIZZLE_128B
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_WARPS * PANEL_BYTES + 256 /*barriers*/ + 1024 /*alignment slack*/;

struct Params {
  void* out;
  const void* bias;
  const void* res;
  const void* gate;
  long long ldc, ldr, gate_ld;
  int M, N, K;
  int bias_dtype, out_dtype, res_dtype, act, res_mode, round_y, rows_per_batch;
  int tiles_m, tiles_n;
};

__device__ __forceinline__ void load8(const void* base, int dtype, long long idx, float (&v)[8]) {
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
__device__ __forceinline__ void load8_bf16(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ float load1(const void* base, int dtype, long long idx) {
  return dtype == SA_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx])
                          : reinterpret_cast<const float*>(base)[idx];
}
__device__ __forceinline__ void store8(void* base, int dtype, long long idx, const float (&v)[8]) {
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
__device__ __forceinline__ void store1(void* base, int dtype, long long idx, float v) {
  if (dtype == SA_BF16) reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(base)[idx] = v;
}

template <int ACT>
__device__ __forceinline__ float act_fn(float y) {
  if constexpr (ACT == 1) {  // GELU(tanh): 0.5 y (1 + tanh(sqrt(2/pi) (y + 0.044715 y^3)))
    const float u = y * fmaf(0.7978845608028654f * 0.044715f, y * y, 0.7978845608028654f);
    const float hy = 0.5f * y;
    return fmaf(hy, tanh_approx(u), hy);
  } else if constexpr (ACT == 2) {  // SiLU
    return __fdividef(y, 1.0f + __expf(-y));
  } else if constexpr (ACT == 3) {  // GELU(erf)
    return 0.5f * y * (1.0f + erff(y * 0.7071067811865476f));
  } else {
    return y;
  }
}
__device__ __forceinline__ float act_rt(float y, int act, bool exact) {
  if (exact && act == 1) {  // fp32 mode (round_y == 0): libm tanhf instead of tanh.approx (2^-11 relative error)
    const float u = y * fmaf(0.7978845608028654f * 0.044715f, y * y, 0.7978845608028654f);
    return 0.5f * y * (1.0f + tanhf(u));
  }
  if (exact && act == 2) return y / (1.0f + expf(-y));
  return act == 1 ? act_fn<1>(y) : (act == 2 ? act_fn<2>(y) : (act == 3 ? act_fn<3>(y) : y));
}

// TMA store of one staged panel (global <- shared), bulk-group completion.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// FAST: bf16 output through swizzled smem staging + TMA store, bf16 bias/res, bf16 rounding of the Linear output
// (autocast), ACT / RES fixed at compile time. !FAST: everything decided at run time, direct global stores.
template <bool FAST, int ACT, int RES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_c, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_out = smem + STAGES * STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(stage_out + EPI_WARPS * PANEL_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;  // accumulator stage ready for the epilogue
  uint64_t* tempty = tfull + 2;      // accumulator stage drained by the epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (FAST) tma_prefetch_desc(&tmap_c);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tfull[s], 1);
        mbar_init(&tempty[s], EPI_WARPS);
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

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {  // elect.sync (not lane == 0): ptxas then knows a single lane is active and feeds the uniform
                        // operands of UTMALDG / UTCHMMA with plain R2UR instead of a per-instruction waterfall loop
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.tiles_n) * BM;
        const int n0 = (tile % p.tiles_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1, 0x0100 | s);
          mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
          uint8_t* sa_ = smem + s * STAGE_BYTES;
          tma_load_2d(sa_, &tmap_a, &full[s], kb * BK, m0);
          tma_load_2d(sa_ + A_BYTES, &tmap_b, &full[s], kb * BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
      uint32_t it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t as = tcount & 1;
        mbar_wait(&tempty[as], ((tcount >> 1) & 1) ^ 1, 0x0200 | as);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full[s], ph, 0x0300 | s);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
          const uint64_t adesc = umma_smem_desc(a_addr, 16, 1024, kSwz128);
          const uint64_t bdesc = umma_smem_desc(a_addr + A_BYTES, 16, 1024, kSwz128);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advancing 16 bf16 (32 B) along K inside the 128-byte swizzle atom = +2 in the (>>4) address field
            umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty[s]);  // frees the smem stage once these MMAs have read it
        }
        umma_commit(&tfull[as]);  // accumulator complete
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: warps 2..9
    const int ew = warp - 2;
    const int q = warp & 3;        // TMEM lane quarter this warp may read
    const int half = ew >> 2;      // which 128 of the tile's 256 columns
    uint8_t* panel = stage_out + ew * PANEL_BYTES;
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
      const int m0 = (tile / p.tiles_n) * BM;
      const int n0 = (tile % p.tiles_n) * BN;
      const uint32_t as = tcount & 1;
      mbar_wait(&tfull[as], (tcount >> 1) & 1, 0x0400 | as);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + as * BN + half * 128;

      if constexpr (FAST) {
        const __nv_bfloat16* bias = reinterpret_cast<const __nv_bfloat16*>(p.bias);
        const __nv_bfloat16* res_row =
            RES ? reinterpret_cast<const __nv_bfloat16*>(p.res) + (long long)(row_ok ? row : 0) * p.ldr : nullptr;
        const __nv_bfloat16* gate_row =
            RES == 2 ? reinterpret_cast<const __nv_bfloat16*>(p.gate) + (long long)((row_ok ? row : 0) / p.rows_per_batch) * p.gate_ld
                     : nullptr;
#pragma unroll 1
        for (int pn = 0; pn < 2; ++pn) {  // two 64-column panels
          const int nc = n0 + half * 128 + pn * 64;
          if (nc >= p.N) break;  // warp-uniform
          uint32_t r0[32], r1[32];
          tmem_ld_x32(t_row + pn * 64, r0);
          tmem_ld_x32(t_row + pn * 64 + 32, r1);
          tmem_ld_wait();
          if (lane == 0) bulk_wait_read0();  // previous TMA store has finished reading this warp's panel
          __syncwarp();
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8) {  // 8 chunks of 8 columns = 16 bytes of bf16 each
            const int n = nc + c8 * 8;
            float y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = __uint_as_float(c8 < 4 ? r0[c8 * 8 + i] : r1[(c8 - 4) * 8 + i]);
            const bool col_ok = n < p.N;  // N % 8 == 0 on this path
            if (bias != nullptr && col_ok) {
              float b[8];
              load8_bf16(bias + n, b);
#pragma unroll
              for (int i = 0; i < 8; ++i) y[i] += b[i];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = bf16_round(y[i]);
            if constexpr (ACT != 0) {
#pragma unroll
              for (int i = 0; i < 8; ++i) y[i] = act_fn<ACT>(y[i]);
            }
            if constexpr (RES != 0) {
              if (row_ok && col_ok) {
                float rs[8];
                load8_bf16(res_row + n, rs);
                if constexpr (RES == 2) {
                  float g[8];
                  load8_bf16(gate_row + n, g);
#pragma unroll
                  for (int i = 0; i < 8; ++i) y[i] = rs[i] + bf16_round(bf16_round(y[i]) * g[i]);
                } else {
#pragma unroll
                  for (int i = 0; i < 8; ++i) y[i] = rs[i] + bf16_round(y[i]);
                }
              }
            }
            uint4 u;
            u.x = pack_bf16x2(y[0], y[1]);
            u.y = pack_bf16x2(y[2], y[3]);
            u.z = pack_bf16x2(y[4], y[5]);
            u.w = pack_bf16x2(y[6], y[7]);
            *reinterpret_cast<uint4*>(panel + lane * 128 + ((c8 ^ (lane & 7)) << 4)) = u;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmap_c, panel, nc, m0 + q * 32);
            bulk_commit();
          }
        }
      } else {
        const long long gate_row = (p.res_mode == 2 && row_ok) ? (long long)(row / p.rows_per_batch) * p.gate_ld : 0;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int nc = n0 + half * 128 + c * 32;
          if (nc >= p.N) break;  // warp-uniform
          uint32_t r[32];
          tmem_ld_x32(t_row + c * 32, r);
          tmem_ld_wait();
          if (!row_ok) continue;
#pragma unroll 1
          for (int j = 0; j < 32; ++j) {
            const int n = nc + j;
            if (n >= p.N) break;
            float y = __uint_as_float(r[j]);
            if (p.bias) y += load1(p.bias, p.bias_dtype, n);
            if (p.res_mode == 3) y += load1(p.res, p.res_dtype, (long long)row * p.ldr + n);  // K-chunked accumulation
            if (p.round_y) y = bf16_round(y);
            if (p.act) {
              y = act_rt(y, p.act, !p.round_y);
              if (p.round_y) y = bf16_round(y);
            }
            if (p.res_mode == 1 || p.res_mode == 2) {
              const float rs = load1(p.res, p.res_dtype, (long long)row * p.ldr + n);
              if (p.res_mode == 2) {
                float t = y * load1(p.gate, SA_BF16, gate_row + n);
                if (p.round_y) t = bf16_round(t);
                y = rs + t;
              } else {
                y = rs + y;
              }
            }
            store1(p.out, p.out_dtype, (long long)row * p.ldc + n, y);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
    }
    if (FAST && lane == 0) bulk_wait0();  // all of this warp's TMA stores have landed before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <bool FAST, int ACT, int RES>
static int launch(const CUtensorMap& tma, const CUtensorMap& tmb, const CUtensorMap& tmc, const Params& p, int grid,
                  cudaStream_t stream) {
  auto* kern = gemm_bf16_kernel<FAST, ACT, RES>;
  if (int rc = ensure_dyn_smem(kern, SMEM_BYTES, "gemm_bf16_kernel")) return rc;
  kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tma, tmb, tmc, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "gemm_bf16_kernel launch");
  return SA_OK;
}

}  // namespace gemm
}  // namespace sa

extern "C" int sa_gemm_bf16(const sa_gemm_args* a, sa_stream_t stream_) {
  using namespace sa;
  using namespace sa::gemm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!a || !a->a || !a->w || !a->out) { set_error("sa_gemm_bf16: null pointer"); return SA_ERR_BAD_ARG; }
  if (a->M <= 0 || a->N <= 0 || a->K <= 0) { set_error("sa_gemm_bf16: non-positive dims"); return SA_ERR_BAD_ARG; }
  if (a->K % 8 || a->lda % 8 || a->ldw % 8 || a->lda < a->K || a->ldw < a->K) {
    set_error("sa_gemm_bf16: K, lda, ldw must be multiples of 8 and ld >= K (K=%d lda=%lld ldw=%lld)", a->K,
              (long long)a->lda, (long long)a->ldw);
    return SA_ERR_BAD_ARG;
  }
  if (a->res_mode && !a->res) { set_error("sa_gemm_bf16: res_mode set but res is null"); return SA_ERR_BAD_ARG; }
  if (a->res_mode == 2 && (!a->gate || a->rows_per_batch <= 0)) {
    set_error("sa_gemm_bf16: gated residual needs gate and rows_per_batch > 0");
    return SA_ERR_BAD_ARG;
  }
  if (a->act < 0 || a->act > 3 || a->res_mode < 0 || a->res_mode > 3) {
    set_error("sa_gemm_bf16: bad act/res_mode");
    return SA_ERR_BAD_ARG;
  }

  CUtensorMap tma, tmb, tmc;
  {
    uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->M};
    uint64_t strides[1] = {(uint64_t)a->lda * 2};
    uint32_t box[2] = {BK, BM};
    int rc = make_tmap_bf16(&tma, a->a, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->N};
    uint64_t strides[1] = {(uint64_t)a->ldw * 2};
    uint32_t box[2] = {BK, BN};
    int rc = make_tmap_bf16(&tmb, a->w, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  // Fast path: bf16 everywhere, autocast rounding, 16-byte aligned rows.
  const bool fast = a->out_dtype == SA_BF16 && a->round_y && a->res_mode != 3 && a->N % 8 == 0 && a->ldc % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(a->out) & 15) == 0 && (!a->bias || a->bias_dtype == SA_BF16) &&
                    (!a->bias || (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0) &&
                    (a->res_mode == 0 || (a->res_dtype == SA_BF16 && a->ldr % 8 == 0 &&
                                          (reinterpret_cast<uintptr_t>(a->res) & 15) == 0)) &&
                    (a->res_mode != 2 || (a->gate_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(a->gate) & 15) == 0)) &&
                    (a->act == 0 || a->act == 1 || (a->act == 3 && a->res_mode == 0));
  if (fast) {
    uint64_t dims[2] = {(uint64_t)a->N, (uint64_t)a->M};
    uint64_t strides[1] = {(uint64_t)a->ldc * 2};
    uint32_t box[2] = {64, 32};
    int rc = make_tmap_bf16(&tmc, a->out, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  } else {
    tmc = tma;
    if (a->N % 8 == 0 && (a->ldc % 4 || (a->res_mode && a->ldr % 4))) {
      set_error("sa_gemm_bf16: ldc/ldr must be multiples of 4");
      return SA_ERR_BAD_ARG;
    }
  }
  Params p;
  p.out = a->out; p.bias = a->bias; p.res = a->res; p.gate = a->gate;
  p.ldc = a->ldc; p.ldr = a->ldr; p.gate_ld = a->gate_ld;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.bias_dtype = a->bias_dtype; p.out_dtype = a->out_dtype; p.res_dtype = a->res_dtype;
  p.act = a->act; p.res_mode = a->res_mode; p.round_y = a->round_y;
  p.rows_per_batch = a->rows_per_batch > 0 ? a->rows_per_batch : a->M;
  p.tiles_m = (a->M + BM - 1) / BM;
  p.tiles_n = (a->N + BN - 1) / BN;
  const int tiles = p.tiles_m * p.tiles_n;
  const int grid = tiles < sm_count() ? tiles : sm_count();

  if (!fast) return launch<false, 0, 0>(tma, tmb, tmc, p, grid, stream);
  if (a->act == 3) return launch<true, 3, 0>(tma, tmb, tmc, p, grid, stream);
  if (a->act == 1) {
    if (a->res_mode == 0) return launch<true, 1, 0>(tma, tmb, tmc, p, grid, stream);
    if (a->res_mode == 1) return launch<true, 1, 1>(tma, tmb, tmc, p, grid, stream);
    return launch<true, 1, 2>(tma, tmb, tmc, p, grid, stream);
  }
  if (a->res_mode == 0) return launch<true, 0, 0>(tma, tmb, tmc, p, grid, stream);
  if (a->res_mode == 1) return launch<true, 0, 1>(tma, tmb, tmc, p, grid, stream);
  return launch<true, 0, 2>(tma, tmb, tmc, p, grid, stream);
}
