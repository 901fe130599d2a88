// gemm_tcgen05.cu — persistent, warp-specialised bf16 GEMM for sm_100a with fused epilogues.
//
//   out[M,N] = epilogue(A[M,K] . W[N,K]^T)          (both operands K-major == nn.Linear layout)
//
// One CTA per SM, 192 threads: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (+TMEM owner), warps 2-5 = epilogue.
// Tile 128 x 256 x 64, 4 smem stages (48 KB each, SWIZZLE_128B), fp32 accumulators double-buffered in TMEM
// (2 x 256 columns) so tile i's epilogue overlaps tile i+1's MMAs. M/N/K tails are handled by TMA zero-fill on
// the load side and by masking on the store side.
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
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 /*barriers*/ + 1024 /*alignment slack*/;

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

__device__ __forceinline__ float apply_act(float y, int act) {
  if (act == 1) {  // GELU(tanh): 0.5 y (1 + tanh(sqrt(2/pi) (y + 0.044715 y^3)))
    float u = 0.7978845608028654f * (y + 0.044715f * y * y * y);
    return 0.5f * y * (1.0f + tanh_approx(u));
  } else if (act == 2) {  // SiLU
    return y / (1.0f + __expf(-y));
  } else if (act == 3) {  // GELU(erf)
    return 0.5f * y * (1.0f + erff(y * 0.7071067811865476f));
  }
  return y;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
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
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
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

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
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
    if (lane == 0) {
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
    // ------------------------------------------------------------ epilogue (warps 2..5, TMEM lane quarter = warp % 4)
    const int q = warp & 3;
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
      const int m0 = (tile / p.tiles_n) * BM;
      const int n0 = (tile % p.tiles_n) * BN;
      const uint32_t as = tcount & 1;
      mbar_wait(&tfull[as], (tcount >> 1) & 1, 0x0400 | as);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      const long long gate_row = (p.res_mode == 2 && row_ok) ? (long long)(row / p.rows_per_batch) * p.gate_ld : 0;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int nc = n0 + c * 32;
        if (nc >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld_x32(tmem_base + (uint32_t(q * 32) << 16) + as * BN + c * 32, r);
        tmem_ld_wait();
        if (!row_ok) continue;
        if (nc + 32 <= p.N) {
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            const int n = nc + j8 * 8;
            float y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = __uint_as_float(r[j8 * 8 + i]);
            if (p.bias) {
              float b[8];
              load8(p.bias, p.bias_dtype, n, b);
#pragma unroll
              for (int i = 0; i < 8; ++i) y[i] += b[i];
            }
            if (p.round_y) {
#pragma unroll
              for (int i = 0; i < 8; ++i) y[i] = bf16_round(y[i]);
            }
            if (p.act) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                y[i] = apply_act(y[i], p.act);
                if (p.round_y) y[i] = bf16_round(y[i]);
              }
            }
            if (p.res_mode) {
              float rs[8];
              load8(p.res, p.res_dtype, (long long)row * p.ldr + n, rs);
              if (p.res_mode == 2) {
                float g[8];
                load8(p.gate, SA_BF16, gate_row + n, g);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  float t = y[i] * g[i];
                  if (p.round_y) t = bf16_round(t);
                  y[i] = rs[i] + t;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = rs[i] + y[i];
              }
            }
            store8(p.out, p.out_dtype, (long long)row * p.ldc + n, y);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = nc + j;
            if (n >= p.N) continue;
            float y = __uint_as_float(r[j]);
            if (p.bias) y += load1(p.bias, p.bias_dtype, n);
            if (p.round_y) y = bf16_round(y);
            if (p.act) {
              y = apply_act(y, p.act);
              if (p.round_y) y = bf16_round(y);
            }
            if (p.res_mode) {
              float rs = load1(p.res, p.res_dtype, (long long)row * p.ldr + n);
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
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
  if (a->N % 8 == 0) {
    if (a->ldc % 8 || (a->res_mode && a->ldr % 8) || (a->res_mode == 2 && a->gate_ld % 8)) {
      set_error("sa_gemm_bf16: ldc/ldr/gate_ld must be multiples of 8");
      return SA_ERR_BAD_ARG;
    }
  }
  if (a->res_mode && !a->res) { set_error("sa_gemm_bf16: res_mode set but res is null"); return SA_ERR_BAD_ARG; }
  if (a->res_mode == 2 && (!a->gate || a->rows_per_batch <= 0)) {
    set_error("sa_gemm_bf16: gated residual needs gate and rows_per_batch > 0");
    return SA_ERR_BAD_ARG;
  }
  if (a->act < 0 || a->act > 3 || a->res_mode < 0 || a->res_mode > 2) {
    set_error("sa_gemm_bf16: bad act/res_mode");
    return SA_ERR_BAD_ARG;
  }

  CUtensorMap tma, tmb;
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
  Params p;
  p.out = a->out; p.bias = a->bias; p.res = a->res; p.gate = a->gate;
  p.ldc = a->ldc; p.ldr = a->ldr; p.gate_ld = a->gate_ld;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.bias_dtype = a->bias_dtype; p.out_dtype = a->out_dtype; p.res_dtype = a->res_dtype;
  p.act = a->act; p.res_mode = a->res_mode; p.round_y = a->round_y;
  p.rows_per_batch = a->rows_per_batch > 0 ? a->rows_per_batch : a->M;
  p.tiles_m = (a->M + BM - 1) / BM;
  p.tiles_n = (a->N + BN - 1) / BN;

  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(gemm_bf16_kernel)");
    attr_set = true;
  }
  const int tiles = p.tiles_m * p.tiles_n;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  gemm_bf16_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tma, tmb, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "gemm_bf16_kernel launch");
  return SA_OK;
}
