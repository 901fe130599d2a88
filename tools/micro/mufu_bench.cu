// Microbenchmark: per-SM throughput of the candidate softmax building blocks on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float a0 = seed + threadIdx.x * 1e-3f, a1 = a0 * 0.5f, a2 = a0 * 0.25f, a3 = a0 * 0.125f;
  uint32_t b0 = __float_as_uint(a0) & 0x3fff3fff, b1 = b0 ^ 0x01010101, b2 = b0 ^ 0x02020202, b3 = b0 ^ 0x03030303;
  unsigned long long c0 = ((unsigned long long)__float_as_uint(a0) << 32) | __float_as_uint(a1), c1 = c0 + 7, c2 = c0 + 9, c3 = c0 + 11;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
    } else if (MODE == 1) {
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(b0)); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(b1));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(b2)); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(b3));
    } else if (MODE == 2) {
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(b0)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(b1));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(b2)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(b3));
    } else if (MODE == 3) {
      asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a0)); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a1));
      asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a2)); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a3));
    } else if (MODE == 4) {
      asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(c0)); asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(c1));
      asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(c2)); asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(c3));
    } else if (MODE == 5) {
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(b0) : "f"(a0), "f"(a1)); a0 += __uint_as_float(b0 << 16);
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(b1) : "f"(a2), "f"(a3)); a2 += __uint_as_float(b1 << 16);
    } else if (MODE == 6) {
      asm volatile("max.f32 %0, %0, %1;" : "+f"(a0) : "f"(a1)); asm volatile("max.f32 %0, %0, %1;" : "+f"(a1) : "f"(a2));
      asm volatile("max.f32 %0, %0, %1;" : "+f"(a2) : "f"(a3)); asm volatile("max.f32 %0, %0, %1;" : "+f"(a3) : "f"(a0));
    } else if (MODE == 7) {
      asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a0) : "f"(a1), "f"(a2)); asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a1) : "f"(a2), "f"(a3));
      asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a2) : "f"(a3), "f"(a0)); asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a3) : "f"(a0), "f"(a1));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __uint_as_float(b0 ^ b1 ^ b2 ^ b3) + (float)(c0 ^ c1 ^ c2 ^ c3);
}

template <int MODE>
void run(const char* name, int per_iter_elems) {
  float* out; cudaMalloc(&out, 148 * 1024 * 4 * 4);
  const int iters = 20000, blocks = 148 * 2, threads = 512;
  k<MODE><<<blocks, threads>>>(out, 100, 0.5f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, iters, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double ops = (double)blocks * threads * iters * 4;      // instructions (thread-level)
  double per_sm_per_ns = ops / 148 / (ms * 1e6);
  printf("%-28s %8.3f ms  %7.2f thread-instr/ns/SM  (~%.1f /clk/SM @1.9GHz)  elems/instr=%d  err=%s\n", name, ms, per_sm_per_ns,
         per_sm_per_ns / 1.9, per_iter_elems, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.ftz.bf16x2", 2);
  run<2>("ex2.approx.f16x2", 2);
  run<3>("fma.rn.f32", 1);
  run<4>("fma.rn.f32x2", 2);
  run<5>("cvt.rn.bf16x2.f32 (+fadd)", 2);
  run<6>("max.f32 2-input", 1);
  run<7>("max.f32 3-input", 2);
  return 0;
}
