// sa_host.h — host-side helpers shared by the C-ABI entry points: error reporting and TMA descriptor encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sa {

// Error codes returned by every extern "C" entry point (0 = ok).
enum : int {
  SA_OK = 0,
  SA_ERR_BAD_ARG = -1,     // shape / alignment / null pointer violates the documented contract
  SA_ERR_CUDA = -2,        // a CUDA runtime/driver call failed (see sa_last_error)
  SA_ERR_UNSUPPORTED = -3  // valid request the kernels do not implement (e.g. head_dim != 128)
};

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

// Encode a tiled TMA descriptor for a bf16 tensor. dims/strides are innermost-first; strides_bytes has rank-1
// entries (the innermost stride is the element size). Returns SA_OK or an error code.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, CUtensorMapSwizzle swizzle);

int sm_count();

// Opt a kernel into `bytes` of dynamic shared memory on the CURRENT device, once per (device, kernel). Thread-safe:
// the reference's app.py calls the pipeline from gradio worker threads.
int ensure_dyn_smem(const void* kernel, int bytes, const char* name);
template <class K>
inline int ensure_dyn_smem(K* kernel, int bytes, const char* name) {
  return ensure_dyn_smem(reinterpret_cast<const void*>(kernel), bytes, name);
}

}  // namespace sa
