// sp_exchange.cu — the sequence-parallel (Ulysses) all-to-all of the DiT self-attention as direct NVLink peer stores.
//
// Replaces xFuserLongContextAttention's NCCL all-to-alls (wan/dist/wan_xfuser.py:102-110) and this repo's own
// dist.all_to_all_single path: every rank maps the receive buffers of all ranks of the NVSwitch domain (CUDA IPC) and
// ONE kernel reads the rank's local q|k|v rows once and stores every 16-byte chunk straight into its destination
// rank's receive buffer, already in the token-major layout the attention kernel's TMA descriptors read — pack, transfer
// and unpack fused, no staging copy, no NCCL launch, and (being plain kernels) capturable in the step's CUDA graph.
// A flag barrier over the same peer mappings (release/acquire at system scope, monotonic epochs, bounded spin) orders
// the stores against the readers.
//
// Head / token split (sequence_parallel.plan): hg = gcd(heads, P) head groups x qs = P / hg query splits; rank
// r = g * qs + s owns head group g (hp = heads / hg heads) and the query tokens of the source ranks r' with r' % qs == s.
//   kv_recv on rank r : [B, P (source rank), Ll, 2 (k, v), hp, d]   = K / V of [B, L, hp, d] with token stride 2*hp*d
//   q_recv  on rank r : [B, P / qs (source ranks r' % qs == s, ascending), Ll, hp, d]
//   o_recv  on rank r : [B, Ll, heads, d]   — the layout the output projection reads, every head group filled by its owner
// The batch (CFG sample) index is outermost everywhere, and every entry point takes a sample range [b_first, b_first +
// b_count): the host pipelines the exchange of sample b + 1 under the attention of sample b (sequence_parallel.py).
#include "../../include/stableavatar_b200.h"
#include "sa_host.h"
#include "sa_ptx.cuh"
#include <string.h>

#include <atomic>

namespace sa {
namespace sp {

constexpr int MAX_RANKS = 8;
// Spin bound of sa_sp_barrier. Default 10 min: long enough for host-side skew between ranks (one rank writing a video,
// a graph capture, GC), short enough that a dead peer does not hang the GPU for ever. 0 = unbounded. sa_sp_set_barrier_timeout_ms.
static std::atomic<uint64_t> g_barrier_timeout_ns{600ull * 1000000000ull};
constexpr int THREADS = 192;  // 3 * 12 heads * 16 chunks = 576 = 3 x 192 columns per token for the 1.3B model

struct ScatterParams {
  const uint4* src;
  uint4* dst_a[MAX_RANKS];  // qkv: kv_recv bases; o: o_recv bases
  uint4* dst_b[MAX_RANKS];  // qkv: q_recv bases
  long long ld8;            // source row stride in 16-byte chunks
  int B, Ll, nh, P, rank, hg, qs, hp, n_src;
  int b_first, b_count;     // CFG samples handled by this launch
};

// d = 128 -> 16 chunks of 16 bytes per head. A thread owns ONE 16-byte column e of the token row [3, nh, 16] (its
// destination ranks and intra-token offsets are loop invariants) and walks tokens with a grid stride, so the per-chunk
// cost is one 32-bit divmod and a few IMADs; a warp reads 512 contiguous bytes and each destination receives
// contiguous runs of hp * 256 bytes.
__global__ void __launch_bounds__(THREADS) scatter_qkv_kernel(const ScatterParams p) {
  const int per_tok = 3 * p.nh * 16;
  const int e = blockIdx.y * THREADS + threadIdx.x;
  if (e >= per_tok) return;
  const int c = e & 15, h = (e >> 4) % p.nh, which = (e >> 4) / p.nh;  // which: 0 q, 1 k, 2 v
  const int g = h / p.hp, hl = h % p.hp;
  const int s_me = p.rank % p.qs;
  const int n_tok = p.b_count * p.Ll;
  // four tokens per iteration: the loads are issued before the (possibly aliasing, as far as the compiler knows) stores
  constexpr int U = 4;
  const bool is_q = which == 0;
  const int n_slots = is_q ? p.n_src : p.P, slot = is_q ? p.rank / p.qs : p.rank;
  const int row_chunks = is_q ? p.hp * 16 : 2 * p.hp * 16;
  const int inner = is_q ? hl * 16 + c : ((which - 1) * p.hp + hl) * 16 + c;
  uint4* const q_dst = p.dst_b[g * p.qs + s_me];
  for (int bt0 = blockIdx.x; bt0 < n_tok; bt0 += U * gridDim.x) {
    uint4 val[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int bt = bt0 + u * gridDim.x;
      if (bt < n_tok) val[u] = __ldg(p.src + ((long long)p.b_first * p.Ll + bt) * p.ld8 + e);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int bt = bt0 + u * gridDim.x;
      if (bt >= n_tok) continue;
      const int t = bt % p.Ll, b = p.b_first + bt / p.Ll;
      const long long off = (((long long)b * n_slots + slot) * p.Ll + t) * row_chunks + inner;
      if (is_q) {
        q_dst[off] = val[u];
      } else {
        for (int s = 0; s < p.qs; ++s) p.dst_a[g * p.qs + s][off] = val[u];
      }
    }
  }
}

// src: attention output [B, n_src, Ll, hp, 16 chunks] of head group g for the query tokens of source ranks
// i * qs + s_me; chunk goes to rank i * qs + s_me at [b, t, g * hp + hl, c]. One 16-byte chunk per thread and iteration,
// any hp (14B at P = 2: 20 heads per rank).
__global__ void __launch_bounds__(256) scatter_o_kernel(const ScatterParams p) {
  const int cols = p.hp * 16;
  const int s_me = p.rank % p.qs, g = p.rank / p.qs;
  const unsigned rows_b = (unsigned)p.n_src * p.Ll;                 // rows per sample
  const unsigned total = (unsigned)p.b_count * rows_b * cols;      // < 2^31, checked by the host
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const unsigned row = idx / cols, col = idx - row * cols;
    const unsigned bl = row / rows_b, lt = row - bl * rows_b;
    const unsigned src = lt / p.Ll, t = lt - src * p.Ll;
    const int b = p.b_first + bl;
    p.dst_a[src * p.qs + s_me][((long long)b * p.Ll + t) * (p.nh * 16) + g * cols + col] =
        __ldg(p.src + ((long long)b * rows_b + lt) * cols + col);
  }
}

struct BarrierParams {
  uint32_t* sig[MAX_RANKS];  // sig[r]: rank r's flag array [P] (peer mapping; sig[rank] is local)
  uint32_t* epoch;           // local device counter: number of barriers passed so far
  int P, rank;
  uint64_t timeout_ns;       // 0 = wait for ever (what NCCL does)
};

// One block, one thread per rank. All stores issued by earlier kernels of this stream (the scatters) are ordered before
// the flag by the fence + release; the acquire orders the peers' stores before everything launched after this kernel.
__global__ void barrier_kernel(const BarrierParams p) {
  __shared__ uint32_t e_sh;
  if (threadIdx.x == 0) e_sh = *p.epoch + 1;
  __syncthreads();
  const uint32_t e = e_sh;
  const int r = threadIdx.x;
  if (r < p.P) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.sig[r] + p.rank), "r"(e) : "memory");
    const uint32_t* mine = p.sig[p.rank] + r;
    uint32_t v;
    const uint64_t t0 = globaltimer_ns();
    uint32_t spins = 0;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int32_t)(v - e) >= 0) break;
      if (((++spins) & 0xff) == 0 && p.timeout_ns && globaltimer_ns() - t0 > p.timeout_ns) {
        printf("sa_sp_barrier: rank %d timed out waiting for rank %d (epoch %u, saw %u)\n", p.rank, r, e, v);
        __trap();
      }
    }
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0) *p.epoch = e;
}

static int fill(ScatterParams& p, const sa_sp_args* a, const char* who) {
  if (!a || !a->src || a->P < 1 || a->P > MAX_RANKS || a->rank < 0 || a->rank >= a->P || a->B <= 0 || a->Ll <= 0 ||
      a->heads <= 0 || a->hg <= 0 || a->P % a->hg || a->heads % a->hg || a->head_dim != 128) {
    set_error("%s: bad argument (1 <= P <= %d, hg | P, hg | heads, head_dim 128)", who, MAX_RANKS);
    return SA_ERR_BAD_ARG;
  }
  p.src = reinterpret_cast<const uint4*>(a->src);
  p.B = a->B; p.Ll = a->Ll; p.nh = a->heads; p.P = a->P; p.rank = a->rank; p.hg = a->hg; p.qs = a->P / a->hg;
  p.hp = a->heads / a->hg; p.n_src = a->P / p.qs;
  p.ld8 = a->ld / 8;
  p.b_first = a->b_first;
  p.b_count = a->b_count > 0 ? a->b_count : a->B - a->b_first;
  if (p.b_first < 0 || p.b_first + p.b_count > a->B) { set_error("%s: sample range [%d, %d) outside batch %d", who, p.b_first, p.b_first + p.b_count, a->B); return SA_ERR_BAD_ARG; }
  for (int r = 0; r < MAX_RANKS; ++r) {
    p.dst_a[r] = r < a->P ? reinterpret_cast<uint4*>(a->dst_a[r]) : nullptr;
    p.dst_b[r] = r < a->P ? reinterpret_cast<uint4*>(a->dst_b[r]) : nullptr;
  }
  return SA_OK;
}

}  // namespace sp
}  // namespace sa

extern "C" int sa_sp_scatter_qkv(const sa_sp_args* a, sa_stream_t stream) {
  using namespace sa;
  sp::ScatterParams p;
  int rc = sp::fill(p, a, "sa_sp_scatter_qkv");
  if (rc) return rc;
  if (a->ld % 8 || a->ld < 3LL * a->heads * 128) { set_error("sa_sp_scatter_qkv: ld must be a multiple of 8 and >= 3*heads*128"); return SA_ERR_BAD_ARG; }
  for (int r = 0; r < a->P; ++r)
    if (!a->dst_a[r] || !a->dst_b[r]) { set_error("sa_sp_scatter_qkv: null destination for rank %d", r); return SA_ERR_BAD_ARG; }
  const int per_tok = 3 * p.nh * 16;
  const int gy = (per_tok + sp::THREADS - 1) / sp::THREADS;
  int gx = sm_count() * 16 / gy;
  if (gx > p.b_count * p.Ll) gx = p.b_count * p.Ll;
  if (gx < 1) gx = 1;
  sp::scatter_qkv_kernel<<<dim3(gx, gy), sp::THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "scatter_qkv_kernel launch");
  return SA_OK;
}

extern "C" int sa_sp_scatter_o(const sa_sp_args* a, sa_stream_t stream) {
  using namespace sa;
  sp::ScatterParams p;
  int rc = sp::fill(p, a, "sa_sp_scatter_o");
  if (rc) return rc;
  for (int r = 0; r < a->P; ++r)
    if (!a->dst_a[r]) { set_error("sa_sp_scatter_o: null destination for rank %d", r); return SA_ERR_BAD_ARG; }
  const long long total = (long long)p.b_count * p.n_src * p.Ll * p.hp * 16;
  if (total >= (1LL << 31)) { set_error("sa_sp_scatter_o: more than 2^31 chunks in one launch"); return SA_ERR_UNSUPPORTED; }
  long long gx = (total + 255) / 256;
  if (gx > (long long)sm_count() * 8) gx = (long long)sm_count() * 8;
  sp::scatter_o_kernel<<<(int)gx, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "scatter_o_kernel launch");
  return SA_OK;
}

extern "C" int sa_sp_set_barrier_timeout_ms(int64_t ms) {
  if (ms < 0) { sa::set_error("sa_sp_set_barrier_timeout_ms: negative timeout"); return sa::SA_ERR_BAD_ARG; }
  sa::sp::g_barrier_timeout_ns.store((uint64_t)ms * 1000000ull, std::memory_order_relaxed);
  return sa::SA_OK;
}

extern "C" int sa_sp_barrier(void* const* sig, void* epoch, int32_t P, int32_t rank, sa_stream_t stream) {
  using namespace sa;
  if (!sig || !epoch || P < 1 || P > sp::MAX_RANKS || rank < 0 || rank >= P) { set_error("sa_sp_barrier: bad argument"); return SA_ERR_BAD_ARG; }
  sp::BarrierParams p;
  for (int r = 0; r < sp::MAX_RANKS; ++r) p.sig[r] = r < P ? reinterpret_cast<uint32_t*>(sig[r]) : nullptr;
  for (int r = 0; r < P; ++r)
    if (!p.sig[r]) { set_error("sa_sp_barrier: null flag array for rank %d", r); return SA_ERR_BAD_ARG; }
  p.epoch = reinterpret_cast<uint32_t*>(epoch);
  p.P = P; p.rank = rank;
  p.timeout_ns = sp::g_barrier_timeout_ns.load(std::memory_order_relaxed);
  sp::barrier_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "barrier_kernel launch");
  return SA_OK;
}

// ------------------------------------------------------------------------------------------------ CUDA IPC plumbing
// The handle describes the whole cudaMalloc allocation that contains `ptr` (torch's caching allocator sub-allocates),
// so the byte offset of `ptr` inside it travels with the handle. sa_ipc_open must be called with the device that will
// run the scatter kernels current: cudaIpcMemLazyEnablePeerAccess enables peer access for THAT device's context.
extern "C" int sa_ipc_export(const void* ptr, void* handle64, int64_t* offset) {
  using namespace sa;
  if (!ptr || !handle64 || !offset) { set_error("sa_ipc_export: null pointer"); return SA_ERR_BAD_ARG; }
  typedef CUresult (*PFN_range)(CUdeviceptr*, size_t*, CUdeviceptr);
  static PFN_range range = nullptr;
  if (!range) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fp) { set_error("cuMemGetAddressRange entry point unavailable"); return SA_ERR_CUDA; }
    range = reinterpret_cast<PFN_range>(fp);
  }
  CUdeviceptr base = 0;
  size_t size = 0;
  CUresult r = range(&base, &size, reinterpret_cast<CUdeviceptr>(ptr));
  if (r != CUDA_SUCCESS) { set_error("cuMemGetAddressRange failed (%d)", (int)r); return SA_ERR_CUDA; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaError_t e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), reinterpret_cast<void*>(base));
  if (e != cudaSuccess) return cuda_fail(e, "cudaIpcGetMemHandle");
  *offset = (int64_t)(reinterpret_cast<CUdeviceptr>(ptr) - base);
  return SA_OK;
}

extern "C" int sa_ipc_open(const void* handle64, void** base) {
  using namespace sa;
  if (!handle64 || !base) { set_error("sa_ipc_open: null pointer"); return SA_ERR_BAD_ARG; }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(base, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return cuda_fail(e, "cudaIpcOpenMemHandle");
  return SA_OK;
}

extern "C" int sa_ipc_close(void* base) {
  using namespace sa;
  if (!base) return SA_OK;
  cudaError_t e = cudaIpcCloseMemHandle(base);
  if (e != cudaSuccess) return cuda_fail(e, "cudaIpcCloseMemHandle");
  return SA_OK;
}
