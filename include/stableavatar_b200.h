/* stableavatar_b200.h — C-ABI of the B200-native StableAvatar denoising hot path.
 *
 * The reference (yangyifeng1128/StableAvatar) has no FFI layer: its operator surface is Python/PyTorch
 * (SURVEY.md §8b). This header is the boundary a maintainer would bind from that Python code (ctypes, see
 * INTEGRATION.md): every entry point takes plain device pointers + sizes + a cudaStream_t, returns 0 or a negative
 * error code (text via sa_last_error), never throws, never allocates and never synchronises. Each op cites the
 * reference call site (file:line under /root/reference) whose arithmetic it replaces.
 *
 * All tensors are row-major with the innermost dimension contiguous. "bf16" = __nv_bfloat16, "f32" = float.
 * dtype codes: SA_BF16 = 0, SA_F32 = 1.
 */
#ifndef STABLEAVATAR_B200_H_
#define STABLEAVATAR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* sa_stream_t; /* cudaStream_t */

#define SA_BF16 0
#define SA_F32 1

/* ---- library ---------------------------------------------------------------------------------------------- */
int sa_version(void);
const char* sa_last_error(void);
int sa_device_sm_count(void);

/* ---- GEMM with fused epilogue (tcgen05 / TMEM / TMA) -------------------------------------------------------
 * C[M,N] = epilogue(A[M,K] . W[N,K]^T): replaces every nn.Linear on the path — self-attn q/k/v/o
 * (wan/models/wan_fantasy_transformer3d_1B.py:395-397,412), cross-attn q/k/v/o/k_img/v_img/k_vocal/v_vocal (:550-554,
 * :577-578,:604), FFN (:644-646,690), text/time/CLIP MLPs (:832-838,:731-734), head (:710,722), patch embedding
 * (:830-831 as a K=in_dim*4 GEMM), audio adapter Linears (wan/models/vocal_projector_fantasy_1B.py:238-241,313-316,
 * 374,393) — with the elementwise work that follows each of them folded in:
 *   y   = acc + bias                      (f32)          ; then y = bf16(y) if round_y (autocast Linear output)
 *   y   = act(y)                          act: 0 none, 1 GELU-tanh (1B.py:645), 2 SiLU (1B.py:837), 3 GELU-erf (:733)
 *   out = y                               res_mode 0
 *       = res + y                         res_mode 1   (1B.py:684  x + cross_attn(...))
 *       = res + bf16(y * gate[row/rows_per_batch, n])   res_mode 2   (1B.py:679,691  x + y * e[k])
 * A, W: bf16 (lda/ldw = row strides in elements, multiples of 8). bias: bf16/f32 [N] or NULL. gate: bf16 [*, N] with
 * row stride gate_ld. res/out: bf16 or f32 per res_dtype/out_dtype with row strides ldr/ldc. res may alias out.
 */
typedef struct {
  const void* a;
  const void* w;
  void* out;
  const void* bias;
  const void* res;
  const void* gate;
  int64_t lda, ldw, ldc, ldr, gate_ld;
  int32_t M, N, K;
  int32_t bias_dtype, out_dtype, res_dtype;
  int32_t act, res_mode, round_y;
  int32_t rows_per_batch;
} sa_gemm_args;
int sa_gemm_bf16(const sa_gemm_args* args, sa_stream_t stream);

/* ---- flash attention, head_dim 128 (tcgen05 / TMEM / TMA) --------------------------------------------------
 * out[b, i, h, :] (+)= softmax_j(q[b,i,h,:] . k[b,j,h,:] * scale) v[b,j,h,:], no mask, no dropout: the SDPA branch of
 * attention() (1B.py:158-207; k_lens ignored there, 1B.py:190-194) used by self-attention (:402-407), text and CLIP
 * cross-attention (:556-569) and the per-latent-frame audio cross-attention (:575-586, batch = B*G groups).
 * q/out: [batch, q_len, heads, 128] with strides (q_bs, q_ls, 128, 1) elements; k/v: [batch, kv_len, heads, 128] with
 * strides (kv_bs, kv_ls, 128, 1). accumulate != 0 adds into out (bf16 add: x + img_x + vocal_x, 1B.py:603).
 */
typedef struct {
  const void* q;
  const void* k;
  const void* v;
  void* out;
  int64_t q_bs, q_ls, k_bs, k_ls, v_bs, v_ls, o_bs, o_ls; /* element strides of batch and token dims */
  int32_t batch, heads, q_len, kv_len;
  float scale;
  int32_t accumulate;
} sa_attn_args;
int sa_flash_attn_d128(const sa_attn_args* args, sa_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* STABLEAVATAR_B200_H_ */
