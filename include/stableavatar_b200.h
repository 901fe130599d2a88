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
 *   res_mode 3: res is added BEFORE the activation, out = act(acc + bias + res) — K-chunked accumulation of the fp32 mode
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
/* The same attention with the sequence-parallel O exchange fused into its epilogue (replaces the second all-to-all of
 * wan/dist/wan_xfuser.py:102-110): args->out is ignored; query row r is stored to dst[r / rows_per_dst] at
 * [b, r % rows_per_dst, head, :] (element strides dst_bs / dst_ls, heads 128 elements apart), each destination typically
 * the o_recv buffer of the rank that owns those tokens, mapped over NVLink (sa_ipc_open) and already offset to this rank's
 * head columns. A 128-row tile leaves as TMA stores from a shared-memory staging tile; a tile that straddles two destinations
 * is written row by row (16-byte stores) instead.
 * 1 <= n_dst <= 8, n_dst * rows_per_dst >= q_len, accumulate must be 0. */
int sa_flash_attn_d128_sp(const sa_attn_args* args, void* const* dst, int32_t n_dst, int32_t rows_per_dst, int64_t dst_bs,
                          int64_t dst_ls, sa_stream_t stream);

/* The three cross-attentions of WanI2VTalkingCrossAttention (1B.py:534-605) in one launch: out (+)= sum over the key sets
 * of softmax(q K_s^T * scale) V_s, each partial result rounded to bf16 before the bf16 add (set order = argument order),
 * q loaded once. q/out: bf16 [batch, q_len, heads, 128] views as in sa_attn_args. A set is [batch, kv_total, heads, 128]
 * (strides k_bs/k_ls, v_bs/v_ls). windowed = 0: all kv_len = kv_total keys. windowed = 1 (audio, 1B.py:575-586): the keys
 * are kv_total / kv_len consecutive windows of kv_len keys and query row r attends only to window
 * (tok_offset + r) / rows_per_group; requires (255 / rows_per_group + 2) * kv_len <= 64. */
typedef struct {
  const void* k;
  const void* v;
  int64_t k_bs, k_ls, v_bs, v_ls;
  int32_t kv_len, kv_total, windowed;
} sa_cross_attn_set;
typedef struct {
  const void* q;
  void* out;
  int64_t q_bs, q_ls, o_bs, o_ls;
  int32_t batch, heads, q_len, n_sets;
  float scale;
  int32_t accumulate, rows_per_group, tok_offset;
  sa_cross_attn_set set[3];
} sa_cross_attn_args;
int sa_cross_attn3_d128(const sa_cross_attn_args* args, sa_stream_t stream);

/* Same contract for a handful of queries per (batch, head) and any head_dim % 8 == 0: the audio adapter's
 * cross-attention, 15 audio tokens x 1560 video tokens x 8 heads of 192
 * (wan/models/vocal_projector_fantasy_1B.py:259-277; SDPA branch :178-203), and 8 heads of 640 for the 14B adapter
 * (vocal_projector_fantasy_14B.py). q_len <= 32, head_dim <= 768, any kv_len (keys are tiled through shared memory
 * with a running softmax). accumulate must be 0. */
int sa_attn_small_q(const sa_attn_args* args, int32_t head_dim, sa_stream_t stream);

/* ---- LayerNorm (+affine) (+AdaLN modulation) (+gated self-residual) ------------------------------------------
 * y = (x - mean) * rstd [* weight + bias]                     WanLayerNorm, 1B.py:345-355 (norm1/norm2/norm3/head.norm),
 *                                                             nn.LayerNorm in MLPProj (:731-734) and VocalProjModel (vp1B:396-399)
 * y = y * (1 + scale[b]) + shift[b]                           1B.py:675, 687, 721-722; vp1B.py:345, 354, 386
 * out = res + y * gate[b]                                     vp1B.py:345-347 (adapter "pseudo self-attention")
 * b = row / rows_per_batch; shift/scale/gate: bf16, batch stride mod_bs elements. round_bf16 = 1 replicates the bf16
 * rounding after every elementwise op (bf16 residual stream under autocast); 0 keeps fp32 (adapter stream).
 * x: x_dtype, out (and res): out_dtype, weight/bias: w_dtype. C % 8 == 0, C <= 2048.
 */
typedef struct {
  const void* x;
  void* out;
  const void* weight;
  const void* bias;
  const void* shift;
  const void* scale;
  const void* gate;
  const void* res;
  int64_t ldx, ldo, ldr, mod_bs;
  int32_t rows, C, rows_per_batch;
  int32_t x_dtype, out_dtype, w_dtype, round_bf16;
  float eps;
} sa_ln_args;
int sa_layernorm_modulate(const sa_ln_args* args, sa_stream_t stream);

/* ---- RMSNorm over the full channel dim (+ 3-D RoPE), in place on bf16 ------------------------------------------
 * x = bf16(bf16(x * rsqrt(mean(x^2) + eps)) * weight)          WanRMSNorm, 1B.py:326-342 (norm_q/norm_k/norm_k_img, vp1B norm_q/k)
 * then, if freqs != NULL, token t = tok_offset + row % rows_per_batch < F*H*W is rotated pairwise by freqs[pos][j] (cos, sin),
 * pos = frame / row / column index for j < 22 / < 43 / < 64       rope_apply, 1B.py:296-323; table 1B.py:855-862
 * Up to two segments (q and k of a fused QKV GEMM output) share the launch: x/weight and x2/weight2 (x2 may be NULL).
 * freqs: float [1024][64][2]. C % 8 == 0 (C % 128 == 0 with RoPE), ld % 8 == 0.
 */
typedef struct {
  void* x;
  const void* weight;
  void* x2;
  const void* weight2;
  const void* freqs;
  int64_t ld;
  int32_t rows, C, rows_per_batch, F, H, W;
  int32_t tok_offset; /* global index of this shard's first token (sequence parallelism), else 0 */
  float eps;
} sa_rms_args;
int sa_rmsnorm_rope(const sa_rms_args* args, sa_stream_t stream);

/* out[i, j, :] = bf16(a[i, :] + b[j, :]) (bf16): e = modulation + e0 of every block in one launch (1B.py:672). */
int sa_add_bcast_bf16(const void* a, const void* b, void* out, int32_t na, int32_t nb, int32_t n, sa_stream_t stream);

/* ---- patch embedding operand / output rearrangement ------------------------------------------------------------
 * sa_patchify: A[b, tok, c*4 + q*2 + r] = cat(x, y)[b, c, f, 2h+q, 2w+r] (bf16), rows past F*H/2*W/2 and columns past
 * 4*(Cx+Cy) zero — the K-major operand that turns patch_embedding (Conv3d k=s=(1,2,2), 1B.py:830-831, 972-983) into
 * a GEMM against weight.flatten(1). sa_unpatchify: 1B.py:1161-1184 on the head output u[b, tok, (q*2+r)*Cout + c].
 */
int sa_patchify(const void* x, const void* y, void* out, int32_t B, int32_t Cx, int32_t Cy, int32_t F, int32_t H,
                int32_t W, int32_t seq_len, int32_t K_pad, sa_stream_t stream);
int sa_unpatchify(const void* u, void* out, int64_t u_bs, int64_t u_ls, int32_t B, int32_t Cout, int32_t F, int32_t H,
                  int32_t W, sa_stream_t stream);

/* out[m, n] = sum_k pre(x[m, k]) * W[n, k] + bias[n] in fp32 (M <= 8): the time-embedding island that the reference
 * runs under autocast(dtype=float32) (1B.py:986-990). pre: 0 none, 1 SiLU, 2 x is t[M] and the input row is
 * sinusoidal_embedding_1d(K, t) computed in fp64 (1B.py:210-220). Writes fp32 and/or bf16 copies. */
int sa_small_linear_f32(const void* x, const void* w, const void* bias, void* out_f32, void* out_bf16, int32_t M,
                        int32_t N, int32_t K, int32_t pre, int32_t w_dtype, sa_stream_t stream);

/* out[r, :] = idx[r] >= 0 ? src[idx[r], :] : 0 (fp32 rows): split_tensor_with_padding, vocal_projector_fantasy.py:81-131. */
int sa_gather_rows_f32(const void* src, const void* idx, void* out, int32_t rows, int32_t C, sa_stream_t stream);

/* ---- scheduler step ------------------------------------------------------------------------------------------
 * cfg != 0: pred = [uncond, drop_audio, cond] (3 x n bf16); noise = uncond + audio_scale*(drop_audio - uncond) +
 * text_scale*(cond - drop_audio), each op rounded to bf16 (wan/pipeline/wan_inference_long_pipeline.py:751-753).
 * out = bf16(float(latents) + bf16(dsigma * noise)): FlowMatchEulerDiscreteScheduler.step of diffusers 0.30.1 (the
 * 0-dim fp32 (sigma_next - sigma) times the bf16 prediction is a bf16 tensor under torch type promotion)
 * (pipe.py:754). noise_out (optional) receives the combined prediction.
 * dsigma_dev (optional, device float): when non-NULL it replaces dsigma, so a captured CUDA graph of the step can be
 * replayed with a new schedule value.
 * latents_dtype: SA_BF16, or SA_F32 for a caller-supplied fp32 sample (the reference keeps the caller's dtype for the
 * first step's `sample.float()`, pipe.py:393-396); out is always bf16 = model_output.dtype. */
int sa_cfg_euler_step(const void* pred, const void* latents, void* out, void* noise_out, int64_t n, float audio_scale,
                      float text_scale, float dsigma, const void* dsigma_dev, int32_t cfg, int32_t latents_dtype,
                      sa_stream_t stream);

/* ---- sliding-window write-back with the overlap blend (wan/pipeline/wan_inference_long_pipeline.py:756-779) --------
 * All windows of one denoise step in one launch, in window order: for window k (frames[k] latent frames starting at
 * latent frame start[k], previous window ending at prev_end[k]) with blend[k] != 0 the first `overlap` frames become
 * new * w_j + pred_latents[(prev_end - overlap + j) % N] * (1 - w_j) — rounded like the reference's bf16 tensor ops —
 * then every frame is written to pred_latents[(start + i) % N]. new_latents: bf16 [n_windows, C, f_max, HW];
 * pred_latents: [C, N, HW] bf16 or f32 (pred_dtype), zero-initialised by the caller; weight / one_minus_weight: the
 * bf16-rounded values of the reference's weight tensor and of (1 - weight). */
typedef struct {
  const void* new_latents;
  void* pred_latents;
  int32_t n_windows, C, N, HW, f_max, overlap, pred_dtype;
  int32_t start[64], frames[64], prev_end[64], blend[64];
  float weight[64], one_minus_weight[64];
} sa_window_blend_args;
int sa_window_blend(const sa_window_blend_args* args, sa_stream_t stream);

/* ---- Wan VAE decode (wan/models/wan_vae.py) --------------------------------------------------------------------
 * Activations are channels-last bf16 [T, H, W, C] (the reference is NCTHW fp32; layout and compute dtype are internal
 * to the decode, the boundary stays AutoencoderKLWan.decode's NCTHW fp32 in / out).
 *
 * sa_conv3d_cl: causal conv as an implicit GEMM on tcgen05 — CausalConv3d (wan_vae.py:20-39) incl. the 3x3x3 convs
 * of ResidualBlock (:197-203), Decoder3d.conv1 / head (:394, 421-424), Resample's (3,1,1) time_conv and 3x3 Conv2d
 * (:79-88). `in` holds Tout + KT - 1 frames: the KT - 1 leading ones are the causal cache (zeros before the first
 * chunk), replacing the per-chunk clone/cat of :208-220; H/W 'same' zero padding is implicit.
 *   w: bf16 [ceil16(Cout), KT*KH*KW*Cin], K index = ((kt*KH + kh)*KW + kw)*Cin + c; bias: f32 [Cout]; Cin % 32 == 0.
 *   out_mode 0: out bf16 [Tout,H,W,Cout] = acc + bias (+ res, same layout: the x + h of :223)
 *   out_mode 1: Cout = 2C; channel n of frame t goes to frame 2t + n/C, channel n%C of bf16 [2*Tout,H,W,C] (:137-140)
 *   out_mode 2: out f32 planar [Cout, out_T_total, H, W] at frame out_t0 + t, clamped to [-1, 1] (:668)
 *   out_mode 3: out f32 channels-last [Tout,H,W,Cout], unclamped (Encoder3d.head, :320-322)
 * Encode side (Encoder3d :268-369, Resample downsample2d/3d :95-105, 145-162): pad_h / pad_w = zero rows / columns in
 * front of H / W (negative = KH/2, KW/2, the 'same' padding; the back padding is whatever the taps reach, zero filled;
 * output H x W always equals input H x W); stride_t = temporal stride (0 = 1), `in` then holds (Tout-1)*stride_t + KT
 * frames. ZeroPad2d(0,1,0,1) + Conv2d(3, stride 2) (:96-98) is run as KH = KW = 2, pad 0 over the space-to-depth
 * input of sa_vae_space_to_depth with the 3x3 weights scattered into the 2x2x(4C) taps; time_conv stride (2,1,1)
 * (:103-104) is KT = 3, stride_t = 2.
 */
typedef struct {
  const void* in;
  const void* w;
  const void* bias;
  const void* res;
  void* out;
  int32_t Tout, H, W, Cin, Cout, KT, KH, KW;
  int32_t out_mode, out_T_total, out_t0;
  int32_t pad_h, pad_w, stride_t;
} sa_conv_args;
int sa_conv3d_cl(const sa_conv_args* args, sa_stream_t stream);
/* The same convolution for the shapes that carry the decoder and encoder (KT x 3 x 3 with KT = 1 or 3, 'same' padding,
 * stride 1, Cout 96 with Cin % 48 == 0 or Cout 192 / 384 with Cin % 96 == 0, out_mode 0 / 1; sa_conv3d_halo_supported says
 * which): the input halo of an output tile is staged once in shared memory and the 9 spatial taps are shared-memory
 * descriptor offsets; weights are shared by 2-4 output tiles. args->w is the weight PRE-PACKED as bf16
 * [Cout / BN][KT*9 taps (kt,kh,kw)][Cin / 8][BN][8], BN = 96 for Cout 96, else 192 (wan_vae.py:20-39, 69-143). */
int sa_conv3d_halo_supported(int32_t Cin, int32_t Cout, int32_t KT, int32_t KH, int32_t KW, int32_t stride_t, int32_t out_mode);
int sa_conv3d_halo_cl(const sa_conv_args* args, sa_stream_t stream);

/* out = x / max(||x||_2, 1e-12) * sqrt(C) * gamma per position, then SiLU if silu != 0: RMS_norm (wan_vae.py:42-57) +
 * nn.SiLU of ResidualBlock / head / AttentionBlock.norm. x/out bf16 [P, C], gamma f32 [C]. */
int sa_vae_rmsnorm_silu(const void* x, const void* gamma, void* out, int64_t P, int32_t C, int32_t silu, sa_stream_t stream);
/* out[t, y, x, :] = in[t, y/2, x/2, :]: Upsample(scale 2, nearest-exact) of Resample (wan_vae.py:60-66, 79-84). */
int sa_vae_upsample2x(const void* in, void* out, int32_t T, int32_t H, int32_t W, int32_t C, sa_stream_t stream);
/* out[r, :] = bf16(softmax(in[r, :] * scale)), in f32: the softmax of AttentionBlock's SDPA (wan_vae.py:254-259). */
int sa_softmax_rows(const void* in, void* out, int32_t rows, int32_t n, int64_t ld_in, int64_t ld_out, float scale,
                    sa_stream_t stream);
/* x = conv2(z / (1/std) + mean): latent de-normalisation + 1x1x1 conv (wan_vae.py:552-559). z f32 [Cz, P] planar ->
 * bf16 [P, Cpad] channels-last, channels >= Cz zero. wc f32 [Cz, Cz], bc/mean/stdv f32 [Cz]. */
int sa_vae_latent_in(const void* z, const void* wc, const void* bc, const void* mean, const void* stdv, void* out,
                     int32_t Cz, int64_t P, int32_t Cpad, sa_stream_t stream);


/* ---- sequence-parallel exchange over NVLink peer memory (replaces wan/dist/wan_xfuser.py:102-110) -----------------
 * Ulysses all-to-all of the DiT self-attention as direct peer stores: the caller maps every rank's receive buffers
 * (CUDA IPC) and passes the P base pointers. Head split: hg head groups x qs = P / hg query splits, rank = g * qs + s.
 *   sa_sp_scatter_qkv: src = local q|k|v rows [B, Ll, 3, heads, 128] bf16 (row stride ld elements) ->
 *       dst_a[r] = rank r's kv_recv [B, P, Ll, 2, heads/hg, 128] (slot = this rank) for the qs ranks of each head group,
 *       dst_b[r] = rank r's q_recv [B, hg, Ll, heads/hg, 128] (slot = this rank / qs) for the rank with s == this rank % qs.
 *   sa_sp_scatter_o: src = attention output [B, hg, Ll, heads/hg, 128] (source ranks i * qs + s) ->
 *       dst_a[r] = rank r's o_recv [B, Ll, heads, 128], head columns of this rank's group.
 *   sa_sp_barrier: sig[r] = rank r's flag array (uint32 [P], zero-initialised), epoch = local uint32 counter. Orders all
 *       earlier stores of this stream before, and all peers' earlier stores after. The spin is bounded (default 10 min,
 *       sa_sp_set_barrier_timeout_ms; 0 = unbounded like NCCL): on expiry the kernel prints the missing rank and traps,
 *       which leaves a sticky context error — ranks must enter every forward within the timeout of each other.
 * No launch is issued on a peer's device; the kernels only store through the mapped pointers. */
typedef struct {
  const void* src;
  void* dst_a[8];
  void* dst_b[8];
  int64_t ld;
  int32_t B, Ll, heads, head_dim, P, rank, hg;
  int32_t b_first, b_count; /* CFG samples [b_first, b_first + b_count) handled by this launch; b_count 0 = up to B */
} sa_sp_args;
int sa_sp_scatter_qkv(const sa_sp_args* args, sa_stream_t stream);
int sa_sp_scatter_o(const sa_sp_args* args, sa_stream_t stream);
int sa_sp_barrier(void* const* sig, void* epoch, int32_t P, int32_t rank, sa_stream_t stream);
int sa_sp_set_barrier_timeout_ms(int64_t ms);
/* Producer fusion of sa_rmsnorm_rope + sa_sp_scatter_qkv: q and k of the fused QKV rows (args->src, un-normalised) get the
 * WanRMSNorm + 3-D RoPE of 1B.py:296-342 (token index = tok_offset + row % Ll, same arithmetic and rounding as
 * sa_rmsnorm_rope), v is copied, and every chunk goes straight to its destination rank's kv_recv / q_recv (layouts as
 * above). The local rows are left untouched. weight_q / weight_k: bf16 [heads * 128]; freqs as in sa_rms_args. */
int sa_sp_norm_rope_scatter(const sa_sp_args* args, const void* weight_q, const void* weight_k, const void* freqs,
                            int32_t F, int32_t H, int32_t W, int32_t tok_offset, float eps, sa_stream_t stream);
/* CUDA IPC for the mappings above. sa_ipc_export: 64-byte handle of the cudaMalloc allocation containing ptr + the byte
 * offset of ptr inside it. sa_ipc_open: map a PEER process's allocation; call with the device that will launch the
 * scatter kernels current (peer access is enabled for that device); returns the allocation base. sa_ipc_close unmaps. */
int sa_ipc_export(const void* ptr, void* handle64, int64_t* offset);
int sa_ipc_open(const void* handle64, void** base);
int sa_ipc_close(void* base);

/* ---- Wan VAE encode helpers (wan/models/wan_vae.py:519-547) -------------------------------------------------------
 * out[t, y, x, (dy*2 + dx)*C + c] = in[t, 2y + dy, 2x + dx, c]: bf16 [T,H,W,C] -> [T,H/2,W/2,4C] (H, W even), the input
 * layout of the stride-2 Conv2d of Resample('downsample2d'/'downsample3d') (:96-104). */
int sa_vae_space_to_depth(const void* in, void* out, int32_t T, int32_t H, int32_t W, int32_t C, sa_stream_t stream);
/* Video in: x f32 planar [Cx, P] (P = T*H*W positions of one sample) -> bf16 channels-last [P, Cpad], channels >= Cx
 * zero: the input of Encoder3d.conv1 (:291). */
int sa_vae_video_in(const void* x, void* out, int32_t Cx, int64_t P, int32_t Cpad, sa_stream_t stream);
/* Latent out: h f32 channels-last [P, 2*Cz] (Encoder3d.head output) -> conv1 (1x1x1, wc f32 [2Cz, 2Cz], bc f32 [2Cz])
 * -> out f32 planar [2*Cz, P]: channels [0,Cz) = (mu - mean) / stdv, channels [Cz,2Cz) = log-variance (:539-545). */
int sa_vae_latent_out(const void* h, const void* wc, const void* bc, const void* mean, const void* stdv, void* out,
                      int32_t Cz, int64_t P, sa_stream_t stream);

/* ---- fp32 mode (BASELINE config 1: fp32 weights, per-block tolerance 1e-4) ----------------------------------------
 * A fp32 Linear runs on the bf16 tensor cores as ONE sa_gemm_bf16 over a six-fold K: both operands are split into three
 * bf16 terms (x = x0 + x1 + x2) and laid out as [x0|x0|x1|x1|x0|x2] (pattern 0, activation side) and
 * [w0|w1|w0|w1|w2|w0] (pattern 1, weight side), so the fp32 accumulator sums the six significant cross products
 * (error ~2^-24). sa_f32_split3: x f32 [M, K] (row stride ld) -> out bf16 [M, 6*K_pad], columns K..K_pad zero. Call sa_gemm_bf16 on the two split
 * operands with out_dtype f32, round_y 0 (GELU then uses libm tanhf). The rest are the fp32 forms of the elementwise
 * steps (same reference lines as their bf16 versions): patchify/unpatchify (1B.py:972-983, 1161-1184), modulation
 * out = x (1 + scale[b]) + shift[b] and gated residual h += y * gate[b] (1B.py:675-691; gate NULL = plain add),
 * RMSNorm (+3-D RoPE) in place (1B.py:296-342), in-place row softmax of x * scale (1B.py:158-207), broadcast add
 * out[i,j,:] = a[i,:] + b[j,:] (1B.py:672), CFG + Euler (pipe.py:752-754). */
int sa_f32_split3(const void* x, int64_t ld, int64_t M, int32_t K, int32_t K_pad, void* out, int32_t pattern,
                  sa_stream_t stream);
int sa_f32_patchify(const void* x, const void* y, void* out, int32_t B, int32_t Cx, int32_t Cy, int32_t F, int32_t H,
                    int32_t W, int32_t seq_len, int32_t K_pad, sa_stream_t stream);
int sa_f32_unpatchify(const void* u, void* out, int64_t u_bs, int64_t u_ls, int32_t B, int32_t Cout, int32_t F, int32_t H,
                      int32_t W, sa_stream_t stream);
int sa_f32_modulate(const void* x, const void* shift, const void* scale, void* out, int64_t rows, int32_t C,
                    int32_t rows_per_batch, int64_t mod_bs, sa_stream_t stream);
int sa_f32_gated_add(void* h, const void* y, int64_t ldy, const void* gate, int64_t rows, int32_t C, int32_t rows_per_batch,
                     int64_t gate_bs, sa_stream_t stream);
int sa_f32_rmsnorm_rope(void* x, int64_t ld, const void* weight, const void* freqs, int32_t rows, int32_t C,
                        int32_t rows_per_batch, int32_t F, int32_t H, int32_t W, float eps, sa_stream_t stream);
int sa_f32_softmax_rows(void* x, int32_t rows, int32_t n, int64_t ld, float scale, sa_stream_t stream);
int sa_f32_add_bcast(const void* a, const void* b, void* out, int32_t na, int32_t nb, int32_t n, sa_stream_t stream);
int sa_f32_cfg_euler_step(const void* pred, const void* latents, void* out, void* noise_out, int64_t n, float audio_scale,
                          float text_scale, float dsigma, const void* dsigma_dev, int32_t cfg, sa_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* STABLEAVATAR_B200_H_ */
