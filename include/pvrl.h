/* pvrl.h -- C ABI of libpvrl_sm100.so: the Blackwell (sm_100a) kernels behind the ProcedureVRL
 * TimeSformer hot path.
 *
 * The reference has no FFI layer: every op below replaces a PyTorch-eager call site inside
 * lib/models/vit.py / tools/train_net.py (cited per function, paths relative to the reference root).
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller
 *     (the library allocates nothing persistent except a cache of TMA descriptors);
 *   - every call only enqueues work on `stream` (a cudaStream_t passed as void*), never syncs;
 *   - return value: 0 = ok, < 0 = argument / shape / arch error, > 0 = cudaError_t;
 *     pvrl_last_error() returns a thread-local message for the last non-zero return;
 *   - activations "act" are bf16 (act_dtype 0) or fp32 (act_dtype 1); the residual stream,
 *     statistics, parameters' gradients and all reductions are fp32;
 *   - token order inside a clip is (h w t) with t fastest, cls first: vit.py:131,142,406.
 */
#ifndef PVRL_H_
#define PVRL_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PVRL_ABI_VERSION 1

/* dtypes of activation buffers */
#define PVRL_BF16 0
#define PVRL_F32 1

/* Row maps: how logical row m of an op addresses the [Bc, S = 1 + HW*T, D] residual stream. */
#define PVRL_MAP_IDENT 0   /* row = m                                                              */
#define PVRL_MAP_SKIPCLS 1 /* m over Bc*HW*T : row = m + m/(HW*T) + 1        (vit.py:130 x[:,1:])   */
#define PVRL_MAP_SPATIAL 2 /* m=(b,t,n) over Bc*T*(HW+1): n>0 -> b*S+1+(n-1)*T+t, n==0 -> cls (vit.py:138-143) */
#define PVRL_MAP_PATCH 3   /* m=(b,t,n) over Bc*T*HW : row = b*S+1+n*T+t      (vit.py:393-407)      */
#define PVRL_MAP_CLS 4     /* m over Bc : row = m*S                            (vit.py:421)          */

/* GEMM epilogues */
#define PVRL_EPI_STORE 0  /* out[map(m)] = rowscale*(acc + bias)                     -> act dtype        */
#define PVRL_EPI_GELU 1   /* z = acc + bias ; out2 = gelu_erf(z) ; out = gelu_erf'(z) -> act dtype (vit.py:54-60) */
#define PVRL_EPI_DGELU 2  /* out = acc * aux[m]   (aux = the gelu_erf' saved by EPI_GELU)   -> act dtype        */
#define PVRL_EPI_RESID 3  /* out[map(m)] = resid[map(m)] + rowscale*(acc + bias) (+pos+time) -> fp32     */
#define PVRL_EPI_ATOMIC 4 /* out[m] += acc  (fp32 red.add; split-K weight gradients)                     */

typedef struct pvrl_geom {
  int32_t T;  /* frames                     */
  int32_t HW; /* patches per frame (196)    */
} pvrl_geom_t;

/* D[M,N] = A * B^T over a contraction of length K, bf16 operands, fp32 accumulate (tcgen05 + TMA).
 *   trans == 0 ("NT"): A is [M,K] row-major (lda), B is [N,K] row-major (ldb)       -- nn.Linear forward / dX
 *   trans == 1 ("TN"): A is [K,M] row-major (lda), B is [K,N] row-major (ldb)       -- dW = dY^T X
 * Replaces cuBLAS addmm/mm behind nn.Linear (vit.py:47-60,72-90,118), nn.Conv2d patch embed (vit.py:172-179)
 * and their autograd (SURVEY 2a K1,K4,K6,K8,K10,K15). */
typedef struct pvrl_gemm {
  int32_t M, N, K;
  int32_t trans;
  const void* A;
  int64_t lda;
  const void* B;
  int64_t ldb;
  int32_t epilogue;  /* PVRL_EPI_*                                        */
  int32_t out_dtype; /* PVRL_BF16 / PVRL_F32 (STORE, GELU, DGELU only)    */
  void* out;
  int64_t ldo;
  void* out2;        /* GELU: activations; RESID+MAP_SPATIAL: cls side buffer fp32 [Bc*T, N] */
  const float* bias; /* [N] or NULL                                       */
  const float* rowscale; /* DropPath factors mask/keep, indexed m / rs_div, or NULL (vit_utils.py:140-155) */
  int32_t rs_div;
  int32_t map;       /* PVRL_MAP_* applied to out / resid rows            */
  const void* aux;   /* DGELU: saved GELU derivative [M, ld_aux], act dtype = out_dtype */
  int64_t ld_aux;
  const float* resid;    /* RESID: fp32 residual, same row map / ldo as out */
  const float* add_pos;  /* RESID+MAP_PATCH: pos_embed [(1+HW), N]  (vit.py:373-389) */
  const float* add_time; /* RESID+MAP_PATCH: time_embed [T, N]      (vit.py:393-404) */
  pvrl_geom_t g;
  int32_t k_splits;  /* ATOMIC only: 0 = auto                             */
  float* colsum;     /* STORE / DGELU: NULL or fp32 [N] += column sums of the stored output (fused bias gradient) */
} pvrl_gemm_t;

int pvrl_gemm_bf16(const pvrl_gemm_t* d, void* stream);

/* ---- elementwise / normalisation / layout ------------------------------------------------------ */

/* frames fp32 [Bc,3,T,H,W] -> im2col rows (b,t,ph,pw) x K=(c,kh,kw), act dtype.  vit.py:176-179 */
int pvrl_patchify(const float* frames, void* out, int32_t out_dtype, int32_t Bc, int32_t T, int32_t H, int32_t W,
                  int32_t patch, void* stream);
/* The same im2col from uint8 frames [Bc,3,T,H,W] with the loader's normalisation fused (datasets/utils.py:309-326,
 * defaults.py:510,516): value = (u8/255 - mean3[c]) / std3[c].  mean3 / std3 are HOST arrays of 3 floats. */
int pvrl_patchify_u8(const uint8_t* frames, void* out, int32_t out_dtype, int32_t Bc, int32_t T, int32_t H, int32_t W,
                     int32_t patch, const float* mean3, const float* std3, void* stream);

/* x[b,0,:] = cls_token + pos_embed[0]  (vit.py:371-389) */
int pvrl_cls_init(float* x, const float* cls_token, const float* pos_embed, int32_t Bc, int32_t S, int32_t D,
                  void* stream);

/* y[m] = LN(x[src(m)]) * w + b, eps; stats[m] = (mean, rstd).  nn.LayerNorm(768, eps=1e-6): vit.py:102,108,115,225.
 * MAP_SPATIAL reads cls rows from x_cls (block input) and token rows from x. */
int pvrl_layernorm_fwd(const float* x, const float* x_cls, const float* w, const float* b, void* y, int32_t y_dtype,
                       float* stats, int32_t M, int32_t D, float eps, int32_t map, pvrl_geom_t g, void* stream);

/* dx[src(m)] += LN'(dy[m]); dw += sum dy*xhat; db += sum dy.  (autograd of the above) */
int pvrl_layernorm_bwd(const void* dy, int32_t dy_dtype, const float* x, const float* x_cls, const float* w,
                       const float* stats, float* dx, float* dw, float* db, int32_t M, int32_t D, int32_t map,
                       pvrl_geom_t g, void* stream);

/* The same, and while each updated dx row is still in registers it is also written out as the operand of the NEXT
 * sub-layer's backward GEMMs: emit_out (act dtype = dy_dtype) receives exactly what
 * pvrl_gather_cast(dx, emit_out, emit_map, emit_rowscale, emit_rs_div, emit_colsum) would produce after this call, so
 * that pass over dx disappears.  Supported (map -> emit_map), the three hand-overs of Block.forward's backward
 * (vit.py:130-157): IDENT -> SPATIAL (MLP -> spatial branch, cls rows replicated x 1/T), SPATIAL -> SKIPCLS (spatial ->
 * temporal branch), SKIPCLS -> IDENT (temporal branch -> previous block's MLP; the clips' cls rows, which a SKIPCLS
 * pass does not touch, are emitted from dx as they are) and SKIPCLS -> PATCH (first block -> patch embedding).
 * M must cover whole clips. */
int pvrl_layernorm_bwd_emit(const void* dy, int32_t dy_dtype, const float* x, const float* x_cls, const float* w,
                            const float* stats, float* dx, float* dw, float* db, int32_t M, int32_t D, int32_t map,
                            pvrl_geom_t g, void* emit_out, int32_t emit_map, const float* emit_rowscale,
                            int32_t emit_rs_div, float* emit_colsum, void* stream);

/* out[m] = act( rowscale[m/rs_div] * clsf(m) * src[map(m)] ), fp32 -> act dtype.  clsf = 1/T on the cls rows of
 * MAP_SPATIAL (backward of the mean over frames, vit.py:147-149), else 1.  colsum (NULL or fp32 [D]) += sum_m out[m]:
 * the bias gradient of the Linear whose dY this is, fused so that dY is not read a second time. */
int pvrl_gather_cast(const float* src, void* out, int32_t out_dtype, const float* rowscale, int32_t rs_div,
                     int32_t M, int32_t D, int32_t map, pvrl_geom_t g, float* colsum, void* stream);

/* x2[b,0,:] = x0[b,0,:] + mean_t side[b*T+t,:]   (vit.py:147-149,156) */
int pvrl_cls_merge(const float* x0, const float* side, float* x2, int32_t Bc, int32_t T, int32_t S, int32_t D,
                   void* stream);

/* out[n] += sum_m a[m,n]   (bias gradients) */
int pvrl_colsum(const void* a, int32_t a_dtype, int64_t lda, float* out, int32_t M, int32_t N, void* stream);

/* fp32 [rows, cols] -> copy w_out [rows, cols] and transposed copy wT_out [cols, rows] in out_dtype
 * (either output may be NULL): the bf16 operand copies of the fp32 master weights. */
int pvrl_cast_weight(const float* w, void* w_out, void* wT_out, int32_t out_dtype, int32_t rows, int32_t cols,
                     void* stream);

/* The same for many matrices in one launch.  descs_dev: DEVICE array of n descriptors sorted by tile0, where matrix i
 * owns the 64x64 tiles [tile0, tile0 + tiles_x * ceil(rows/64)), tiles_x = ceil(cols/64); rows and cols must be even;
 * out / outT may be NULL. */
typedef struct pvrl_cast_desc {
  const float* w;
  void* out;
  void* outT;
  int32_t rows, cols;
  int32_t tile0, tiles_x;
} pvrl_cast_desc_t;
int pvrl_cast_weight_multi(const pvrl_cast_desc_t* descs_dev, int32_t n, int32_t total_tiles, int32_t out_dtype,
                           void* stream);

/* Error-compensated operand split for the "bf16x3" parity mode: a = hi + lo with hi = bf16(a), lo = bf16(a - hi).
 * pattern 0: out = [hi | hi | lo], pattern 1: out = [hi | lo | hi]; along = 1 concatenates along columns
 * ([M, 3*K]), along = 0 along rows ([3*M, K]).  A GEMM over the 3x longer contraction then yields
 * hi*hi + hi*lo + lo*hi (relative error ~2^-17 per product instead of 2^-9). */
int pvrl_split3(const float* a, void* out, int32_t M, int32_t K, int32_t pattern, int32_t along, void* stream);

/* embed gradients: dpos[0] = dcls = sum_b dx[b,0]; dpos[1+n] = sum_{b,t}; dtime[t] = sum_{b,n}  (vit.py:371-407) */
int pvrl_embed_bwd(const float* dx, float* dcls, float* dpos, float* dtime, int32_t Bc, int32_t D, pvrl_geom_t g,
                   void* stream);

/* ---- attention (Attention.forward vit.py:84-88 and its autograd) -------------------------------------- */

/* qkv [n_seq*seq, 3*H*64] (act dtype, column order [3][H][64], vit.py:78) -> out [n_seq*seq, H*64],
 * lse fp32 [n_seq, H, seq] (log-sum-exp of the scaled scores).  Any seq; CUDA-core fp32 math. */
int pvrl_attn_fwd(const void* qkv, void* out, float* lse, int32_t dtype, int32_t n_seq, int32_t seq, int32_t H,
                  float scale, void* stream);
int pvrl_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int32_t dtype,
                  int32_t n_seq, int32_t seq, int32_t H, float scale, void* stream);

/* Tensor-core (tcgen05 + TMA) attention for the spatial axis: bf16, head_dim 64, seq <= 256.  Same contract as
 * pvrl_attn_fwd / pvrl_attn_bwd (qkv column order [3][H][64]; lse = log-sum-exp of the scaled scores). */
int pvrl_attn_tc_fwd(const void* qkv, void* out, float* lse, int32_t n_seq, int32_t seq, int32_t H, float scale,
                     void* stream);
int pvrl_attn_tc_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                     int32_t n_seq, int32_t seq, int32_t H, float scale, void* stream);

/* Development aid.  With PVRL_SP_TRACE=1 in the environment the persistent spatial-attention kernels record clock64()
 * stamps of CTA 0 at their phase boundaries ([20 warps][16 problems][8 slots] int64); this copies them to host_out and
 * clears the buffer.  Returns the number of values written, 0 when tracing is off. */
int pvrl_debug_sp_trace(long long* host_out);

/* ---- head + similarity + loss (vit.py:300-307, train_net.py:153-162) ----------------------------------- */

/* y[m, j] = sum_k x[m,k] w[j,k] + b[j]   fp32, small M (head 768->512, vit.py:301) */
int pvrl_linear_small_fwd(const float* x, const float* w, const float* b, float* y, int32_t M, int32_t K, int32_t N,
                          void* stream);
int pvrl_linear_small_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db, int32_t M,
                          int32_t K, int32_t N, void* stream);
/* y = x / ||x||_2 (rows), norms saved.  vit.py:302 */
int pvrl_l2norm_fwd(const float* x, float* y, float* norms, int32_t M, int32_t C, void* stream);
int pvrl_l2norm_bwd(const float* y, const float* norms, const float* dy, float* dx, int32_t M, int32_t C,
                    void* stream);
/* logits[m, c] = emb[m,:] . label[c,:] * inv_temp   (vit.py:307; label rows pre-normalised, vit.py:435-440) */
int pvrl_sim_logits_fwd(const float* emb, const float* label, float* logits, int32_t M, int32_t C, int32_t K,
                        float inv_temp, void* stream);
/* demb[m,:] += sum_c dlogits[m,c] label[c,:] * inv_temp */
int pvrl_sim_logits_bwd(const float* dlogits, const float* label, float* demb, int32_t M, int32_t C, int32_t K,
                        float inv_temp, void* stream);
/* Pre-training loss, train_net.py:153-162: teacher = renorm(top-k(softmax(teacher_logits))),
 * loss = KLDiv(log_softmax(pred), teacher, batchmean); dpred = (softmax(pred)*sum(t) - t) / M * gscale.
 * row_loss fp32 [M] (sum to get the loss), teacher_out optional fp32 [M,K]. */
int pvrl_kl_topk_loss(const float* pred, const float* teacher_logits, float* row_loss, float* dpred,
                      float* teacher_out, int32_t M, int32_t K, int32_t topk, float gscale, void* stream);
/* eval-mode softmax over rows (vit.py:355-356) */
int pvrl_softmax_rows(const float* x, float* y, int32_t M, int32_t K, void* stream);

/* ---- clip-order ("diffusion") transformer of the pre-training branch, fp32 -------------------------------------
 * Replaces the op-by-op eager execution of DiffusionTransformer (tfm_model.py:70-289; call site vit.py:330):
 * ResidualAttentionBlock :32-53 = LayerNorm -> nn.MultiheadAttention(8 heads of 64, key_padding_mask) -> residual ->
 * LayerNorm -> Linear -> QuickGELU -> Linear -> residual, run per denoising level over M = B*S rows (a few dozen).
 * Rows are b-major: token (b, s) is row b*S + s.  Every pointer is fp32 unless noted. */

/* y[M,N] = (resid ? resid : 0) + prologue(x)[M,K] W[N,K]^T + bias.  x_mode 0: x; 1: LayerNorm(x) with ln_w/ln_b/eps
 * (also writes xhat[M,K] and rstd[M] when non-NULL, for pvrl_ot_ln_bwd / pvrl_ot_linear_dw); 2: QuickGELU(x)
 * (tfm_model.py:27-29).  act_out (NULL or [M,N]) also receives QuickGELU(y): the input of the next Linear and of its dW,
 * computed once.  K % 128 == 0, K <= 2048. */
int pvrl_ot_linear_fwd(const float* x, int32_t x_mode, const float* ln_w, const float* ln_b, float eps, float* xhat_out,
                       float* rstd_out, const float* W, const float* bias, const float* resid, float* y, float* act_out,
                       int32_t M, int32_t N, int32_t K, void* stream);
/* dA[M,K] = dY[M,N] W[N,K], times QuickGELU'(pre[M,K]) when pre != NULL. */
int pvrl_ot_linear_dx(const float* dY, const float* W, const float* pre, float* dA, int32_t M, int32_t N, int32_t K,
                      void* stream);
/* dW[N,K] += dY[M,N]^T a[M,K]; db[N] += colsum(dY) (db may be NULL).  a_mode 0: a = A; 1: a = A*ln_w + ln_b (A = xhat
 * saved by pvrl_ot_linear_fwd); 2: a = QuickGELU(A). */
int pvrl_ot_linear_dw(const float* dY, const float* A, int32_t a_mode, const float* ln_w, const float* ln_b, float* dW,
                      float* db, int32_t M, int32_t N, int32_t K, void* stream);
/* LayerNorm backward on saved xhat / rstd: dh[M,C] += dx, dw[C] += sum dA*xhat, db[C] += sum dA. */
int pvrl_ot_ln_bwd(const float* dA, const float* xhat, const float* rstd, const float* w, float* dh, float* dw,
                   float* db, int32_t M, int32_t C, void* stream);
/* Multi-head attention over S <= 16 tokens per sequence, head_dim 64: qkv[M, 3*H*64] ([q | k | v], heads contiguous, as
 * nn.MultiheadAttention.in_proj), keys s >= pad_start[b] masked (pad_start int64[B], NULL = no padding);
 * probs[B,H,S,S] saved for the backward; o[M, H*64]. */
int pvrl_ot_attn_fwd(const float* qkv, const int64_t* pad_start, float* probs, float* o, int32_t B, int32_t S, int32_t H,
                     void* stream);
int pvrl_ot_attn_bwd(const float* qkv, const float* probs, const float* dO, float* dqkv, int32_t B, int32_t S, int32_t H,
                     void* stream);
/* Level input (tfm_model.py:171-191, ennoise :291-302): h[(b,s)] = in + type_emb[s == mask_b] + pos_emb[s] + tvec with
 * in = ca*src[b] + cb*noise[b] at the mask position, pad_emb at s >= pad_start[b], video[(b,s)] elsewhere. */
int pvrl_ot_embed_fwd(const float* video, const float* src, const float* noise, float ca, float cb,
                      const int64_t* mask_inds, const int64_t* pad_start, const float* type_w, const float* pos_w,
                      const float* pad_w, const float* tvec, float* h, int32_t B, int32_t S, int32_t C, void* stream);
/* Its backward: dvideo[M,C] / dtype[2,C] / dpos[S,C] / dpad[C] are accumulated (+=), dtvec[C] is written. */
int pvrl_ot_embed_bwd(const float* dh, const int64_t* mask_inds, const int64_t* pad_start, float* dvideo, float* dtype,
                      float* dpos, float* dpad, float* dtvec, int32_t B, int32_t S, int32_t C, void* stream);

/* ---- optimizer step over flat buffers ------------------------------------------------------ */
/* Replaces torch.optim.AdamW / Adam / SGD as constructed by lib/models/optimizer.py:90-114 and the
 * optimizer.step(); optimizer.zero_grad() pair of tools/train_net.py:176-192.  One call per parameter group
 * (optimizer.py:35-84: weight decay and lr multiplier differ per group); p, g and the state arrays are the group's
 * slices of flat fp32 buffers laid out identically (same offset from a 16-byte boundary).
 * lr_dev / step_dev are DEVICE scalars: the learning rate of this iteration (lr_policy.get_lr_at_epoch, set_lr) and
 * the number of the step being taken (1 for the first), so a CUDA graph of the step can be replayed unchanged;
 * pvrl_optim_tick increments the counter once per step, before the group calls.
 * g is multiplied by grad_scale before use (1/world after a SUM all-reduce, 1/micro-steps under accumulation) and,
 * with zero_grad != 0, cleared for the next backward (whose dW kernels accumulate). */
int pvrl_optim_tick(float* step_dev, void* stream);
/* decoupled != 0: AdamW (p *= 1 - lr*wd); decoupled == 0: Adam with the L2 term added to the gradient.  The betas are
 * doubles because torch derives 1 - beta and the bias corrections from the Python doubles before rounding to fp32. */
int pvrl_adam_flat(float* p, float* g, float* m, float* v, int64_t n, const float* lr_dev, const float* step_dev,
                   float lr_mult, double beta1, double beta2, float eps, float weight_decay, int32_t decoupled,
                   float grad_scale, int32_t zero_grad, void* stream);
/* dst[i] = (dst[i] + sum_{s < n_src} src[s * stride + i]) * scale, i < n (fp32; vectorised when dst, src are 16-byte aligned):
 * the local reduction of the copy-engine gradient exchange -- the peers' copies of this rank's chunk of the flat
 * gradient buffer (pulled over NVLink by the DMA engines, no SMs) are summed into it and averaged.  Replaces the
 * reduction half of the NCCL all-reduce behind DistributedDataParallel, reference lib/models/build.py:49-53. */
int pvrl_reduce_chunks(float* dst, const float* src, int32_t n_src, int64_t stride, int64_t n, float scale, void* stream);
/* torch.optim.SGD semantics: g += wd*p; buf = g on step 1 else momentum*buf + (1-dampening)*g;
 * update = g + momentum*buf (nesterov) or buf; p -= lr*update.  momentum == 0 ignores buf's contents. */
int pvrl_sgd_flat(float* p, float* g, float* buf, int64_t n, const float* lr_dev, const float* step_dev, float lr_mult,
                  float momentum, float dampening, int32_t nesterov, float weight_decay, float grad_scale,
                  int32_t zero_grad, void* stream);

/* ---- MViTv2 video encoder (BASELINE config 5; reference lib/models/slowfast_mvit/) --------------
 * Tokens of a clip are [1 + T*H*W, width] rows, cls first, grid order (t h w) with w fastest (mvit.py:338-360).
 * The encoder's Linear layers (qkv, proj, skip proj, MLP, Conv3d stem as im2col) go through pvrl_gemm_bf16. */

/* nn.LayerNorm over the last axis of [M, D], any D <= 1024 (attention.py:239-280 norm_q/k/v over 96; :530,:556 norm1 /
 * norm2 over 96 .. 768; mvit.py:401 final norm).  x, y: bf16 or fp32; stats [M, 2] = (mean, rstd) for the backward.
 * Backward: dx (dtype of x) is written; dw / db (fp32, may be NULL) are ACCUMULATED (+=, atomics). */
int pvrl_ln_any_fwd(const void* x, int32_t x_dtype, const float* w, const float* b, void* y, int32_t y_dtype, float* stats,
                    int32_t M, int32_t D, float eps, void* stream);
int pvrl_ln_any_bwd(const void* dy, int32_t dy_dtype, const void* x, int32_t x_dtype, const float* w, const float* stats,
                    void* dx, float* dw, float* db, int32_t M, int32_t D, void* stream);

/* Geometry of a 3-D pooling window over a clip's token grid. out[d] = floor((in[d] + 2 pad[d] - kernel[d]) / stride[d]) + 1. */
typedef struct pvrl_pool3d {
  int32_t B;         /* clips                                                                         */
  int32_t heads;     /* heads sharing the pooling weights (1 for the skip max-pool / the stem)        */
  int32_t C;         /* channels per head (pool3d: <= 128), token width (maxpool3d)                   */
  int32_t T, H, W;   /* input grid                                                                    */
  int32_t kernel[3], stride[3], pad[3], out[3];
  int64_t ld;        /* pool3d: row pitch (elements) of the token-major input [B, 1 + T*H*W, ld]      */
} pvrl_pool3d_t;

/* attention_pool (attention.py:14-48) for one of Q / K / V: `in` points at that tensor's first column inside the qkv
 * GEMM output [B, 1 + T*H*W, ld = 3 * heads * C] (head h = columns [h*C, h*C + C)); depth-wise Conv3d(C, C, kernel,
 * stride, pad, groups = C, bias = False) with weights w [C, kt*kh*kw] shared by all heads, the cls row bypasses the
 * convolution.  out [B, heads, 1 + To*Ho*Wo, C] (same dtype).  w == NULL: no pooling, only the head-major re-layout.
 * Backward: din gets the same strided layout as `in` (every element of the slice is written); dw [C, 27] (may be
 * NULL) is ACCUMULATED and needs the forward input `in` and a 3 x 3 x 3 kernel. */
int pvrl_pool3d_fwd(const void* in, const float* w, void* out, int32_t dtype, const pvrl_pool3d_t* p, void* stream);
int pvrl_pool3d_bwd(const void* dout, const void* in, const float* w, void* din, float* dw, int32_t dtype,
                    const pvrl_pool3d_t* p, void* stream);

/* MaxPool3d of the skip path (attention.py:521-543): x [B, 1 + T*H*W, C] -> y [B, 1 + To*Ho*Wo, C], cls row copied;
 * arg [B, To*Ho*Wo, C] = winning input token of every window (first maximum in scan order, as ATen).
 * Backward: dx fp32, ZERO-INITIALISED by the caller, receives dy through arg (windows overlap: atomics). */
int pvrl_maxpool3d_fwd(const void* x, void* y, int32_t* arg, int32_t dtype, const pvrl_pool3d_t* p, void* stream);
int pvrl_maxpool3d_bwd(const void* dy, const int32_t* arg, float* dx, int32_t dtype, const pvrl_pool3d_t* p, void* stream);

/* Rows of the Conv3d patch stem (stem_helper.py:290-322) for the GEMM: frames fp32 [B, Cin, T, H, W] ->
 * out [B * To*Ho*Wo, Kpad], column ((c*kt + dt)*kh + dh)*kw + dw, zero-padded to Kpad and outside the clip. */
int pvrl_im2col3d(const float* frames, void* out, int32_t out_dtype, int32_t Cin, int32_t Kpad, const pvrl_pool3d_t* p,
                  void* stream);

/* Pooled attention with decomposed relative-position bias and residual pooling (attention.py:51-159, :360-411).
 * q [B, heads, Nq, C], k / v [B, heads, Nk, C] (row 0 = cls; C = 96), bq [B, heads, Nq - 1, Kt + Kh + Kw] fp32 = each non-cls
 * query's dot products with the rows of the relative-position tables that its position selects:
 *   score(i, j) = scale * q_i . k_j + [i > 0 and j > 0] (bq[i-1, kt] + bq[i-1, Kt + kh] + bq[i-1, Kt + Kh + kw]),
 *   j - 1 = (kt * Kh + kh) * Kw + kw;    out[b, i, h*C ..] = softmax_j(score) v + [i > 0 and residual_pooling] q_i.
 * out [B, Nq, heads * C] (what the projection Linear reads), lse [B, heads, Nq] fp32.
 * Backward (recomputes the probabilities from q, k, bq and lse; the forward output is not needed): dq (layout / dtype of
 * q), dbq (layout of bq) and delta [B, heads, Nq] = sum_j p_ij dP_ij are written; dk / dv are fp32
 * [B, heads, Nk, C], ZERO-INITIALISED by the caller (atomics over query slices). */
typedef struct pvrl_pooled_attn {
  int32_t B, heads, Nq, Nk, C;
  int32_t Kt, Kh, Kw;          /* key grid: Nk = 1 + Kt*Kh*Kw */
  float scale;
  int32_t residual_pooling;
} pvrl_pooled_attn_t;
int pvrl_pooled_attn_fwd(const void* q, const void* k, const void* v, const float* bq, void* out, float* lse, int32_t dtype,
                         const pvrl_pooled_attn_t* a, void* stream);
int pvrl_pooled_attn_bwd(const void* q, const void* k, const void* v, const float* bq, const void* dout, const float* lse,
                         void* dq, float* dk, float* dv, float* dbq, float* delta, int32_t dtype,
                         const pvrl_pooled_attn_t* a, void* stream);

/* ---- misc ----------------------------------------------------------------------------------------------- */
const char* pvrl_last_error(void);
int pvrl_abi_version(void);
/* number of kernels this library has launched in this process (bench.py "gpu_launches") */
int64_t pvrl_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PVRL_H_ */
