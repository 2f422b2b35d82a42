/*
 * simvg_b200 — C ABI of libsimvg_b200.so (sm_100a kernels for SimVG's vision-language fusion train step).
 *
 * The reference (Dmmm1997/SimVG) has no FFI: its hot path is PyTorch eager behind the mmcv registry
 * (simvg/models/builder.py:4-36).  Each entry point below therefore names the reference *Python* op
 * sequence it replaces (file:line under /root/reference) — see INTEGRATION.md for the binding a maintainer
 * adds on the reference side (ctypes, from simvg/models/vis_encs/beit/beit3_base.py etc.).
 *
 * Conventions (all entry points):
 *   - return 0 on success, <0 on error; the message is in simvgb_last_error() (thread-local).
 *   - plain device pointers + sizes; the library never allocates, frees or synchronises; every launch goes
 *     to the caller's stream (a cudaStream_t passed as void*).
 *   - pointers must be 16-byte aligned; bf16 = raw uint16 storage of __nv_bfloat16; row-major.
 *   - no C++ exceptions cross this boundary.
 */
#ifndef SIMVG_B200_H
#define SIMVG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIMVGB_VERSION 200

int simvgb_version(void);
const char* simvgb_last_error(void);
/* 0 iff `device` is a compute-capability-10.x GPU (tcgen05/TMEM present). */
int simvgb_device_check(int device);

/* ------------------------------------------------------------------------------------------------
 * GEMM  C[M,N] = epilogue( A(M,K) · B(N,K)^T ), bf16 operands, fp32 accumulation in TMEM (tcgen05.mma).
 * Replaces every nn.Linear on the path: torchscale q/k/v/out_proj + fc1/fc2
 * (beit3_base.py:57-63,112-121 -> torchscale MultiheadAttention / FeedForwardNetwork), Conv2d patch-embed
 * (beit3_base.py:417-424) and 1x1 input_proj (tgqs_kd_detr_head.py:74,377), and their autograd backward
 * (dgrad + wgrad; simvg/apis/train.py:80).
 * ------------------------------------------------------------------------------------------------ */
enum simvgb_gemm_epilogue {
  SIMVGB_EPI_BF16 = 0,     /* out_bf16 = (acc + bias[n]) * (n < scale_cols ? scale : 1)                 */
  SIMVGB_EPI_GELU = 1,     /* out_bf16 = u = acc + bias ; out2_bf16 = gelu_erf(u)   (fc1, A.5)           */
  SIMVGB_EPI_RESID = 2,    /* out_f32 = res_f32 + row_scale[row / rows_per_scale] * (acc + bias)         */
  SIMVGB_EPI_F32 = 3,      /* out_f32 = acc + bias            (accumulate=1: out_f32 += ...)             */
  SIMVGB_EPI_ATOMIC = 4    /* atomicAdd(out_f32, acc)         (split-K weight gradients)                 */
};

typedef struct simvgb_gemm_args {
  int32_t M, N, K;
  /* operand storage: 0 = K contiguous ([M,K] / [N,K] row-major), 1 = M/N contiguous ([K,M] / [K,N]) */
  int32_t a_mn_major, b_mn_major;
  int64_t lda, ldb;          /* leading dimension of the stored matrices, in elements */
  const void* A;             /* bf16 */
  const void* B;             /* bf16 */
  int32_t epilogue;          /* enum simvgb_gemm_epilogue */
  int32_t k_splits;          /* >= 1; > 1 only with SIMVGB_EPI_ATOMIC */
  const float* bias;         /* [N] or NULL */
  void* out_bf16;            /* [M, ldo] */
  void* out2_bf16;           /* [M, ldo] (GELU) */
  float* out_f32;            /* [M, ldo] */
  const float* res_f32;      /* [M, ldo] (RESID) */
  int64_t ldo;
  float scale;               /* BF16 epilogue column scale */
  int32_t scale_cols;
  const float* row_scale;    /* per row-group scale (DropPath mask / keep_prob) or NULL */
  int32_t rows_per_scale;
  int32_t accumulate;        /* F32 epilogue: add to existing contents */
} simvgb_gemm_args;

int simvgb_gemm(const simvgb_gemm_args* args, void* stream);
/* Two independent problems in ONE persistent launch: SimVG's multiway layers run every projection twice (vision-token
 * expert, text-token expert: torchscale MultiwayNetwork); the text problem is ~80x smaller and, launched on
 * its own, is pure launch/ramp latency.  Equivalent to simvgb_gemm(a) followed by simvgb_gemm(b). */
int simvgb_gemm_pair(const simvgb_gemm_args* a, const simvgb_gemm_args* b, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused multiway self-attention (flash-style; S, O, dK, dV, dQ accumulators in TMEM).
 * Replaces torchscale MultiheadAttention.forward's bmm -> masked_fill(key_padding_mask, -inf) ->
 * softmax(dtype=fp32) -> bmm (called at beit3_base.py:137-145) and its backward.
 * Vision and text tokens live in separate token-major buffers (multiway experts A / B); per sample the
 * attended sequence is [Lv vision | Lt text].  q must already be scaled by head_dim^-0.5.
 * ------------------------------------------------------------------------------------------------ */
typedef struct simvgb_attn_args {
  int32_t B, H, Lv, Lt, head_dim;   /* head_dim must be 64 */
  const void* qkv_v;      /* bf16 [B*Lv, 3*H*64]   row = q | k | v                       */
  const void* qkv_t;      /* bf16 [B*Lt, 3*H*64]                                          */
  const void* text_pad;   /* uint8 [B, Lt], 1 = padded text token (masked key), or NULL    */
  void* out_v;            /* bf16 [B*Lv, H*64]   fwd: written; bwd: read                  */
  void* out_t;            /* bf16 [B*Lt, H*64]                                             */
  float* lse;             /* fp32 [B, H, simvgb_attn_lse_stride(Lv, Lt)] log2-domain       */
  /* backward only */
  const void* dout_v;     /* bf16 [B*Lv, H*64] */
  const void* dout_t;     /* bf16 [B*Lt, H*64] */
  void* dqkv_v;           /* bf16 [B*Lv, 3*H*64] out: gradient w.r.t. the *unscaled* q, k, v */
  void* dqkv_t;           /* bf16 [B*Lt, 3*H*64] */
  float* delta;           /* fp32 workspace [B, H, lse_stride] */
  float* dq_acc_v;        /* fp32 workspace [B*Lv, H*64] (zeroed by the call) */
  float* dq_acc_t;        /* fp32 workspace [B*Lt, H*64] */
  float q_scale;          /* head_dim^-0.5, applied to dq on the way out */
  int32_t delta_ready;    /* bwd: delta was already produced (simvgb_ln_bwd mode 1 with .delta set; non-token slots zero) */
} simvgb_attn_args;

int simvgb_attn_lse_stride(int Lv, int Lt);
/* Position of text token 0 on the virtual sequence axis the lse / delta rows are indexed by (vision token i sits at i). */
int simvgb_attn_text_offset(int Lv, int Lt);
int simvgb_attn_fwd(const simvgb_attn_args* args, void* stream);
int simvgb_attn_bwd(const simvgb_attn_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm (HBM-bound, coalesced 16/32 B per lane).  C must be a multiple of 256.
 * Replaces the multiway nn.LayerNorm calls self_attn_layer_norm / final_layer_norm / inner_attn_ln /
 * ffn_layernorm / encoder.layer_norm (beit3_base.py:41,86,136,157,228,396-397; torchscale A.4, A.5) — the
 * caller launches once per expert (vision rows with the A parameters, text rows with the B parameters).
 * ------------------------------------------------------------------------------------------------ */
/* act = 1: y = LN(gelu_erf(x)) — the FFN's activation (F.gelu(x.float()), A.5) fused in front of ffn_layernorm. */
int simvgb_ln_fwd(const void* x, int x_is_bf16, void* y, int y_is_bf16, const float* gamma, const float* beta,
                  float* mean, float* rstd, int64_t rows, int C, float eps, int act, void* stream);

/* Backward.  mode 0: residual-stream LN — dres_out = dres_in + LN'(dy), optionally also emits
 *   dyb = bf16(row_scale * dres_out) (the dY operand of the preceding out_proj / fc2 GEMMs) and
 *   dbias_prev += colsum(row_scale * dres_out) (their bias gradient).           (beit3_base.py:123-124,146-169)
 * mode 1: inner attention LN — dx(bf16) = LN'(dy).
 * mode 2: FFN LN fused with GELU backward — the LN input was gelu(u) and is recomputed from u (x is ignored, may be
 *   NULL): dx(bf16) = LN'(dy) * gelu'(u), dbias_prev += colsum(dx) (fc1 bias). */
typedef struct simvgb_ln_bwd_args {
  int32_t mode, C;
  int64_t rows;
  const void* x;          /* LN input: fp32 (mode 0) / bf16 (modes 1, 2) */
  const void* dy;         /* bf16, or fp32 if dy_is_f32 */
  int32_t dy_is_f32;
  const float* gamma;
  const float* mean;
  const float* rstd;
  float* dgamma;          /* [C] accumulated (+=) */
  float* dbeta;           /* [C] accumulated (+=) */
  const float* dres_in;   /* mode 0; NULL = zero */
  float* dres_out;        /* mode 0 (may alias dres_in) */
  void* dyb;              /* mode 0 optional bf16 [rows, C] */
  const float* row_scale; /* optional per-sample DropPath scale */
  int32_t rows_per_scale;
  float* dbias_prev;      /* optional [C] accumulated (+=) */
  void* dx;               /* modes 1, 2: bf16 [rows, C] */
  const void* u;          /* mode 2: bf16 pre-activation */
  /* mode 1, optional (NULL = off): the attention backward's delta = rowsum(O o dO) per head, written in simvgb_attn_bwd's
   * workspace layout delta[(b * H + h) * delta_stride + delta_vbase + l] for row b * delta_L + l (x = O, dx = dO). */
  float* delta;
  int32_t delta_L, delta_H, delta_stride, delta_vbase;
} simvgb_ln_bwd_args;
int simvgb_ln_bwd(const simvgb_ln_bwd_args* args, void* stream);

/* out[c] += sum_r s(r) * in[r, c]; optional out_bf16[r, c] = bf16(s(r) * in[r, c]) with s(r) = row_scale[r / rows_per_scale]
 * (bias gradients: the reduction autograd performs for nn.Linear.bias). */
int simvgb_colsum(const void* in, int in_is_bf16, float* out, void* out_bf16, const float* row_scale,
                  int rows_per_scale, int64_t rows, int C, int64_t ld, void* stream);
int simvgb_cast_bf16(const float* in, void* out_bf16, int64_t n, void* stream);

/* Embedding assembly (torchscale VisionEmbedding / TextEmbedding / PositionalEmbedding, A.6-A.7;
 * Encoder.forward_embedding beit3_base.py:317-334 and the pad-zeroing at :367). */
int simvgb_im2col_patch(const float* img, void* cols_bf16, int B, int S, int P, void* stream);
/* Same patch matrix straight from the un-normalised uint8 [B,S,S,3] (HWC) image of the dataset pipeline: mmcv.imnormalize
 * ((x - mean) / std per channel, BGR->RGB when to_rgb; simvg/datasets/pipelines/transforms.py:126-155) and the HWC->CHW
 * transpose of the collate step (simvg/datasets/utils.py:24-52) fused into the im2col.  mean / std: host float[3]. */
int simvgb_im2col_patch_u8(const void* img_u8, void* cols_bf16, int B, int S, int P, const float* mean, const float* std,
                           int to_rgb, void* stream);
int simvgb_embed_vision(const float* patch, const float* cls, const float* posA, float* xv, int B, int N, int D, void* stream);
int simvgb_embed_text(const float* table, const int64_t* ids, const void* pad_u8, const float* posB, float* xt, int B,
                      int Lt, int D, void* stream);

/* Global-norm clip + Adam(amsgrad) on flat fp32 buffers (apis/train.py:81-83; core/optimizer.py:52-68).
 * grad_sumsq: device scalar holding sum(g^2) over ALL parameters (simvgb_sumsq accumulates into it); the clip
 * coefficient min(1, max_norm / (sqrt(sumsq) + 1e-6)) is evaluated on the device — no host sync. */
int simvgb_sumsq(const float* g, int64_t n, float* out, void* stream);
/* ema (optional, may be NULL): shadow weights updated in the same pass, ema = ema_decay * ema + (1 - ema_decay) * p_new —
 * ExponentialMovingAverage.update_params (simvg/models/utils.py:148-173; decay = min(alpha, (1 + t) / (10 + t)) is the
 * caller's). */
int simvgb_adam_amsgrad(float* p, const float* g, float* m, float* v, float* vmax, int64_t n, float lr, float beta1,
                        float beta2, float eps, float weight_decay, int step, const float* grad_sumsq, float max_norm,
                        float* ema, float ema_decay, void* stream);
/* Same update with the step-dependent scalars taken from device memory: hyper = {lr, 1 - beta1^t, sqrt(1 - beta2^t),
 * ema_decay} (fp32[4]).  The launch arguments are then step-invariant, so a whole train step can be captured in a CUDA graph and the
 * host only refreshes `hyper` before each replay. */
int simvgb_adam_amsgrad_dev(float* p, const float* g, float* m, float* v, float* vmax, int64_t n, const float* hyper,
                            float beta1, float beta2, float eps, float weight_decay, const float* grad_sumsq,
                            float max_norm, float* ema, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Object-token head (fp32): fused building blocks of the DETR decoder layers, text-guided query generation and the MLP / class /
 * box heads.  Replace detrex BaseTransformerLayer / MultiheadAttention / FFN over nn.Linear / nn.LayerNorm /
 * nn.MultiheadAttention and their autograd graph (simvg/models/heads/tgqs_kd_detr_head/transformer.py:93-186,
 * tgqs_kd_detr_head.py:375-454).  All buffers fp32 row-major; `drop_u` (optional) holds uniform [0,1) samples,
 * an element is kept (and scaled by 1/(1-p)) iff u >= drop_p.  Backward entry points ACCUMULATE (+=) into every gradient buffer.
 * ------------------------------------------------------------------------------------------------ */
typedef struct simvgb_head_lin_args {
  /* forward: y[r,n] = dropout(relu?( sum_k (x[r,k] + (n < n_split ? x2[r,k] : 0)) W[n,k] + b[n] ))
   * k_splits > 1: partial sums are atomically added into a ZEROED y (bias from split 0), no relu / dropout.
   * backward: dx (+ dx2 for the columns < n_split) += dY_eff W,  dW += dY_eff^T x_in,  db += colsum(dY_eff); NULL = skip.
   * (nn.Linear forward / backward: heads/utils.py:7-46, the in / out projections of nn.MultiheadAttention and detrex FFN.) */
  const float* x;        /* [R, K] */
  const float* x2;       /* [R, K] or NULL (position embedding added to the input of the first n_split outputs) */
  const float* W;        /* [N, K] */
  const float* b;        /* [N] or NULL */
  const float* drop_u;   /* [R, N] or NULL */
  float* y;              /* [R, N] (backward: the forward output, needed for the ReLU mask) */
  const float* dy;       /* [R, N] backward */
  float* dx;             /* [R, K] */
  float* dx2;            /* [R, K] */
  float* dW;             /* [N, K] */
  float* db;             /* [N] */
  int32_t R, N, K, n_split, relu, k_splits;
  float drop_p;
} simvgb_head_lin_args;
int simvgb_head_lin_fwd(const simvgb_head_lin_args* args, void* stream);
int simvgb_head_lin_bwd(const simvgb_head_lin_args* args, void* stream);

typedef struct simvgb_head_ln_args {
  /* forward: s = a + dropout(b);  y = LayerNorm(s) * gamma + beta  (mean / rstd written).   C = 256 or 512.
   * backward: ds = LN'(dy);  da += ds;  db += ds * dropmask;  dgamma += sum dy * xhat;  dbeta += sum dy. */
  const float* a;        /* [R, C] */
  const float* b;        /* [R, C] or NULL */
  const float* drop_u;   /* [R, C] or NULL (applies to b) */
  const float* gamma;
  const float* beta;
  float* y;
  float* mean;           /* [R] forward: out, backward: in */
  float* rstd;
  const float* dy;
  float* da;
  float* db;
  float* dgamma;
  float* dbeta;
  int32_t R, C;
  float drop_p, eps;
} simvgb_head_ln_args;
int simvgb_head_lnres(const simvgb_head_ln_args* args, int backward, void* stream);

typedef struct simvgb_head_attn_args {
  /* Attention against at most 32 keys per sample (decoder self-attention over the nq queries; cross-attention of the
   * text-guided query generation over the text tokens): H heads of 32 channels, scores = scale * q.k, key padding mask,
   * softmax, dropout on the probabilities, ctx = P V.  q / k / v / ctx may be column slices (row strides ldq / ldk / ldc). */
  const float* q;        /* rows b * nq + i */
  const float* k;        /* rows b * nk + j */
  const float* v;
  const unsigned char* kpm;   /* [B, nk], 1 = ignore, or NULL */
  const float* drop_u;   /* [B, H, nq, nk] or NULL */
  float* ctx;
  float* P;              /* [B, H, nq, nk] softmax before dropout (forward: out, backward: in) */
  const float* dctx;
  float* dq;
  float* dk;
  float* dv;
  int32_t B, nq, nk, H, ldq, ldk, ldc;
  float scale, drop_p;
} simvgb_head_attn_args;
int simvgb_head_attn_small(const simvgb_head_attn_args* args, int backward, void* stream);

typedef struct simvgb_head_xattn_args {
  /* Cross-attention of nq queries against the N-token image memory with the key / value projections absorbed into the query /
   * output side: s[h,n] = (Wk_h^T q_h) . kin[b,n] + q_h . bk_h;  ctx_h = Wv_h (sum_n pd[h,n] val[b,n]) + bv_h sum_n pd[h,n].
   * E = 256, H = 8.  q is the projected (unscaled) query [B*nq, E]; kin = memory + positions, val = memory: [B, N, E]. */
  const float* q;
  const float* kin;
  const float* val;
  const float* Wk;       /* [E, E] rows = output channel (the k block of in_proj_weight) */
  const float* bk;
  const float* Wv;
  const float* bv;
  const unsigned char* kpm;   /* [B, N] or NULL */
  const float* drop_u;   /* [B*nq, H, N] or NULL */
  float* ctx;            /* [B*nq, E] */
  float* P;              /* [B*nq, H, N] */
  float* z;              /* [B*nq, H, E] */
  float* psum;           /* [B*nq, H] */
  const float* dctx;
  float* dq;
  float* dkin;
  float* dval;
  float* dWk;
  float* dbk;
  float* dWv;
  float* dbv;
  int32_t B, nq, N, E, H;
  float drop_p;
  float scale;           /* head_dim^-0.5, applied to q inside */
  float* ws;             /* scratch, >= simvgb_head_xattn_ws_floats(B, nq, N, backward) floats, 16-byte aligned */
  long long ws_floats;
} simvgb_head_xattn_args;
/* Scratch the call needs (absorbed vectors, per-key-chunk partial sums, the backward's score gradients), in floats. */
long long simvgb_head_xattn_ws_floats(int B, int nq, int N, int backward);
int simvgb_head_xattn(const simvgb_head_xattn_args* args, int backward, void* stream);

/* TMA descriptor cache counters (descriptors are cached per (pointer, shape, stride, box); diagnostics / tests). */
void simvgb_tmap_cache_stats(long long* hits, long long* misses);

/* Hungarian matching on the device (detrex HungarianMatcher + scipy.optimize.linear_sum_assignment; called from
 * simvg/core/criterion/criterion.py:226-271).  cost: fp32 [B, nq, ttot], sample b's targets are columns
 * offsets[b] .. offsets[b+1]-1 (int32 [B+1], device); every sample needs nq <= 32 and <= 32 targets.  out_q / out_t: int64
 * [B, kmax], the min(nq, n_b) assignments of sample b in ascending query order, padded with -1. */
int simvgb_hungarian(const float* cost, int B, int nq, int ttot, const int32_t* offsets, int64_t* out_q, int64_t* out_t,
                     int kmax, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SIMVG_B200_H */
