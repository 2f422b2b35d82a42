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

#define SIMVGB_VERSION 100

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
} simvgb_attn_args;

int simvgb_attn_lse_stride(int Lv, int Lt);
int simvgb_attn_fwd(const simvgb_attn_args* args, void* stream);
int simvgb_attn_bwd(const simvgb_attn_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SIMVG_B200_H */
