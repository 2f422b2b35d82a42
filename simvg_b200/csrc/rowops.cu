// HBM-bound row kernels of the encoder: LayerNorm forward / backward (with the residual-stream, GELU and bias-gradient
// work fused in), column sums, dtype casts, embedding assembly.  All of them are coalesced 16/32-byte-per-lane
// streaming kernels: each lane owns 8 consecutive columns, a row is covered by C/256 warps.
//
// Replaces (reference = eager ATen ops): the multiway LayerNorms self_attn_layer_norm / final_layer_norm / inner_attn_ln /
// ffn_layernorm / encoder.layer_norm (/root/reference/simvg/models/vis_encs/beit/beit3_base.py:41,86,228 and torchscale
// A.4/A.5), the residual adds (:123-124,151,169), gelu backward, and Encoder.forward_embedding (:317-334).
#include "common.cuh"
#include "simvg_b200.h"

namespace simvgb {

__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                            pack_bf16x2(v[6], v[7]));
}

// Sum of `val[0..N)` over the W warps covering one row group.  red: smem [rows_per_block * W * N].
template <int N>
__device__ __forceinline__ void row_reduce(float (&val)[N], float* red, int rg, int ww, int W, int lane) {
#pragma unroll
  for (int i = 0; i < N; ++i) val[i] = warp_sum(val[i]);
  if (W == 1) return;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) red[(rg * W + ww) * N + i] = val[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float s = 0.f;
    for (int w = 0; w < W; ++w) s += red[(rg * W + w) * N + i];
    val[i] = s;
  }
  __syncthreads();
}

// Rows are processed RB at a time per row group: RB independent loads in flight per thread and one pair of block
// barriers per RB rows instead of per row.
constexpr int kLnMaxRB = 4;

// ------------------------------------------------------------------------------------------------ LayerNorm forward
template <typename TIn, typename TOut, int RB, bool GELU>
__global__ void ln_fwd_kernel(const TIn* __restrict__ x, TOut* __restrict__ y, const float* __restrict__ gamma,
                              const float* __restrict__ beta, float* __restrict__ mean, float* __restrict__ rstd,
                              long long R, int C, int W, int rpb, float eps) {
  extern __shared__ float red[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rg = warp / W, ww = warp % W;
  const int col = (ww * 32 + lane) * 8;
  float g[8], bt[8];
  load8(gamma + col, g);
  load8(beta + col, bt);
  const long long rows_per_iter = (long long)gridDim.x * rpb * RB;
  const long long niter = (R + rows_per_iter - 1) / rows_per_iter;
  for (long long it = 0; it < niter; ++it) {
    const long long row0 = ((it * gridDim.x + blockIdx.x) * rpb + rg) * RB;
    float v[RB][8];
    float s[RB];
#pragma unroll
    for (int rr = 0; rr < RB; ++rr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[rr][i] = 0.f;
      if (row0 + rr < R) {
        load8(x + (row0 + rr) * C + col, v[rr]);
        if (GELU) {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[rr][i] = gelu_fwd(v[rr][i]);
        }
      }
    }
#pragma unroll
    for (int rr = 0; rr < RB; ++rr)
      s[rr] = v[rr][0] + v[rr][1] + v[rr][2] + v[rr][3] + v[rr][4] + v[rr][5] + v[rr][6] + v[rr][7];
    row_reduce<RB>(s, red, rg, ww, W, lane);
    float q[RB];
#pragma unroll
    for (int rr = 0; rr < RB; ++rr) {
      s[rr] /= C;
      q[rr] = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = v[rr][i] - s[rr]; q[rr] += d * d; }
    }
    row_reduce<RB>(q, red, rg, ww, W, lane);
#pragma unroll
    for (int rr = 0; rr < RB; ++rr) {
      const long long row = row0 + rr;
      if (row < R) {
        const float rs = rsqrtf(q[rr] / C + eps);
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = (v[rr][i] - s[rr]) * rs * g[i] + bt[i];
        store8(y + row * C + col, o);
        if (ww == 0 && lane == 0) { mean[row] = s[rr]; rstd[row] = rs; }
      }
    }
  }
}

// Warp-per-row forward: a warp owns a whole row (NCH chunks of 8 columns per lane), issues all its loads up front, keeps
// the fp32 values in registers and reduces with shuffles only — no shared memory, no block barriers, rows independent.
template <typename TIn, typename TOut, int NCH, bool GELU>
__global__ void __launch_bounds__(128) ln_fwd_warp_kernel(const TIn* __restrict__ x, TOut* __restrict__ y,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float* __restrict__ mean, float* __restrict__ rstd, long long R,
                                                          float eps) {
  constexpr int C = NCH * 256;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * 4;
  for (long long row = warp0; row < R; row += nwarps) {
    float v[NCH][8];
#pragma unroll
    for (int c = 0; c < NCH; ++c) load8(x + row * C + (c * 32 + lane) * 8, v[c]);
    float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        float2 t = make_float2(v[c][i], v[c][i + 1]);
        if (GELU) t = gelu_fwd2(t);
        v[c][i] = t.x; v[c][i + 1] = t.y;
        s2 = add2(s2, t);
      }
    }
    const float mu = warp_sum(s2.x + s2.y) * (1.0f / C);
    float2 q2 = make_float2(0.f, 0.f);
    const float2 nmu = splat2(-mu);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) { const float2 d = add2(make_float2(v[c][i], v[c][i + 1]), nmu); q2 = fma2(d, d, q2); }
    }
    const float rs = rsqrtf(warp_sum(q2.x + q2.y) * (1.0f / C) + eps);
    const float2 rs2 = splat2(rs), nmurs = splat2(-mu * rs);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = (c * 32 + lane) * 8;
      float g[8], bt[8], o[8];
      load8(gamma + col, g);
      load8(beta + col, bt);
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const float2 xh = fma2(make_float2(v[c][i], v[c][i + 1]), rs2, nmurs);
        const float2 r = fma2(xh, make_float2(g[i], g[i + 1]), make_float2(bt[i], bt[i + 1]));
        o[i] = r.x; o[i + 1] = r.y;
      }
      store8(y + row * C + col, o);
    }
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
  }
}

template <typename TIn, typename TOut, int NCH>
static void launch_ln_fwd_warp(const void* x, void* y, const float* gamma, const float* beta, float* mean, float* rstd,
                               long long R, float eps, int act, cudaStream_t s) {
  long long blocks = (R + 3) / 4;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (act) ln_fwd_warp_kernel<TIn, TOut, NCH, true><<<(unsigned)blocks, 128, 0, s>>>((const TIn*)x, (TOut*)y, gamma, beta, mean, rstd, R, eps);
  else ln_fwd_warp_kernel<TIn, TOut, NCH, false><<<(unsigned)blocks, 128, 0, s>>>((const TIn*)x, (TOut*)y, gamma, beta, mean, rstd, R, eps);
}

template <typename TIn, typename TOut>
static bool dispatch_ln_fwd_warp(int C, const void* x, void* y, const float* gamma, const float* beta, float* mean,
                                 float* rstd, long long R, float eps, int act, cudaStream_t s) {
  switch (C) {
    case 256: launch_ln_fwd_warp<TIn, TOut, 1>(x, y, gamma, beta, mean, rstd, R, eps, act, s); return true;
    case 768: launch_ln_fwd_warp<TIn, TOut, 3>(x, y, gamma, beta, mean, rstd, R, eps, act, s); return true;
    case 1024: launch_ln_fwd_warp<TIn, TOut, 4>(x, y, gamma, beta, mean, rstd, R, eps, act, s); return true;
    case 3072: launch_ln_fwd_warp<TIn, TOut, 12>(x, y, gamma, beta, mean, rstd, R, eps, act, s); return true;
    case 4096: launch_ln_fwd_warp<TIn, TOut, 16>(x, y, gamma, beta, mean, rstd, R, eps, act, s); return true;
    default: return false;   // other widths: generic multi-warp-per-row kernel below
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm backward
// mode 0 (residual-stream LN): dres_out = dres_in + LN'(dy);  optional dyb = bf16(row_scale * dres_out) and
//                              dbias_prev += colsum(row_scale * dres_out)  (out_proj / fc2 bias gradient of the sub-layer
//                              whose output joined the stream at this point)
// mode 1 (inner attention LN): dx (bf16) = LN'(dy)
// mode 2 (FFN LN + GELU):      the LN input was gelu(u) (recomputed here from u, never stored);
//                              du (bf16) = LN'(dy) * gelu'(u);  dbias_prev += colsum(du)  (fc1 bias gradient)
struct LnBwdParams {
  const void* x;         // LN input: fp32 (mode 0) / bf16 (modes 1,2)
  const void* dy;        // bf16, or fp32 when dy_f32
  int dy_f32;
  const float* gamma;
  const float* mean;
  const float* rstd;
  float* dgamma;
  float* dbeta;
  const float* dres_in;  // mode 0, may be null (treated as zero)
  float* dres_out;       // mode 0
  bf16* dyb;             // mode 0 optional
  const float* row_scale;
  int rows_per_scale;
  float* dbias_prev;     // optional (modes 0, 2)
  bf16* dx;              // modes 1, 2
  const bf16* u;         // mode 2
  long long R;
  int C, W, rpb, mode;
  // mode 1 only (optional): delta[(b * H + h) * stride + vbase + l] = sum over head h's 64 columns of x * bf16(dx) for row b*L + l
  // — the attention backward's rowsum(O o dO), produced here where O and dO are already in registers
  float* delta;
  int delta_L, delta_H, delta_stride, delta_vbase;
};

// Raw (still packed) operands of one row chunk: loaded one row ahead so the global-load latency of row i+1 overlaps the
// reduction / barriers / stores of row i.
template <int MODE, bool DYF32>
struct LnRaw {
  uint4 x0, dy0;
  uint4 x1;    // only meaningful for MODE 0 (fp32 x)
  uint4 dy1;   // only meaningful for DYF32
  uint4 u0;    // only meaningful for MODE 2
  float mu, rs;
};

__device__ __forceinline__ void unpack8(const uint4& a, const uint4& b, bool is_f32, float (&v)[8]) {
  if (is_f32) {
    v[0] = __uint_as_float(a.x); v[1] = __uint_as_float(a.y); v[2] = __uint_as_float(a.z); v[3] = __uint_as_float(a.w);
    v[4] = __uint_as_float(b.x); v[5] = __uint_as_float(b.y); v[6] = __uint_as_float(b.z); v[7] = __uint_as_float(b.w);
  } else {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
}

template <int MODE, bool DYF32>
__device__ __forceinline__ void ln_load_raw(const LnBwdParams& p, long long row, int col, LnRaw<MODE, DYF32>& r) {
  const uint4 z = make_uint4(0, 0, 0, 0);
  r.x0 = r.dy0 = z;
  if (MODE == 0) r.x1 = z;
  if (DYF32) r.dy1 = z;
  if (MODE == 2) r.u0 = z;
  r.mu = 0.f;
  r.rs = 0.f;
  if (row >= p.R) return;
  const long long off = row * p.C + col;
  if (MODE == 0) {
    const uint4* px = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.x) + off);
    r.x0 = __ldg(px);
    r.x1 = __ldg(px + 1);
  } else if (MODE == 1) {
    r.x0 = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.x) + off));
  }   // MODE 2: the LN input is gelu(u), recomputed from u
  if (DYF32) {
    const uint4* pd = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.dy) + off);
    r.dy0 = __ldg(pd);
    r.dy1 = __ldg(pd + 1);
  } else {
    r.dy0 = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.dy) + off));
  }
  if (MODE == 2) r.u0 = __ldg(reinterpret_cast<const uint4*>(p.u + off));
  r.mu = __ldg(p.mean + row);
  r.rs = __ldg(p.rstd + row);
}

template <int MODE, bool DYF32, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) ln_bwd_kernel(const LnBwdParams p) {
  extern __shared__ float red[];
  const int W = p.W, C = p.C, rpb = p.rpb;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rg = warp / W, ww = warp % W;
  const int col = (ww * 32 + lane) * 8;
  float g[8];
  load8(p.gamma + col, g);
  float acc_g[8] = {0, 0, 0, 0, 0, 0, 0, 0}, acc_b[8] = {0, 0, 0, 0, 0, 0, 0, 0}, acc_p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long rows_per_iter = (long long)gridDim.x * rpb;
  const long long niter = (p.R + rows_per_iter - 1) / rows_per_iter;
  LnRaw<MODE, DYF32> cur, nxt;
  ln_load_raw<MODE, DYF32>(p, (long long)blockIdx.x * rpb + rg, col, cur);
  for (long long it = 0; it < niter; ++it) {
    const long long row = (it * gridDim.x + blockIdx.x) * rpb + rg;
    ln_load_raw<MODE, DYF32>(p, it + 1 < niter ? row + rows_per_iter : p.R, col, nxt);   // prefetch (row >= R loads nothing)
    const bool ok = row < p.R;
    float xh[8], dyg[8], gg[8];   // gg: gelu'(u) (MODE 2)
    if (MODE == 2) {
      unpack8(cur.u0, cur.u0, false, xh);
#pragma unroll
      for (int i = 0; i < 8; ++i) { float gv; gg[i] = gelu_fwd_grad(xh[i], gv); xh[i] = gv; }
    } else {
      unpack8(cur.x0, MODE == 0 ? cur.x1 : cur.x0, MODE == 0, xh);
    }
    unpack8(cur.dy0, DYF32 ? cur.dy1 : cur.dy0, DYF32, dyg);
    const float rs = cur.rs;
    float s[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      xh[i] = (xh[i] - cur.mu) * rs;
      acc_g[i] += dyg[i] * xh[i];
      acc_b[i] += dyg[i];
      dyg[i] *= g[i];
      s[0] += dyg[i];
      s[1] += dyg[i] * xh[i];
    }
    row_reduce<2>(s, red, rg, ww, W, lane);
    if (ok) {
      const float c1 = s[0] / C, c2 = s[1] / C;
      float dx[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) dx[i] = rs * (dyg[i] - c1 - xh[i] * c2);
      if (MODE == 0) {
        float dr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (p.dres_in != nullptr) load8(p.dres_in + row * C + col, dr);
#pragma unroll
        for (int i = 0; i < 8; ++i) dr[i] += dx[i];
        store8(p.dres_out + row * C + col, dr);
        if (p.dyb != nullptr || p.dbias_prev != nullptr) {
          const float sc = p.row_scale != nullptr ? __ldg(p.row_scale + (unsigned)row / (unsigned)p.rows_per_scale) : 1.0f;
#pragma unroll
          for (int i = 0; i < 8; ++i) { dr[i] *= sc; acc_p[i] += dr[i]; }
          if (p.dyb != nullptr) store8(p.dyb + row * C + col, dr);
        }
      } else if (MODE == 1) {
        store8(p.dx + row * C + col, dx);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) { dx[i] *= gg[i]; acc_p[i] += dx[i]; }
        store8(p.dx + row * C + col, dx);
      }
    }
    cur = nxt;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    atomicAdd(p.dgamma + col + i, acc_g[i]);
    atomicAdd(p.dbeta + col + i, acc_b[i]);
    if (MODE != 1 && p.dbias_prev != nullptr) atomicAdd(p.dbias_prev + col + i, acc_p[i]);
  }
}

// Warp-per-row backward for narrow rows (C <= 1024, i.e. the residual-stream and inner-attention LayerNorms): a warp owns
// whole rows, all loads of a row are in flight at once, row reductions are shuffles only, column accumulators live in
// registers and are combined across the block's 4 warps through shared memory before one atomic per column per block.
template <int MODE, bool DYF32, int NCH>
__global__ void __launch_bounds__(128) ln_bwd_warp_kernel(const LnBwdParams p) {
  constexpr int C = NCH * 256;
  __shared__ float sacc[3][C];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 3 * C; i += 128) (&sacc[0][0])[i] = 0.f;
  __syncthreads();
  float g[NCH][8], acc_g[NCH][8], acc_b[NCH][8], acc_p[NCH][8];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    load8(p.gamma + (c * 32 + lane) * 8, g[c]);
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc_g[c][i] = 0.f; acc_b[c][i] = 0.f; acc_p[c][i] = 0.f; }
  }
  const bool want_p = MODE == 0 && (p.dyb != nullptr || p.dbias_prev != nullptr);
  const long long nwarps = (long long)gridDim.x * 4;
  for (long long row = (long long)blockIdx.x * 4 + warp; row < p.R; row += nwarps) {
    float xh[NCH][8], dyg[NCH][8], dr[NCH][8];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const long long off = row * C + (c * 32 + lane) * 8;
      if (MODE == 0) load8(reinterpret_cast<const float*>(p.x) + off, xh[c]);
      else load8(reinterpret_cast<const bf16*>(p.x) + off, xh[c]);
      if (DYF32) load8(reinterpret_cast<const float*>(p.dy) + off, dyg[c]);
      else load8(reinterpret_cast<const bf16*>(p.dy) + off, dyg[c]);
      if (MODE == 0) {
        if (p.dres_in != nullptr) load8(p.dres_in + off, dr[c]);
        else {
#pragma unroll
          for (int i = 0; i < 8; ++i) dr[c][i] = 0.f;
        }
      }
    }
    const float mu = __ldg(p.mean + row), rs = __ldg(p.rstd + row);
    float s1 = 0.f, s2 = 0.f;
    float xraw[MODE == 1 ? NCH : 1][8];
    if (MODE == 1) {
#pragma unroll
      for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) xraw[MODE == 1 ? c : 0][i] = xh[c][i];
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        xh[c][i] = (xh[c][i] - mu) * rs;
        acc_g[c][i] += dyg[c][i] * xh[c][i];
        acc_b[c][i] += dyg[c][i];
        dyg[c][i] *= g[c][i];
        s1 += dyg[c][i];
        s2 += dyg[c][i] * xh[c][i];
      }
    }
    const float c1 = warp_sum(s1) * (1.0f / C), c2 = warp_sum(s2) * (1.0f / C);
    const float sc = (want_p && p.row_scale != nullptr) ? __ldg(p.row_scale + (unsigned)row / (unsigned)p.rows_per_scale) : 1.0f;
    long long delta_base = 0;
    if (MODE == 1 && p.delta != nullptr) {      // one 32-bit division per row (rows < 2^31), not a 64-bit one per chunk
      const unsigned r32 = (unsigned)row, bb = r32 / (unsigned)p.delta_L, l = r32 - bb * (unsigned)p.delta_L;
      delta_base = (long long)bb * p.delta_H * p.delta_stride + p.delta_vbase + l;
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const long long off = row * C + (c * 32 + lane) * 8;
      float dx[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) dx[i] = rs * (dyg[c][i] - c1 - xh[c][i] * c2);
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dr[c][i] += dx[i];
        store8(p.dres_out + off, dr[c]);
        if (want_p) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { dr[c][i] *= sc; acc_p[c][i] += dr[c][i]; }
          if (p.dyb != nullptr) store8(p.dyb + off, dr[c]);
        }
      } else {
        store8(p.dx + off, dx);
        if (MODE == 1 && p.delta != nullptr) {
          // lanes 8k .. 8k+7 hold the 64 columns of head 4c + k: three shuffles give the head's rowsum(O o dO), with dO rounded
          // to bf16 exactly as it was just stored (the attention kernels read that rounded value)
          float d = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) d += xraw[MODE == 1 ? c : 0][i] * __bfloat162float(__float2bfloat16_rn(dx[i]));
          d += __shfl_xor_sync(0xffffffffu, d, 1);
          d += __shfl_xor_sync(0xffffffffu, d, 2);
          d += __shfl_xor_sync(0xffffffffu, d, 4);
          if ((lane & 7) == 0) p.delta[delta_base + (long long)(c * 4 + (lane >> 3)) * p.delta_stride] = d;
        }
      }
    }
  }
  // combine the 4 warps' column partials in shared memory, then one atomic per column per block
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int col = (c * 32 + lane) * 8 + i;
      atomicAdd(&sacc[0][col], acc_g[c][i]);
      atomicAdd(&sacc[1][col], acc_b[c][i]);
      if (want_p) atomicAdd(&sacc[2][col], acc_p[c][i]);
    }
  }
  __syncthreads();
  for (int col = threadIdx.x; col < C; col += 128) {
    atomicAdd(p.dgamma + col, sacc[0][col]);
    atomicAdd(p.dbeta + col, sacc[1][col]);
    if (want_p && p.dbias_prev != nullptr) atomicAdd(p.dbias_prev + col, sacc[2][col]);
  }
}

// Wide rows (FFN LayerNorm + GELU backward, C = 3072 / 4096): one 128-thread block per row, each warp owns a contiguous
// quarter of the columns (NCH chunks of 8 per lane, register accumulators), one block barrier per row (double-buffered
// partials), next row's packed operands prefetched.  Blocks are small and independent, so several rows are in flight per SM.
constexpr int kWideStages = 3;   // rows of dy / u in flight per block (bulk-copied into shared memory)
template <int NCH>
__global__ void __launch_bounds__(128) ln_bwd_wide_kernel(const LnBwdParams p) {
  constexpr int C = NCH * 4 * 256;
  // dy and u rows are streamed through a kWideStages-deep shared-memory ring with cp.async.bulk (TMA): with register
  // prefetch of one row, two resident 252-register blocks keep only ~24 KB in flight per SM (3.3 TB/s); the ring keeps
  // kWideStages x 2 x C x 2 bytes per block in flight independent of the register budget.
  extern __shared__ __align__(128) uint8_t wide_smem[];
  bf16* ring = reinterpret_cast<bf16*>(wide_smem);                                   // [kWideStages][2][C]
  uint64_t* full = reinterpret_cast<uint64_t*>(wide_smem + kWideStages * 2 * C * 2);   // [kWideStages]
  __shared__ float red[2][4][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col0 = warp * NCH * 256 + lane * 8;
  float g[NCH][8], acc_g[NCH][8], acc_b[NCH][8], acc_p[NCH][8];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    load8(p.gamma + col0 + c * 256, g[c]);
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc_g[c][i] = 0.f; acc_b[c][i] = 0.f; acc_p[c][i] = 0.f; }
  }
  auto issue = [&](long long row, int stage) {   // thread 0 only
    if (row < p.R) {
      mbar_expect_tx(&full[stage], 2 * C * 2);
      bulk_g2s(ring + (stage * 2 + 0) * C, reinterpret_cast<const bf16*>(p.dy) + row * C, C * 2, &full[stage]);
      bulk_g2s(ring + (stage * 2 + 1) * C, p.u + row * C, C * 2, &full[stage]);
    }
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < kWideStages; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kWideStages; ++s) issue(blockIdx.x + (long long)s * gridDim.x, s);
  }
  float nmu = 0.f, nrs = 0.f;
  if (blockIdx.x < p.R) { nmu = __ldg(p.mean + blockIdx.x); nrs = __ldg(p.rstd + blockIdx.x); }
  int par = 0, stage = 0;
  uint32_t phase = 0;
  for (long long row = blockIdx.x; row < p.R; row += gridDim.x, par ^= 1) {
    const float mu = nmu, rs = nrs;
    if (row + gridDim.x < p.R) { nmu = __ldg(p.mean + row + gridDim.x); nrs = __ldg(p.rstd + row + gridDim.x); }
    mbar_wait(&full[stage], phase);
    uint4 cdy[NCH], cu[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      cdy[c] = *reinterpret_cast<const uint4*>(ring + (stage * 2 + 0) * C + col0 + c * 256);
      cu[c] = *reinterpret_cast<const uint4*>(ring + (stage * 2 + 1) * C + col0 + c * 256);
    }
    float xh[NCH][8], dyg[NCH][8], gg[NCH][8];
    float s1 = 0.f, s2 = 0.f;
    float2 s1v = make_float2(0.f, 0.f), s2v = make_float2(0.f, 0.f);
    const float2 rs2 = splat2(rs), nmurs = splat2(-mu * rs);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      unpack8(cu[c], cu[c], false, xh[c]);
      unpack8(cdy[c], cdy[c], false, dyg[c]);
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        float2 gv;
        const float2 gd = gelu_fwd_grad2(make_float2(xh[c][i], xh[c][i + 1]), gv);
        gg[c][i] = gd.x; gg[c][i + 1] = gd.y;
        const float2 xn = fma2(gv, rs2, nmurs);
        xh[c][i] = xn.x; xh[c][i + 1] = xn.y;
        float2 dy2 = make_float2(dyg[c][i], dyg[c][i + 1]);
        const float2 ag = fma2(dy2, xn, make_float2(acc_g[c][i], acc_g[c][i + 1]));
        acc_g[c][i] = ag.x; acc_g[c][i + 1] = ag.y;
        const float2 ab = add2(dy2, make_float2(acc_b[c][i], acc_b[c][i + 1]));
        acc_b[c][i] = ab.x; acc_b[c][i + 1] = ab.y;
        dy2 = mul2(dy2, make_float2(g[c][i], g[c][i + 1]));
        dyg[c][i] = dy2.x; dyg[c][i + 1] = dy2.y;
        s1v = add2(s1v, dy2);
        s2v = fma2(dy2, xn, s2v);
      }
    }
    s1 = s1v.x + s1v.y;
    s2 = s2v.x + s2v.y;
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) { red[par][warp][0] = s1; red[par][warp][1] = s2; }
    __syncthreads();   // also: every thread has read this ring stage
    if (threadIdx.x == 0) issue(row + (long long)kWideStages * gridDim.x, stage);
    if (++stage == kWideStages) { stage = 0; phase ^= 1; }
    const float c1 = (red[par][0][0] + red[par][1][0] + red[par][2][0] + red[par][3][0]) * (1.0f / C);
    const float c2 = (red[par][0][1] + red[par][1][1] + red[par][2][1] + red[par][3][1]) * (1.0f / C);
    const float2 nc1 = splat2(-c1), nc2 = splat2(-c2);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      float dx[8];
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        // rs * (dy*gamma - c1 - xhat * c2) * gelu'(u)
        const float2 t = fma2(make_float2(xh[c][i], xh[c][i + 1]), nc2, add2(make_float2(dyg[c][i], dyg[c][i + 1]), nc1));
        const float2 d = mul2(mul2(t, rs2), make_float2(gg[c][i], gg[c][i + 1]));
        dx[i] = d.x; dx[i + 1] = d.y;
        const float2 ap = add2(d, make_float2(acc_p[c][i], acc_p[c][i + 1]));
        acc_p[c][i] = ap.x; acc_p[c][i + 1] = ap.y;
      }
      store8(p.dx + row * C + col0 + c * 256, dx);
    }
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int col = col0 + c * 256 + i;
      atomicAdd(p.dgamma + col, acc_g[c][i]);
      atomicAdd(p.dbeta + col, acc_b[c][i]);
      if (p.dbias_prev != nullptr) atomicAdd(p.dbias_prev + col, acc_p[c][i]);
    }
  }
}

template <int MODE, bool DYF32>
static bool dispatch_ln_bwd_warp(const LnBwdParams& p, cudaStream_t st) {
  long long blocks = (p.R + 3) / 4;
  const long long cap = (long long)sm_count() * 3;
  if (blocks > cap) blocks = cap;
  switch (p.C) {
    case 256: ln_bwd_warp_kernel<MODE, DYF32, 1><<<(unsigned)blocks, 128, 0, st>>>(p); return true;
    case 768: ln_bwd_warp_kernel<MODE, DYF32, 3><<<(unsigned)blocks, 128, 0, st>>>(p); return true;
    case 1024: ln_bwd_warp_kernel<MODE, DYF32, 4><<<(unsigned)blocks, 128, 0, st>>>(p); return true;
    default: return false;
  }
}

// ------------------------------------------------------------------------------------------------ column sum / scaled cast
// out[c] += sum_r scale(r) * in[r, c];  optional ob = bf16(scale(r) * in[r, c])   (in: bf16 or fp32)
template <typename TIn>
__global__ void colsum_kernel(const TIn* __restrict__ in, float* __restrict__ out, bf16* __restrict__ ob,
                              const float* __restrict__ row_scale, int rows_per_scale, long long R, int C, int ld) {
  // block = 256 threads: 32 column-chunks (8 cols each) x 8 row lanes
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + cx) * 8;
  const bool col_ok = col < C;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long stride = (long long)gridDim.y * 8;
  for (long long row = (long long)blockIdx.y * 8 + ry; col_ok && row < R; row += 4 * stride) {
    float v[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {     // 4 independent loads in flight per thread
#pragma unroll
      for (int i = 0; i < 8; ++i) v[k][i] = 0.f;
      if (row + k * stride < R) load8(in + (row + k * stride) * ld + col, v[k]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long rk = row + k * stride;
      if (rk >= R) break;
      const float sc = row_scale != nullptr ? __ldg(row_scale + rk / rows_per_scale) : 1.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) { v[k][i] *= sc; acc[i] += v[k][i]; }
      if (ob != nullptr) store8(ob + rk * C + col, v[k]);
    }
  }
  if (out != nullptr) {
    // combine the block's 8 row lanes in shared memory, then one atomic per column per block
    __shared__ float part[8][32 * 8 + 1];
#pragma unroll
    for (int i = 0; i < 8; ++i) part[ry][cx * 8 + i] = acc[i];
    __syncthreads();
    const int c = threadIdx.x;   // 256 threads <-> 256 columns of this block
    const int gc = blockIdx.x * 256 + c;
    if (gc < C) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) t += part[r][c];
      atomicAdd(out + gc, t);
    }
  }
}

// ------------------------------------------------------------------------------------------------ casts
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long n8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float v[8];
  load8(in + i * 8, v);
  store8(out + i * 8, v);
}

// ------------------------------------------------------------------------------------------------ embedding assembly
// x_v[b, 0] = cls + posA[2];  x_v[b, 1+n] = patch[b*N+n] + posA[3+n]      (VisionEmbedding + PositionalEmbedding, A.6/A.7)
__global__ void assemble_vision_kernel(const float* __restrict__ patch, const float* __restrict__ cls,
                                       const float* __restrict__ posA, float* __restrict__ xv, int B, int N, int D) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 4 elements
  const int d4 = D / 4;
  const long long total = (long long)B * (N + 1) * d4;
  if (idx >= total) return;
  const int c = (idx % d4) * 4;
  const long long tok = idx / d4;
  const int l = tok % (N + 1);
  const long long b = tok / (N + 1);
  const float4 pos = __ldg(reinterpret_cast<const float4*>(posA + (long long)(2 + l) * D + c));
  float4 t;
  if (l == 0) t = __ldg(reinterpret_cast<const float4*>(cls + c));
  else t = __ldg(reinterpret_cast<const float4*>(patch + (b * N + (l - 1)) * D + c));
  *reinterpret_cast<float4*>(xv + tok * D + c) = make_float4(t.x + pos.x, t.y + pos.y, t.z + pos.z, t.w + pos.w);
}

// x_t[b, i] = (text_embed[ids[b,i]] + posB[2+i]) * (1 - pad[b,i])         (beit3_base.py:325-330,367)
__global__ void assemble_text_kernel(const float* __restrict__ table, const long long* __restrict__ ids,
                                     const unsigned char* __restrict__ pad, const float* __restrict__ posB,
                                     float* __restrict__ xt, int B, int Lt, int D) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int d4 = D / 4;
  const long long total = (long long)B * Lt * d4;
  if (idx >= total) return;
  const int c = (idx % d4) * 4;
  const long long tok = idx / d4;
  const int i = tok % Lt;
  const float keep = (pad != nullptr && pad[tok]) ? 0.f : 1.f;
  const float4 e = __ldg(reinterpret_cast<const float4*>(table + ids[tok] * D + c));
  const float4 pos = __ldg(reinterpret_cast<const float4*>(posB + (long long)(2 + i) * D + c));
  *reinterpret_cast<float4*>(xt + tok * D + c) =
      make_float4((e.x + pos.x) * keep, (e.y + pos.y) * keep, (e.z + pos.z) * keep, (e.w + pos.w) * keep);
}

// im2col for the stride-P patch conv: img [B,3,S,S] fp32 -> cols [B*N, 3*P*P] bf16 (channel-major, then ky, kx:
// the flattening order of Conv2d weight [D,3,P,P]).
__global__ void im2col_patch_kernel(const float* __restrict__ img, bf16* __restrict__ cols, int B, int S, int P) {
  const int G = S / P;
  const long long K = 3LL * P * P;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 8 consecutive kx
  const long long total = (long long)B * G * G * K / 8;
  if (idx >= total) return;
  const long long e = idx * 8;
  const long long row = e / K;
  const int k = e % K;
  const int ch = k / (P * P), ky = (k / P) % P, kx = k % P;
  const int n = row % (G * G);
  const long long b = row / (G * G);
  const int gy = n / G, gx = n % G;
  const float* src = img + ((b * 3 + ch) * S + (gy * P + ky)) * (long long)S + gx * P + kx;
  float v[8];
  load8(src, v);
  store8(cols + e, v);
}

// The same im2col fed from the image as the dataset pipeline holds it BEFORE `Normalize`: uint8 [B,S,S,3] (HWC, cv2 channel
// order).  mmcv.imnormalize (/root/reference/simvg/datasets/pipelines/transforms.py:126-155: optional BGR->RGB swap, then
// (x - mean) * (1 / std) in fp32) and the HWC -> CHW transpose of DefaultFormatBundle are applied on the fly, so a batch
// crosses PCIe and HBM as 1 byte per sample instead of 4.  One thread = 8 consecutive pixels of one patch row, all channels.
__global__ void im2col_patch_u8_kernel(const uint8_t* __restrict__ img, bf16* __restrict__ cols, int B, int S, int P,
                                       float m0, float m1, float m2, float i0, float i1, float i2, int to_rgb) {
  const int G = S / P;
  const long long K = 3LL * P * P;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int per_row = P * P / 8;                    // threads per patch
  const long long total = (long long)B * G * G * per_row;
  if (idx >= total) return;
  const long long row = idx / per_row;
  const int k = (idx % per_row) * 8;               // ky * P + kx
  const int ky = k / P, kx = k % P;
  const int n = row % (G * G);
  const long long b = row / (G * G);
  const int gy = n / G, gx = n % G;
  const uint8_t* src = img + ((b * S + (gy * P + ky)) * (long long)S + gx * P + kx) * 3;   // 24 contiguous bytes, 8-byte aligned
  uint2 raw[3];
  raw[0] = __ldg(reinterpret_cast<const uint2*>(src));
  raw[1] = __ldg(reinterpret_cast<const uint2*>(src) + 1);
  raw[2] = __ldg(reinterpret_cast<const uint2*>(src) + 2);
  const uint8_t* px = reinterpret_cast<const uint8_t*>(raw);
  const float mean[3] = {m0, m1, m2}, inv[3] = {i0, i1, i2};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int sc = to_rgb ? 2 - c : c;               // output channel c reads stored channel sc
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = ((float)px[3 * i + sc] - mean[c]) * inv[c];
    store8(cols + row * K + (long long)c * P * P + k, v);
  }
}

static int ln_geometry(int C, int* W, int* rpb) {
  if (C % 256 != 0 || C > 8192) return -1;
  *W = C / 256;
  int r = 8 / *W;
  if (r < 1) r = 1;
  *rpb = r;
  return 0;
}

static int ln_grid(long long R, int rows_per_block) {
  long long want = (R + rows_per_block - 1) / rows_per_block;
  long long cap = (long long)sm_count() * 4;
  return (int)(want < cap ? want : cap);
}
// Rows batched per thread: 4 when there are plenty of rows, fewer for small problems (text tokens) so the grid stays wide.
static int ln_rb(long long R, int rpb, int threads) {
  // the 4-row variant of the backward kernel needs ~160 registers/thread: only legal for blocks of <= 384 threads
  if (threads <= 384 && R >= (long long)sm_count() * 4 * rpb * 4) return 4;
  if (R >= (long long)sm_count() * 2 * rpb * 2) return 2;
  return 1;
}

}  // namespace simvgb

using namespace simvgb;

extern "C" int simvgb_ln_fwd(const void* x, int x_is_bf16, void* y, int y_is_bf16, const float* gamma, const float* beta,
                             float* mean, float* rstd, int64_t rows, int C, float eps, int act, void* stream) {
  int W, rpb;
  SIMVGB_CHECK(ln_geometry(C, &W, &rpb) == 0, "simvgb_ln_fwd: C=%d must be a multiple of 256 and <= 8192", C);
  SIMVGB_CHECK(x && y && gamma && beta && mean && rstd, "simvgb_ln_fwd: null pointer");
  if (rows <= 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  {
    bool done;
    if (!x_is_bf16 && y_is_bf16) done = dispatch_ln_fwd_warp<float, bf16>(C, x, y, gamma, beta, mean, rstd, rows, eps, act, s);
    else if (x_is_bf16 && y_is_bf16) done = dispatch_ln_fwd_warp<bf16, bf16>(C, x, y, gamma, beta, mean, rstd, rows, eps, act, s);
    else if (!x_is_bf16 && !y_is_bf16) done = dispatch_ln_fwd_warp<float, float>(C, x, y, gamma, beta, mean, rstd, rows, eps, act, s);
    else done = dispatch_ln_fwd_warp<bf16, float>(C, x, y, gamma, beta, mean, rstd, rows, eps, act, s);
    if (done) {
      SIMVGB_CUDA(cudaGetLastError());
      return 0;
    }
  }
  const int threads = 32 * W * rpb;
  const int rb = x_is_bf16 ? ln_rb(rows, rpb, threads) : 1;
  const int grid = ln_grid(rows, rpb * rb);
  const size_t sm = sizeof(float) * rpb * W * kLnMaxRB * 2;
#define SIMVGB_LN_FWD(TI, TO, RBV)                                                                                        \
  do {                                                                                                                  \
    if (act) ln_fwd_kernel<TI, TO, RBV, true><<<grid, threads, sm, s>>>((const TI*)x, (TO*)y, gamma, beta, mean, rstd, rows, C, W, rpb, eps); \
    else ln_fwd_kernel<TI, TO, RBV, false><<<grid, threads, sm, s>>>((const TI*)x, (TO*)y, gamma, beta, mean, rstd, rows, C, W, rpb, eps);   \
  } while (0)
#define SIMVGB_LN_FWD_RB(TI, TO)            \
  do {                                      \
    if (rb == 4) SIMVGB_LN_FWD(TI, TO, 4);  \
    else if (rb == 2) SIMVGB_LN_FWD(TI, TO, 2); \
    else SIMVGB_LN_FWD(TI, TO, 1);          \
  } while (0)
  if (!x_is_bf16 && y_is_bf16) SIMVGB_LN_FWD_RB(float, bf16);
  else if (x_is_bf16 && y_is_bf16) SIMVGB_LN_FWD_RB(bf16, bf16);
  else if (!x_is_bf16 && !y_is_bf16) SIMVGB_LN_FWD_RB(float, float);
  else SIMVGB_LN_FWD_RB(bf16, float);
  SIMVGB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int simvgb_ln_bwd(const simvgb_ln_bwd_args* a, void* stream) {
  SIMVGB_CHECK(a != nullptr, "simvgb_ln_bwd: null args");
  int W, rpb;
  SIMVGB_CHECK(ln_geometry(a->C, &W, &rpb) == 0, "simvgb_ln_bwd: C=%d must be a multiple of 256 and <= 8192", a->C);
  SIMVGB_CHECK(a->mode >= 0 && a->mode <= 2, "simvgb_ln_bwd: bad mode %d", a->mode);
  SIMVGB_CHECK((a->x || a->mode == 2) && a->dy && a->gamma && a->mean && a->rstd && a->dgamma && a->dbeta, "simvgb_ln_bwd: null pointer");
  SIMVGB_CHECK(a->mode != 0 || a->dres_out, "simvgb_ln_bwd: mode 0 needs dres_out");
  SIMVGB_CHECK(a->mode == 0 || a->dx, "simvgb_ln_bwd: modes 1/2 need dx");
  SIMVGB_CHECK(a->mode != 2 || a->u, "simvgb_ln_bwd: mode 2 needs u");
  SIMVGB_CHECK(a->mode == 0 || !a->dy_is_f32, "simvgb_ln_bwd: fp32 dy is only supported in mode 0");
  if (a->rows <= 0) return 0;
  LnBwdParams p;
  p.x = a->x; p.dy = a->dy; p.dy_f32 = a->dy_is_f32; p.gamma = a->gamma; p.mean = a->mean; p.rstd = a->rstd;
  p.dgamma = a->dgamma; p.dbeta = a->dbeta; p.dres_in = a->dres_in; p.dres_out = a->dres_out;
  p.dyb = reinterpret_cast<bf16*>(a->dyb); p.row_scale = a->row_scale;
  p.rows_per_scale = a->rows_per_scale > 0 ? a->rows_per_scale : 1;
  p.dbias_prev = a->dbias_prev; p.dx = reinterpret_cast<bf16*>(a->dx); p.u = reinterpret_cast<const bf16*>(a->u);
  p.R = a->rows; p.C = a->C; p.W = W; p.rpb = rpb; p.mode = a->mode;
  p.delta = a->delta; p.delta_L = a->delta_L; p.delta_H = a->delta_H; p.delta_stride = a->delta_stride; p.delta_vbase = a->delta_vbase;
  if (p.delta != nullptr) {
    SIMVGB_CHECK(a->mode == 1 && (a->C == 256 || a->C == 768 || a->C == 1024) && a->delta_H * 64 == a->C && a->delta_L > 0 &&
                     a->rows % a->delta_L == 0,
                 "simvgb_ln_bwd: delta needs mode 1, C = 64 * H in {256, 768, 1024} and rows = B * L");
  }
  const int threads = 32 * W * rpb;
  const int grid = ln_grid(a->rows, rpb);
  const size_t sm = sizeof(float) * rpb * W * kLnMaxRB * 2;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (p.mode != 2) {   // narrow rows: warp-per-row kernel
    bool done;
    if (p.mode == 0 && p.dy_f32) done = dispatch_ln_bwd_warp<0, true>(p, st);
    else if (p.mode == 0) done = dispatch_ln_bwd_warp<0, false>(p, st);
    else done = dispatch_ln_bwd_warp<1, false>(p, st);
    if (done) {
      SIMVGB_CUDA(cudaGetLastError());
      return 0;
    }
  }
  if (p.mode == 2 && !p.dy_f32 && (p.C == 3072 || p.C == 4096)) {
    long long blocks = p.R;
    const long long cap = (long long)sm_count() * 2;   // 255 registers x 128 threads: two resident blocks per SM
    if (blocks > cap) blocks = cap;
    const size_t ring_bytes = (size_t)kWideStages * 2 * p.C * 2 + kWideStages * 8;
    if (ensure_dynamic_smem(reinterpret_cast<const void*>(ln_bwd_wide_kernel<3>), 3 * 2 * 3072 * 2 + 64)) return -2;
    if (ensure_dynamic_smem(reinterpret_cast<const void*>(ln_bwd_wide_kernel<4>), 3 * 2 * 4096 * 2 + 64)) return -2;
    if (p.C == 3072) ln_bwd_wide_kernel<3><<<(unsigned)blocks, 128, ring_bytes, st>>>(p);
    else ln_bwd_wide_kernel<4><<<(unsigned)blocks, 128, ring_bytes, st>>>(p);
    SIMVGB_CUDA(cudaGetLastError());
    return 0;
  }
  // <= 384 threads per block (C <= 3072): cap registers so two blocks are resident per SM (the kernel is latency-bound)
#define SIMVGB_LN_BWD(M, F)                                                        \
  do {                                                                             \
    if (threads <= 384) ln_bwd_kernel<M, F, 384, 2><<<grid, threads, sm, st>>>(p); \
    else if (threads <= 512) ln_bwd_kernel<M, F, 512, 1><<<grid, threads, sm, st>>>(p); \
    else ln_bwd_kernel<M, F, 1024, 1><<<grid, threads, sm, st>>>(p);               \
  } while (0)
  if (p.mode == 0 && p.dy_f32) SIMVGB_LN_BWD(0, true);
  else if (p.mode == 0) SIMVGB_LN_BWD(0, false);
  else if (p.mode == 1) SIMVGB_LN_BWD(1, false);
  else SIMVGB_LN_BWD(2, false);
  SIMVGB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int simvgb_colsum(const void* in, int in_is_bf16, float* out, void* out_bf16, const float* row_scale,
                             int rows_per_scale, int64_t rows, int C, int64_t ld, void* stream) {
  SIMVGB_CHECK(in && (out || out_bf16), "simvgb_colsum: null pointer");
  SIMVGB_CHECK(C % 8 == 0 && ld % 8 == 0, "simvgb_colsum: C and ld must be multiples of 8");
  if (rows <= 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((C / 8 + 31) / 32, 1);
  long long gy = (rows + 7) / 8;
  const long long cap = (long long)sm_count() * 6 / grid.x + 1;
  grid.y = (unsigned)(gy < cap ? gy : cap);
  const int rps = rows_per_scale > 0 ? rows_per_scale : 1;
  if (in_is_bf16)
    colsum_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)in, out, (bf16*)out_bf16, row_scale, rps, rows, C, (int)ld);
  else
    colsum_kernel<float><<<grid, 256, 0, s>>>((const float*)in, out, (bf16*)out_bf16, row_scale, rps, rows, C, (int)ld);
  SIMVGB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int simvgb_cast_bf16(const float* in, void* out, int64_t n, void* stream) {
  SIMVGB_CHECK(in && out, "simvgb_cast_bf16: null pointer");
  SIMVGB_CHECK(n % 8 == 0, "simvgb_cast_bf16: n must be a multiple of 8 (got %lld)", (long long)n);
  if (n == 0) return 0;
  const long long n8 = n / 8;
  cast_f32_bf16_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, (bf16*)out, n8);
  SIMVGB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int simvgb_embed_vision(const float* patch, const float* cls, const float* posA, float* xv, int B, int N, int D,
                                   void* stream) {
  SIMVGB_CHECK(patch && cls && posA && xv, "simvgb_embed_vision: null pointer");
  SIMVGB_CHECK(D % 4 == 0, "simvgb_embed_vision: D must be a multiple of 4");
  const long long total = (long long)B * (N + 1) * (D / 4);
  assemble_vision_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(patch, cls, posA, xv, B, N, D);
  SIMVGB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int simvgb_embed_text(const float* table, const int64_t* ids, const void* pad, const float* posB, float* xt,
                                 int B, int Lt, int D, void* stream) {
  SIMVGB_CHECK(table && ids && posB && xt, "simvgb_embed_text: null pointer");
  SIMVGB_CHECK(D % 4 == 0, "simvgb_embed_text: D must be a multiple of 4");
  const long long total = (long long)B * Lt * (D / 4);
  if (total == 0) return 0;
  assemble_text_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      table, reinterpret_cast<const long long*>(ids), reinterpret_cast<const unsigned char*>(pad), posB, xt, B, Lt, D);
  SIMVGB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int simvgb_im2col_patch_u8(const void* img_u8, void* cols, int B, int S, int P, const float* mean, const float* std,
                                      int to_rgb, void* stream) {
  SIMVGB_CHECK(img_u8 && cols && mean && std, "simvgb_im2col_patch_u8: null pointer");
  SIMVGB_CHECK(P % 8 == 0 && S % P == 0, "simvgb_im2col_patch_u8: need P %% 8 == 0 and S %% P == 0 (S=%d P=%d)", S, P);
  SIMVGB_CHECK((reinterpret_cast<uintptr_t>(img_u8) & 7) == 0, "simvgb_im2col_patch_u8: image must be 8-byte aligned");
  SIMVGB_CHECK(std[0] != 0.f && std[1] != 0.f && std[2] != 0.f, "simvgb_im2col_patch_u8: zero std");
  const long long total = (long long)B * (S / P) * (S / P) * (P * P / 8);
  // 1 / std evaluated in double and rounded once, as mmcv.imnormalize does (stdinv = 1 / np.float64(std))
  im2col_patch_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint8_t*>(img_u8), (bf16*)cols, B, S, P, mean[0], mean[1], mean[2], (float)(1.0 / (double)std[0]),
      (float)(1.0 / (double)std[1]), (float)(1.0 / (double)std[2]), to_rgb);
  SIMVGB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int simvgb_im2col_patch(const float* img, void* cols, int B, int S, int P, void* stream) {
  SIMVGB_CHECK(img && cols, "simvgb_im2col_patch: null pointer");
  SIMVGB_CHECK(P % 8 == 0 && S % P == 0, "simvgb_im2col_patch: need P %% 8 == 0 and S %% P == 0 (S=%d P=%d)", S, P);
  const long long total = (long long)B * (S / P) * (S / P) * 3 * P * P / 8;
  im2col_patch_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(img, (bf16*)cols, B, S, P);
  SIMVGB_CUDA(cudaGetLastError());
  return 0;
}
