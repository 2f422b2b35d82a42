// Fused multiway self-attention forward for sm_100a:  O = softmax(Q K^T + key_padding_mask) V  per (sample, head).
//
// Replaces torchscale MultiheadAttention's bmm -> masked_fill(-inf) -> softmax(fp32) -> bmm sequence
// (called at /root/reference/simvg/models/vis_encs/beit/beit3_base.py:137-145; semantics SURVEY Appendix A.4) without
// ever materialising the [B*H, L, L] score tensor.  q arrives pre-scaled by head_dim^-0.5 (fused into the QKV GEMM
// epilogue, matching `q *= self.scaling` after the bias add).
//
// Round-2 design (round 1 ran one 128-query tile per CTA, two CTAs per SM, and sat at 0.3 of the tensor peak because every
// CTA serialised S -> softmax -> P -> PV and the two halves of a row exchanged their maxima through shared memory):
//   * persistent, one CTA per SM, each work item = TWO 128-query tiles (A, B) of one (sample, head) that share every K/V
//     tile: the tensor pipe works on one tile while the MUFU/FMA pipes run the other tile's softmax (ping-pong), and K/V
//     cross L2 -> smem once per 256 queries instead of once per 128;
//   * sixteen softmax warps per CTA (eight per query tile: two per TMEM lane quarter, each thread owns one query row x 64 key
//     columns, the two halves of a row exchange their maxima through 4 KB of shared memory): every SM sub-partition has four
//     softmax warps in different phases, so the MUFU unit (the bound at head_dim 64: 16 exp2 / clk / SM against 8192 tensor
//     MACs / clk) stays fed while other warps sit in TMEM loads, row-max chains or barrier waits;
//   * S_X(j+1) is issued as soon as the softmax has pulled S_X(j) out of TMEM (P has its own TMEM columns), so scores are
//     always waiting when a softmax warpgroup comes back; the next work item's Q / K / V loads and first S products overlap
//     this item's epilogue;
//   * P goes to TMEM as packed bf16 pairs (tcgen05.st) and is the A operand of O += P V (TS-form MMA); V is consumed
//     MN-major straight from its token-major TMA tile; online softmax with lazy rescaling (O is touched only when a row
//     maximum grows by more than 2^8).
// Warp roles: 0 = TMA producer, 1 = tcgen05.mma issuer, 2-9 = softmax of tile A, 10-17 = softmax of tile B.  TMEM (512 columns): S_A [0,128)  S_B [128,256)  O_A [256,320)  O_B [320,384)  P_A [384,448)  P_B [448,512).
#include <stdlib.h>

#include "attn_common.cuh"
#include "simvg_b200.h"

namespace simvgb {

#ifndef SIMVGB_FWD_POLY
#define SIMVGB_FWD_POLY 0   // of every 8 exponentials, this many are evaluated on the FMA pipe instead of MUFU (full tiles only)
#endif
#ifndef SIMVGB_FWD_KS
#define SIMVGB_FWD_KS 3
#endif
#ifndef SIMVGB_FWD_EXP16
#define SIMVGB_FWD_EXP16 0    // 1: exponentials as ex2.approx.f16x2 (experiment: ptxas splits it into two MUFU.EX2.F16, see exp_chunk16)
#endif
#ifndef SIMVGB_FWD_STAGGER
#define SIMVGB_FWD_STAGGER 0   // 1: hold tile B's first S product back until tile A is half-way through key tile 0 (anti-phase)
#endif
constexpr int kFwdThreads = 576;
constexpr int kKS = SIMVGB_FWD_KS;   // K ring depth = V ring depth
constexpr int kFwdSmem = (2 + 2 * kKS) * kTileBytes + 1024 /*align*/ + 512 /*barriers*/ + 4096 /*row-max exchange*/;
constexpr float kLog2e = 1.4426950408889634f;

struct AttnFwdParams {
  AttnGeom g;
  const unsigned char* pad;  // [B, Lt] 1 = padded text token, or null
  bf16* out_v;               // [B*Lv, D]
  bf16* out_t;               // [B*Lt, D]
  float* lse;                // [B, H, ntiles*128]  log2-domain logsumexp of each query row
  int npairs;                // ntiles / 2  (work items with two query tiles)
  int n_full_items;          // B * H * npairs
  int n_items;               // + B * H single-tile items when ntiles is odd
  long long* ts;             // SIMVGB_FWD_TRACE builds only: clock64 trace of CTA 0's first work item (tools/attn_fwd_trace.py)
};

#ifdef SIMVGB_FWD_TRACE
static long long* g_fwd_trace = nullptr;
#define FWD_TS(cond, slot) do { if (cond) p.ts[(slot)] = clock64(); } while (0)
#else
#define FWD_TS(cond, slot) do { } while (0)
#endif

struct FwdItem {
  int b, h, tA, tB;   // tB < 0: only tile A
};

__device__ __forceinline__ FwdItem decode_item(const AttnFwdParams& p, int it) {
  FwdItem w;
  if (it < p.n_full_items) {
    const int bh = it / p.npairs, qp = it - bh * p.npairs;
    w.b = bh / p.g.H; w.h = bh - w.b * p.g.H; w.tA = 2 * qp; w.tB = 2 * qp + 1;
  } else {
    const int bh = it - p.n_full_items;
    w.b = bh / p.g.H; w.h = bh - w.b * p.g.H; w.tA = p.g.ntiles - 1; w.tB = -1;
  }
  return w;
}

// 32 lanes x 8 consecutive fp32 columns (the rare O-rescale path works in small pieces so it needs few registers while the
// 64 scores of the tile are live)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void pair_sync(int id) {   // the two softmax warps that share a (tile, TMEM lane quarter)
  asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

// exp2 of 32 scores (already in registers) -> 16 packed bf16 pairs, returns their fp32 sum.  x = s * log2e - m.
template <bool kPoly>
__device__ __forceinline__ float exp_chunk(const uint32_t (&s)[32], uint32_t (&pk)[16], float neg_m) {
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float x0 = fmaf(__uint_as_float(s[2 * i]), kLog2e, neg_m);
    const float x1 = fmaf(__uint_as_float(s[2 * i + 1]), kLog2e, neg_m);
    float e0, e1;
    // exponent i of this chunk goes to the FMA pipe when (2i mod 8) < SIMVGB_FWD_POLY (resp. 2i+1)
    if (kPoly && ((2 * i) & 7) < SIMVGB_FWD_POLY) e0 = ex2_fma(fmaxf(x0, -100.f)); else e0 = ex2_approx(x0);
    if (kPoly && ((2 * i + 1) & 7) < SIMVGB_FWD_POLY) e1 = ex2_fma(fmaxf(x1, -100.f)); else e1 = ex2_approx(x1);
    sum0 += e0;
    sum1 += e1;
    pk[i] = pack_bf16x2(e0, e1);   // TMEM column c of P holds keys 2c (low half), 2c+1 (high half)
  }
  return sum0 + sum1;
}

// The same with half-precision exponentials: MUFU evaluates ex2.approx.f16x2 at the instruction rate of the fp32 form, i.e. two
// exponentials per operation, and MUFU (16 / clk / SM) is what the softmax phases saturate at head_dim 64 (measured: 8.2 clk
// per MUFU instruction and SM sub-partition during the exp phases).  Accuracy: the argument x = s*log2e - m (<= 8 by the lazy
// rescaling) is rounded to fp16 — |x| < 16 keeps 2^x within 0.27 %, comparable to the bf16 rounding P gets anyway (0.2 %), and
// such x carry weights < 2^-8 of the row maximum; fp16 results keep denormals (no flush: 2^-24 absolute resolution against a
// row maximum >= 2^-8... 2^8), row sums are accumulated in fp32 from the very values that are packed into P.
__device__ __forceinline__ float exp_chunk16(const uint32_t (&s)[32], uint32_t (&pk)[16], float neg_m) {
  float2 acc = make_float2(0.f, 0.f);
  const float2 sc = make_float2(kLog2e, kLog2e), nm = make_float2(neg_m, neg_m);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 x = fma2(make_float2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), sc, nm);
    uint32_t h2, e2;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(x.y), "f"(x.x));       // {hi, lo} = {x.y, x.x}
    asm("ex2.approx.f16x2 %0, %1;" : "=r"(e2) : "r"(h2));
    float lo, hi;
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(lo), "=f"(hi) : "r"(e2));
    acc = add2(acc, make_float2(lo, hi));
    pk[i] = pack_bf16x2(lo, hi);   // TMEM column c of P holds keys 2c (low half), 2c+1 (high half)
  }
  return acc.x + acc.y;
}

__global__ void __launch_bounds__(kFwdThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap map_full, const __grid_constant__ CUtensorMap map_tail,
                const __grid_constant__ CUtensorMap map_text, const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                              // [2]
  uint8_t* sK = smem + 2 * kTileBytes;             // [kKS]
  uint8_t* sV = smem + (2 + kKS) * kTileBytes;     // [kKS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (2 + 2 * kKS) * kTileBytes);
  uint64_t* q_full = bars;                 // [2]
  uint64_t* q_empty = bars + 2;            // [2]
  uint64_t* k_full = bars + 4;             // [kKS]
  uint64_t* k_empty = k_full + kKS;        // [kKS]
  uint64_t* v_full = k_empty + kKS;        // [kKS]
  uint64_t* v_empty = v_full + kKS;        // [kKS]
  uint64_t* s_full = v_empty + kKS;        // [2]  S_X(j) complete
  uint64_t* s_free = s_full + 2;           // [2]  softmax X holds S_X(j) in registers
  uint64_t* p_full = s_full + 4;           // [2]  P_X(j) written (and O_X rescaled if it had to be)
  uint64_t* pv_done = s_full + 6;          // [2]  O_X += P_X(j) V(j) complete
  uint64_t* o_free = s_full + 8;           // [2]  epilogue X holds O_X in registers
  uint64_t* mid = s_full + 10;             // tile A is half-way through the first key tile of a work item (staggers tile B)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 11);
  float* xchg = reinterpret_cast<float*>(smem + (2 + 2 * kKS) * kTileBytes + 512);   // [tile 2][parity 2][half 2][128 rows]

  const AttnGeom& g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = g.ntiles;

  // Zero the staging tiles: partial (tail/text) tiles only overwrite some rows, the rest must stay finite.
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < (2 + 2 * kKS) * kTileBytes / 16; i += kFwdThreads) z[i] = zero;
  }
  if (threadIdx.x == 0) {
    for (int x = 0; x < 2; ++x) {
      mbar_init(&q_full[x], 1); mbar_init(&q_empty[x], 1);
      mbar_init(&s_full[x], 1); mbar_init(&s_free[x], 8);     // one elected arrival per softmax warp
      mbar_init(&p_full[x], 8); mbar_init(&pv_done[x], 1); mbar_init(&o_free[x], 8);
    }
    mbar_init(mid, 8);
    for (int s = 0; s < kKS; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async();  // generic-proxy zero fill -> visible before TMA (async proxy) writes
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // provably warp-uniform (uniform registers for UTCHMMA)

  if (warp < 2) {
    if (warp == 0) {
      // ------------------------------ TMA producer ------------------------------
      if (lane == 0) {
        uint32_t kc = 0, vc = 0, qc[2] = {0, 0};
        for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
          const FwdItem w = decode_item(p, it);
          const int colq = w.h * kHeadDim, colk = g.D + colq, colv = 2 * g.D + colq;
          auto load_q = [&](int x, int t) {
            mbar_wait(&q_empty[x], (qc[x] & 1) ^ 1);          // freed when the previous item's last S_X retired
            load_virtual_tile(sQ + x * kTileBytes, &q_full[x], g, &map_full, &map_tail, &map_text, t, colq, w.b);
            ++qc[x];
          };
          auto load_k = [&](int j) {
            const uint32_t slot = kc % kKS;
            mbar_wait(&k_empty[slot], ((kc / kKS) & 1) ^ 1);
            load_virtual_tile(sK + slot * kTileBytes, &k_full[slot], g, &map_full, &map_tail, &map_text, j, colk, w.b);
            ++kc;
          };
          auto load_v = [&](int j) {
            const uint32_t slot = vc % kKS;
            mbar_wait(&v_empty[slot], ((vc / kKS) & 1) ^ 1);
            load_virtual_tile(sV + slot * kTileBytes, &v_full[slot], g, &map_full, &map_tail, &map_text, j, colv, w.b);
            ++vc;
          };
          load_q(0, w.tA);
          load_k(0);
          if (w.tB >= 0) load_q(1, w.tB);
          load_v(0);
          for (int j = 1; j < nk; ++j) { load_k(j); load_v(j); }
        }
      }
    } else if (warp == 1) {
      // ------------------------------ MMA issuer: warp-uniform control flow, single-lane issue ------------------------------
      const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(128, kHeadDim, 0, 1);  // A = P (TMEM), B = V (MN-major)
      const uint64_t dQ0 = umma_smem_desc(smem_u32(sQ), 16, 1024);
      const uint64_t dK0 = umma_smem_desc(smem_u32(sK), 16, 1024);
      const uint64_t dV0 = umma_smem_desc(smem_u32(sV), 8192, 1024);
      constexpr uint32_t kStep = kTileBytes >> 4;
      uint32_t kc = 0, vc = 0, mc = 0, sc[2] = {0, 0}, pc[2] = {0, 0}, ic[2] = {0, 0};
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
        const FwdItem w = decode_item(p, it);
        const bool both = w.tB >= 0;
        // S_X = Q_X K_j^T into TMEM columns [128 X, 128 X + 128); `last` also releases the Q_X buffer
        auto issue_s = [&](int x, uint32_t kslot, bool last) {
          mbar_wait(&s_free[x], (sc[x] & 1) ^ 1);             // softmax X has pulled the previous S_X out of TMEM
          tc_fence_after();
          const uint64_t dq = dQ0 + x * kStep, dk = dK0 + kslot * kStep;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kHeadDim / 16; ++k) umma_f16_ss(tmem + 128 * x, dq + 2 * k, dk + 2 * k, idesc_s, k > 0);
            umma_commit(&s_full[x]);
            if (last) umma_commit(&q_empty[x]);
          }
          __syncwarp();
          ++sc[x];
        };
        // O_X (+)= P_X V_j
        auto issue_pv = [&](int x, uint32_t vslot, bool first) {
          mbar_wait(&p_full[x], pc[x] & 1);
          if (first) mbar_wait(&o_free[x], (ic[x] & 1) ^ 1);  // the previous item's epilogue has read O_X
          tc_fence_after();
          const uint64_t dv = dV0 + vslot * kStep;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kTile / 16; ++k)   // A = P[:, 16k .. 16k+16) = 8 packed TMEM columns
              umma_f16_ts(tmem + 256 + 64 * x, tmem + 384 + 64 * x + 8 * k, dv + k * 128, idesc_o, (!first || k > 0) ? 1u : 0u);
            umma_commit(&pv_done[x]);
          }
          __syncwarp();
          ++pc[x];
        };
        // SIMVGB_FWD_STAGGER: tile B's first S product is held back until A's softmax is half-way through key tile 0, which keeps
        // the two tiles in anti-phase for the whole item.  Measured (tools/attn_fwd_trace.py): the anti-phase holds (B lags A by
        // ~1600 of 3130 clk) and the per-tile period drops 3350 -> 3130 clk, but every item then ends with half a period of tile B
        // alone: 0.960 ms against 0.841 ms without — off by default.
        uint32_t kslot0 = kc % kKS;
        mbar_wait(&q_full[0], ic[0] & 1);
        mbar_wait(&k_full[kslot0], (kc / kKS) & 1);
        issue_s(0, kslot0, nk == 1);
        if (!SIMVGB_FWD_STAGGER && both) {
          mbar_wait(&q_full[1], ic[1] & 1);
          issue_s(1, kslot0, nk == 1);
        }
        ++kc;
        for (int j = 0; j < nk; ++j) {
          const bool more = j + 1 < nk;
          const uint32_t vslot = vc % kKS;
          const uint32_t kslot = kc % kKS;
          if (more) {
            mbar_wait(&k_full[kslot], (kc / kKS) & 1);
            issue_s(0, kslot, j + 2 == nk);
          }
          if (j == 0) {
            if (SIMVGB_FWD_STAGGER) {
              mbar_wait(mid, mc & 1);
              ++mc;
            }
            if (SIMVGB_FWD_STAGGER && both) {
              mbar_wait(&q_full[1], ic[1] & 1);
              issue_s(1, kslot0, nk == 1);
            }
          }
          mbar_wait(&v_full[vslot], (vc / kKS) & 1);
          issue_pv(0, vslot, j == 0);
          if (both) {
            if (more) issue_s(1, kslot, j + 2 == nk);
            issue_pv(1, vslot, j == 0);
          }
          if (elect_one()) {
            if (j == 0) umma_commit(&k_empty[kslot0]);
            if (more) umma_commit(&k_empty[kslot]);
            umma_commit(&v_empty[vslot]);
          }
          __syncwarp();
          if (more) ++kc;
          ++vc;
        }
        ++ic[0];
        if (both) ++ic[1];
      }
    }
  } else {
// ------------------------------ softmax + epilogue: thread = (query row of tile X, 64-key half) ------------------------------
    const int k16 = warp - 2;
    const int x = k16 >> 3;
    const int half = (k16 >> 2) & 1;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int bar_id = 1 + x * 4 + quarter;
    const uint32_t lane_base = uint32_t(quarter * 32) << 16;
    const uint32_t tmS = tmem + 128 * x + lane_base + 64 * half;
    const uint32_t tmO = tmem + 256 + 64 * x + lane_base + 32 * half;
    const uint32_t tmP = tmem + 384 + 64 * x + lane_base + 32 * half;
    float* xrow = xchg + x * 512;
    uint32_t sc = 0, pc = 0;
    for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
      const FwdItem w = decode_item(p, it);
      const int qt = x == 0 ? w.tA : w.tB;
      if (qt < 0) continue;
      // validity bits of this half of the (at most two) partial key tiles of sample b: bit c of word [t][k] = key
      // t*128 + 64*half + 32k + c is a real, unpadded token.  One ballot per word (warp-uniform result).
      uint32_t mk[2][2];
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int v = (g.nfull + t) * kTile + 64 * half + 32 * k + lane;
          bool ok = v < g.Lv;
          if (!ok && v >= g.T0 && v < g.T0 + g.Lt) ok = (p.pad == nullptr) || (p.pad[w.b * g.Lt + (v - g.T0)] == 0);
          mk[t][k] = __ballot_sync(0xffffffffu, ok);
        }
      float m = -INFINITY, l = 0.f;   // l: partial row sum over this thread's columns
#ifdef SIMVGB_FWD_TRACE
      const bool tr = p.ts != nullptr && blockIdx.x == 0 && it == 0 && lane == 0 && quarter == 0 && half == 0;
      const int tb = x * 256;
#endif
      for (int j = 0; j < nk; ++j) {
        FWD_TS(tr, tb + j * 8 + 0);
        mbar_wait(&s_full[x], sc & 1);
        FWD_TS(tr, tb + j * 8 + 1);
        tc_fence_after();
        uint32_t s0[32], s1[32];
        tmem_ld32(tmS, s0);
        tmem_ld32(tmS + 32, s1);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[x]);   // scores are in registers: the MMA warp may overwrite S_X
        FWD_TS(tr, tb + j * 8 + 2);
        if (j >= g.nfull) {
          const int t = j - g.nfull;
          const uint32_t ba = mk[t][0], bb = mk[t][1];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (!((ba >> i) & 1u)) s0[i] = 0xff800000u;   // -inf
            if (!((bb >> i) & 1u)) s1[i] = 0xff800000u;
          }
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(s0[i]), __uint_as_float(s0[i + 16])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(s1[i]), __uint_as_float(s1[i + 16])));
        }
        mx2 = fmaxf(mx0, mx1);
        float* xm = xrow + (sc & 1) * 256;
        xm[half * 128 + r] = mx2;
        pair_sync(bar_id);
        mx3 = fmaxf(mx2, xm[(half ^ 1) * 128 + r]);
        ++sc;
        FWD_TS(tr, tb + j * 8 + 3);
        const float mx = mx3 * kLog2e;
        const bool need = mx > m + 8.0f;       // lazy rescale threshold (log2 units); true on the first tile
        const float m_use = need ? mx : m;
        const float alpha = need ? ex2_approx(m - m_use) : 1.0f;
        const float neg_m = -m_use;
        uint32_t pk[16];
        // the first 32 exponentials need neither the P buffer nor O: they run while P_X(j-1) V_(j-1) is still on the tensor pipe
#if SIMVGB_FWD_EXP16
        float sum = exp_chunk16(s0, pk, neg_m);
#else
        float sum = exp_chunk<(SIMVGB_FWD_POLY > 0)>(s0, pk, neg_m);
#endif
        if (SIMVGB_FWD_STAGGER && x == 0 && j == 0) {
          __syncwarp();
          if (lane == 0) mbar_arrive(mid);
        }
        FWD_TS(tr, tb + j * 8 + 4);
        mbar_wait(&pv_done[x], (pc & 1) ^ 1);  // P_X buffer free, O_X complete up to tile j-1
        FWD_TS(tr, tb + j * 8 + 5);
        tc_fence_after();
        if (j > 0 && __any_sync(0xffffffffu, need)) {
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t v[8];
            tmem_ld8(tmO + 8 * c, v);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st8(tmO + 8 * c, v);
          }
        }
        tmem_st16(tmP, pk);
#if SIMVGB_FWD_EXP16
        sum += exp_chunk16(s1, pk, neg_m);
#else
        sum += exp_chunk<(SIMVGB_FWD_POLY > 0)>(s1, pk, neg_m);
#endif
        tmem_st16(tmP + 16, pk);
        tmem_wait_st();
        l = l * alpha + sum;
        m = m_use;
        tc_fence_before();     // P (and a rescaled O) are in TMEM: order them before the MMA warp's tcgen05.mma
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[x]);
        ++pc;
        FWD_TS(tr, tb + j * 8 + 6);
      }
      // ------------------------------ epilogue ------------------------------
      float* xs = xrow + (sc & 1) * 256;      // the buffer tile nk would use: free (tile nk-2's readers passed tile nk-1's barrier)
      xs[half * 128 + r] = l;
      pair_sync(bar_id);
      l += xs[(half ^ 1) * 128 + r];
      mbar_wait(&pv_done[x], (pc - 1) & 1);
      tc_fence_after();
      uint32_t o0[32];
      tmem_ld32(tmO, o0);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[x]);    // the next item's first P V may overwrite O_X
      const int qv = qt * kTile + r;
      bf16* dst = nullptr;
      if (qv < g.Lv) dst = p.out_v + ((long long)w.b * g.Lv + qv) * g.D + w.h * kHeadDim;
      else if (qv >= g.T0 && qv < g.T0 + g.Lt) dst = p.out_t + ((long long)w.b * g.Lt + (qv - g.T0)) * g.D + w.h * kHeadDim;
      const float inv = 1.0f / l;
      if (dst != nullptr) {
        uint4* o = reinterpret_cast<uint4*>(dst + 32 * half);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          o[q4] = make_uint4(pack_bf16x2(__uint_as_float(o0[8 * q4]) * inv, __uint_as_float(o0[8 * q4 + 1]) * inv),
                             pack_bf16x2(__uint_as_float(o0[8 * q4 + 2]) * inv, __uint_as_float(o0[8 * q4 + 3]) * inv),
                             pack_bf16x2(__uint_as_float(o0[8 * q4 + 4]) * inv, __uint_as_float(o0[8 * q4 + 5]) * inv),
                             pack_bf16x2(__uint_as_float(o0[8 * q4 + 6]) * inv, __uint_as_float(o0[8 * q4 + 7]) * inv));
      }
      // positions of the virtual axis that are not tokens get LSE = +inf: the backward then computes P = exp2(S - inf) = 0
      // for them without any query-side masking
      if (p.lse != nullptr && half == 0)
        p.lse[((long long)w.b * g.H + w.h) * (g.ntiles * kTile) + qv] = dst != nullptr ? m + log2f(l) : INFINITY;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// Tensor maps over the token-major qkv buffers: dims (3D, tokens-per-sample, B).
int make_attn_maps(CUtensorMap* full, CUtensorMap* tail, CUtensorMap* text, const AttnGeom& g, const void* base_v,
                   const void* base_t, int row_elems) {
  uint64_t dims[3] = {(uint64_t)row_elems, (uint64_t)g.Lv, (uint64_t)g.B};
  uint64_t strides[2] = {(uint64_t)row_elems * 2, (uint64_t)row_elems * 2 * g.Lv};
  uint32_t box[3] = {kHeadDim, kTile, 1};
  if (make_tmap(full, base_v, 2, 3, dims, strides, box, 1)) return -1;
  box[1] = g.tail_rows > 0 ? g.tail_rows : 8;
  if (make_tmap(tail, base_v, 2, 3, dims, strides, box, 1)) return -1;
  if (g.Lt > 0) {
    dims[1] = g.Lt;
    strides[1] = (uint64_t)row_elems * 2 * g.Lt;
    box[1] = g.Ltp;
    if (make_tmap(text, base_t, 2, 3, dims, strides, box, 1)) return -1;
  } else {
    *text = *tail;
  }
  return 0;
}

}  // namespace simvgb

extern "C" int simvgb_attn_fwd(const simvgb_attn_args* a, void* stream) {
  using namespace simvgb;
  SIMVGB_CHECK(a != nullptr, "simvgb_attn_fwd: null args");
  SIMVGB_CHECK(a->B > 0 && a->H > 0 && a->Lv > 0 && a->Lt >= 0, "simvgb_attn_fwd: bad shape");
  SIMVGB_CHECK(a->head_dim == kHeadDim, "simvgb_attn_fwd: head_dim must be 64 (got %d)", a->head_dim);
  SIMVGB_CHECK(a->Lt <= 120, "simvgb_attn_fwd: at most 120 text tokens (got %d)", a->Lt);
  SIMVGB_CHECK(a->qkv_v && a->out_v && (a->Lt == 0 || (a->qkv_t && a->out_t)), "simvgb_attn_fwd: null buffer");
  const int D = a->H * kHeadDim;
  AttnFwdParams p;
  p.g = make_attn_geom(a->B, a->H, a->Lv, a->Lt, D);
  p.pad = reinterpret_cast<const unsigned char*>(a->text_pad);
  p.out_v = reinterpret_cast<bf16*>(a->out_v);
  p.out_t = reinterpret_cast<bf16*>(a->out_t);
  p.lse = a->lse;
  p.npairs = p.g.ntiles / 2;
  p.n_full_items = a->B * a->H * p.npairs;
  p.n_items = p.n_full_items + ((p.g.ntiles & 1) ? a->B * a->H : 0);
#ifdef SIMVGB_FWD_TRACE
  p.ts = g_fwd_trace;
#else
  p.ts = nullptr;
#endif
  CUtensorMap full, tail, text;
  if (make_attn_maps(&full, &tail, &text, p.g, a->qkv_v, a->qkv_t, 3 * D)) return -1;
  if (ensure_dynamic_smem(reinterpret_cast<const void*>(attn_fwd_kernel), kFwdSmem)) return -2;
  int grid = sm_count();
  if (grid > p.n_items) grid = p.n_items;
  attn_fwd_kernel<<<grid, kFwdThreads, kFwdSmem, reinterpret_cast<cudaStream_t>(stream)>>>(full, tail, text, p);
  SIMVGB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int simvgb_attn_lse_stride(int Lv, int Lt) {
  simvgb::AttnGeom g = simvgb::make_attn_geom(1, 1, Lv, Lt, 64);
  return g.ntiles * simvgb::kTile;
}

extern "C" int simvgb_attn_text_offset(int Lv, int Lt) { return simvgb::make_attn_geom(1, 1, Lv, Lt, 64).T0; }

#ifdef SIMVGB_FWD_TRACE
extern "C" void simvgb_debug_attn_fwd_trace(long long* buf) { simvgb::g_fwd_trace = buf; }
#endif
