// Fused multiway self-attention forward for sm_100a:  O = softmax(Q K^T + key_padding_mask) V  per (sample, head).
//
// Replaces torchscale MultiheadAttention's bmm -> masked_fill(-inf) -> softmax(fp32) -> bmm sequence
// (called at /root/reference/simvg/models/vis_encs/beit/beit3_base.py:137-145; semantics SURVEY Appendix A.4) without
// ever materialising the [B*H, L, L] score tensor.  q arrives pre-scaled by head_dim^-0.5 (fused into the QKV GEMM
// epilogue, matching `q *= self.scaling` after the bias add).
//
// One CTA = one 128-query tile of one (b, h); two CTAs are co-resident per SM so one CTA's softmax overlaps the other's
// MMAs.  Warp roles:  0 = TMA producer, 1 = tcgen05.mma issuer, 2-9 = softmax: two warps per TMEM lane quarter, each
// thread owns one query row x 64 key columns (row max exchanged through smem) so every SM sub-partition always has
// several softmax warps to switch between while MUFU / TMEM loads are in flight.
//   S = Q K_j^T   : tcgen05.mma M=128 N=128 K=64, both operands K-major, S in TMEM columns [0,128)
//   O += P V_j    : P (bf16 pairs packed in 32-bit TMEM columns [192,256)) is written by the softmax threads with
//                   tcgen05.st and consumed as the A operand straight from TMEM (shared memory only feeds B: at
//                   head_dim 64 the kernel is bound by shared-memory bandwidth, and P through smem cost 64 KB of the
//                   144 KB moved per tile); V_j read MN-major straight from its token-major TMA tile; O in TMEM
//                   columns [128,192)
// Online softmax with lazy rescaling (O is only rescaled when a row max grows by more than 2^8).
#include <stdlib.h>

#include "attn_common.cuh"

#ifdef SIMVGB_ATTN_ABLATE   // timing ablations (tools/attn_ablate.py): build with -DSIMVGB_ATTN_ABLATE
#define SIMVGB_DBG(p) ((p).dbg)
#else
#define SIMVGB_DBG(p) 0
#endif
#include "simvg_b200.h"

namespace simvgb {

static long long* g_fwd_trace = nullptr;

#ifndef SIMVGB_FWD_POLY
#define SIMVGB_FWD_POLY 0   // exponentials per group of 4 evaluated on the FMA pipe instead of MUFU (0, 1 or 2)
#endif
constexpr int kFwdThreads = 320;
constexpr int kSoftmaxThreads = 256;
constexpr int kSlots = 4;  // staging tiles: K double-buffered (slots 0,1), V double-buffered (slots 2,3); both prefetched a tile ahead
constexpr int kFwdSmem = kTileBytes /*Q*/ + kSlots * kTileBytes + 1024 /*align*/ + 256 /*barriers*/ + 2048 /*row-max exchange*/;
constexpr float kLog2e = 1.4426950408889634f;

struct AttnFwdParams {
  AttnGeom g;
  const unsigned char* pad;  // [B, Lt] 1 = padded text token, or null
  bf16* out_v;               // [B*Lv, D]
  bf16* out_t;               // [B*Lt, D]
  float* lse;                // [B, H, ntiles*128]  log2-domain logsumexp of each query row
  int dbg;                   // timing ablations (SIMVGB_ATTN_DEBUG, tools/attn_ablate.py); 0 in production
  long long* ts;             // optional clock64 trace of CTA (0,0,0) (tools/attn_trace.py)
};

__device__ __forceinline__ void pair_sync(int quarter) {   // the two softmax warps that share a TMEM lane quarter
  asm volatile("bar.sync %0, 64;" ::"r"(quarter + 1) : "memory");
}

__global__ void __launch_bounds__(kFwdThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap map_full, const __grid_constant__ CUtensorMap map_tail,
                const __grid_constant__ CUtensorMap map_text, const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (1 + kSlots) * kTileBytes);
  uint64_t* q_full = bars;
  uint64_t* slot_full = bars + 1;             // [kSlots]  0,1 = K ring, 2,3 = V ring
  uint64_t* slot_empty = bars + 1 + kSlots;   // [kSlots]
  uint64_t* s_full = bars + 1 + 2 * kSlots;
  uint64_t* s_empty = s_full + 1;
  uint64_t* p_full = s_full + 2;
  uint64_t* pv_done = s_full + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 4);
  uint32_t* masks = tmem_slot + 2;            // [2][4] validity bits of the (at most two) partial tiles
  float* xchg = reinterpret_cast<float*>(smem + (1 + kSlots) * kTileBytes + 256);  // [2 parity][2 halves][128 rows]

  const AttnGeom& g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nk = g.ntiles;

  // Zero the staging tiles: partial (tail/text) tiles only overwrite some rows, the rest must stay finite.
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < (1 + kSlots) * kTileBytes / 16; i += kFwdThreads) z[i] = zero;
  }
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < kSlots; ++s) { mbar_init(&slot_full[s], 1); mbar_init(&slot_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(s_empty, kSoftmaxThreads / 32);   // one elected arrival per softmax warp
    mbar_init(p_full, kSoftmaxThreads / 32);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 256) {   // warps 2-5: first partial tile, warps 6-9: second partial tile
    const int which = (threadIdx.x - 64) >> 7;
    const int t = g.nfull + which;
    if (t < nk) build_tile_mask(masks + 4 * which, g, p.pad, b, t, (threadIdx.x - 64) & 127);
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  fence_proxy_async();  // generic-proxy zero fill -> visible before TMA (async proxy) writes
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // provably warp-uniform (uniform registers for UTCHMMA)
  const uint32_t tmS = tmem, tmO = tmem + 128, tmP = tmem + 192;

  if (warp == 0) {
    if (lane == 0) {
      const int colq = h * kHeadDim, colk = g.D + h * kHeadDim, colv = 2 * g.D + h * kHeadDim;
      load_virtual_tile(sQ, q_full, g, &map_full, &map_tail, &map_text, qt, colq, b);
      auto load_k = [&](int j) {
        const int slot = j & 1;
        mbar_wait(&slot_empty[slot], ((j >> 1) & 1) ^ 1);       // freed when S_{j-2} retired
        load_virtual_tile(sKV + slot * kTileBytes, &slot_full[slot], g, &map_full, &map_tail, &map_text, j, colk, b);
      };
      auto load_v = [&](int j) {
        const int slot = 2 + (j & 1);
        mbar_wait(&slot_empty[slot], ((j >> 1) & 1) ^ 1);       // freed when P V_{j-2} retired
        load_virtual_tile(sKV + slot * kTileBytes, &slot_full[slot], g, &map_full, &map_tail, &map_text, j, colv, b);
      };
      load_k(0);
      for (int j = 0; j < nk; ++j) {
        if (j + 1 < nk) load_k(j + 1);
        load_v(j);
      }
    }
  } else if (warp == 1) {
    // MMA issuer: warp-uniform control flow, single-lane issue (keeps descriptors in uniform registers).
    const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
    const uint32_t idesc_o = umma_idesc_bf16(128, kHeadDim, 0, 1);  // A = P (K-major), B = V (MN-major)
    const uint64_t dQ = umma_smem_desc(smem_u32(sQ), 16, 1024);
    const uint64_t dKV_k = umma_smem_desc(smem_u32(sKV), 16, 1024), dKV_mn = umma_smem_desc(smem_u32(sKV), 8192, 1024);
    auto issue_s = [&](int j) {
      const int slot = j & 1;
      mbar_wait(&slot_full[slot], (j >> 1) & 1);
      tc_fence_after();
      const uint64_t dk = dKV_k + slot * (kTileBytes >> 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k) umma_f16_ss(tmS, dQ + 2 * k, dk + 2 * k, idesc_s, k > 0);
        umma_commit(&slot_empty[slot]);
        umma_commit(s_full);
      }
      __syncwarp();
    };
    const bool trace = p.ts != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
    mbar_wait(q_full, 0);
    issue_s(0);
    for (int j = 0; j < nk; ++j) {
      if (trace) p.ts[j * 8 + 0] = clock64();
      mbar_wait(s_empty, j & 1);  // softmax has consumed S_j
      if (trace) p.ts[j * 8 + 1] = clock64();
      if (j + 1 < nk) issue_s(j + 1);
      if (trace) p.ts[j * 8 + 2] = clock64();
      const int slot = 2 + (j & 1);
      mbar_wait(p_full, j & 1);
      if (trace) p.ts[j * 8 + 3] = clock64();
      mbar_wait(&slot_full[slot], (j >> 1) & 1);
      if (trace) p.ts[j * 8 + 4] = clock64();
      tc_fence_after();
      const uint64_t dv = dKV_mn + slot * (kTileBytes >> 4);
      if (elect_one()) {
        if (!(SIMVGB_DBG(p) & 8))
#pragma unroll
        for (int k = 0; k < kTile / 16; ++k)   // A = P[:, 16k .. 16k+16) = 8 packed TMEM columns
          umma_f16_ts(tmO, tmP + 8 * k, dv + k * 128, idesc_o, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(&slot_empty[slot]);
        umma_commit(pv_done);
      }
      __syncwarp();
      if (trace) p.ts[j * 8 + 5] = clock64();
    }
  } else {
    // ------------------------------ softmax: thread = (query row, 64-column half) ------------------------------
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = uint32_t(quarter * 32) << 16;
    const uint32_t col_base = half * 64;
    float m = -INFINITY, l = 0.f;   // l: partial row sum over this thread's columns
    for (int j = 0; j < nk; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const bool partial = j >= g.nfull;
      const uint32_t* mk = masks + 4 * (j - g.nfull) + half * 2;
      uint32_t va[32], vb[32];
      if (!(SIMVGB_DBG(p) & 1)) {
        tmem_ld32(tmS + lane_base + col_base, va);
        tmem_ld32(tmS + lane_base + col_base + 32, vb);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) { va[i] = 0; vb[i] = 0; }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);   // scores are in registers: the MMA warp may overwrite S with Q K_{j+1}^T
      if (partial) {
        const uint32_t ba = mk[0], bb = mk[1];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (!((ba >> i) & 1u)) va[i] = 0xff800000u;   // -inf
          if (!((bb >> i) & 1u)) vb[i] = 0xff800000u;
        }
      }
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(va[i]), __uint_as_float(vb[i])));
      float* xm = xchg + (j & 1) * 256;
      if (!(SIMVGB_DBG(p) & 16)) {
        xm[half * 128 + r] = mx;
        pair_sync(quarter);
        mx = fmaxf(mx, xm[(half ^ 1) * 128 + r]);
      }
      mx *= kLog2e;
      const bool need = mx > m + 8.0f;       // lazy rescale threshold (log2 units); true on the first tile
      const float m_use = need ? mx : m;
      const float alpha = need ? ex2_approx(m - m_use) : 1.0f;
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);     // P buffer free, O up to tile j-1 complete
        if (__any_sync(0xffffffffu, need)) {
          tc_fence_after();
          uint32_t v[32];
          tmem_ld32(tmO + lane_base + half * 32, v);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st32(tmO + lane_base + half * 32, v);
          tmem_wait_st();
        }
      }
      // exp2 and bf16 packing fused (in place: pair i of va/vb lands in word i), so no fp32 probability outlives its pair
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float a0, a1, b0, b1;
        if (SIMVGB_DBG(p) & 2) {
          a0 = __uint_as_float(va[2 * i]); a1 = __uint_as_float(va[2 * i + 1]);
          b0 = __uint_as_float(vb[2 * i]); b1 = __uint_as_float(vb[2 * i + 1]);
        } else {
          a0 = ex2_approx(fmaf(__uint_as_float(va[2 * i]), kLog2e, -m_use));
          a1 = ex2_approx(fmaf(__uint_as_float(va[2 * i + 1]), kLog2e, -m_use));
#if SIMVGB_FWD_POLY >= 2
          const float xb0 = fmaf(__uint_as_float(vb[2 * i]), kLog2e, -m_use);
          b0 = partial ? ex2_approx(xb0) : ex2_fma(fmaxf(xb0, -100.f));
#else
          b0 = ex2_approx(fmaf(__uint_as_float(vb[2 * i]), kLog2e, -m_use));
#endif
#if SIMVGB_FWD_POLY >= 1
          // full tiles only (warp-uniform): masked scores are -inf, which the FMA-pipe exp2 does not handle
          const float xb1 = fmaf(__uint_as_float(vb[2 * i + 1]), kLog2e, -m_use);
          b1 = partial ? ex2_approx(xb1) : ex2_fma(fmaxf(xb1, -100.f));
#else
          b1 = ex2_approx(fmaf(__uint_as_float(vb[2 * i + 1]), kLog2e, -m_use));
#endif
        }
        sum += (a0 + a1) + (b0 + b1);
        va[i] = pack_bf16x2(a0, a1);   // TMEM column c of P holds keys 2c (low half), 2c+1 (high half)
        vb[i] = pack_bf16x2(b0, b1);
      }
      if (!(SIMVGB_DBG(p) & 4)) {
        tmem_st16x2(tmP + lane_base + half * 32, va, vb);
        tmem_wait_st();
      }
      l = l * alpha + sum;
      m = m_use;
      tc_fence_before();     // P is in TMEM (tcgen05.st completed above): order it before the MMA warp's tcgen05.mma
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // ------------------------------ epilogue ------------------------------
    float* xs = xchg + (nk & 1) * 256;
    xs[half * 128 + r] = l;
    pair_sync(quarter);
    l += xs[(half ^ 1) * 128 + r];
    mbar_wait(pv_done, (nk - 1) & 1);
    tc_fence_after();
    const int qv = qt * kTile + r;
    bf16* dst = nullptr;
    if (qv < g.Lv) dst = p.out_v + ((long long)b * g.Lv + qv) * g.D + h * kHeadDim;
    else if (qv >= g.T0 && qv < g.T0 + g.Lt) dst = p.out_t + ((long long)b * g.Lt + (qv - g.T0)) * g.D + h * kHeadDim;
    const float inv = 1.0f / l;
    {
      uint32_t v[32];
      tmem_ld32(tmO + lane_base + half * 32, v);
      tmem_wait_ld();
      if (dst != nullptr) {
        uint4* o = reinterpret_cast<uint4*>(dst + half * 32);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          o[q4] = make_uint4(pack_bf16x2(__uint_as_float(v[8 * q4]) * inv, __uint_as_float(v[8 * q4 + 1]) * inv),
                             pack_bf16x2(__uint_as_float(v[8 * q4 + 2]) * inv, __uint_as_float(v[8 * q4 + 3]) * inv),
                             pack_bf16x2(__uint_as_float(v[8 * q4 + 4]) * inv, __uint_as_float(v[8 * q4 + 5]) * inv),
                             pack_bf16x2(__uint_as_float(v[8 * q4 + 6]) * inv, __uint_as_float(v[8 * q4 + 7]) * inv));
      }
    }
    // positions of the virtual axis that are not tokens get LSE = +inf: the backward then computes P = exp2(S - inf) = 0
    // for them without any query-side masking
    if (p.lse != nullptr && half == 0)
      p.lse[((long long)b * g.H + h) * (g.ntiles * kTile) + qv] = dst != nullptr ? m + log2f(l) : INFINITY;
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
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
  {
    static const int dbg_env = [] { const char* e = getenv("SIMVGB_ATTN_DEBUG"); return e ? atoi(e) : 0; }();
    p.dbg = dbg_env;
    p.ts = g_fwd_trace;
  }
  CUtensorMap full, tail, text;
  if (make_attn_maps(&full, &tail, &text, p.g, a->qkv_v, a->qkv_t, 3 * D)) return -1;
  if (ensure_dynamic_smem(reinterpret_cast<const void*>(attn_fwd_kernel), kFwdSmem)) return -2;
  dim3 grid(p.g.ntiles, a->H, a->B);
  attn_fwd_kernel<<<grid, kFwdThreads, kFwdSmem, reinterpret_cast<cudaStream_t>(stream)>>>(full, tail, text, p);
  SIMVGB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int simvgb_attn_lse_stride(int Lv, int Lt) {
  simvgb::AttnGeom g = simvgb::make_attn_geom(1, 1, Lv, Lt, 64);
  return g.ntiles * simvgb::kTile;
}

extern "C" void simvgb_debug_attn_fwd_trace(long long* buf) { simvgb::g_fwd_trace = buf; }
