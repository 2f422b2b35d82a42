// Fused multiway self-attention backward for sm_100a (flash-style recompute; nothing of size L x L is stored).
//
// Replaces the autograd backward of torchscale MultiheadAttention's bmm/softmax/bmm
// (/root/reference/simvg/models/vis_encs/beit/beit3_base.py:137-145; driven by loss.backward() at
// /root/reference/simvg/apis/train.py:80).
//
// One CTA owns one 128-key tile j of one (b, h) and loops over the query tiles i:
//     S  = Q_i K_j^T                 P  = exp2(S*log2e - LSE_i)      (masked keys / rows -> 0)
//     dP = dO_i V_j^T                dS = P o (dP - delta_i)
//     dV_j += P^T dO_i               dK_j += dS^T Q_i                 dQ_i += dS K_j   (fp32 red.global.add)
// All five products are tcgen05.mma with accumulators in TMEM (S, dP: 128 columns each; dV, dK, dQ: 64 each).
// P and dS are written once to 128B-swizzled smem by the compute threads and consumed twice: K-major (dQ = dS K)
// and MN-major (P^T dO, dS^T Q) — the same bytes, two descriptors — so nothing is transposed.
// Warp roles: 0 TMA producer, 1 MMA issuer, 2-9 compute (two warps per TMEM lane quarter, 64 key columns each — two
// resident warps per SM sub-partition hide each other's MUFU/TMEM latency), 10-13 dQ write-back.
#include <stdlib.h>

#include "attn_common.cuh"

#ifdef SIMVGB_ATTN_ABLATE   // timing ablations (tools/attn_ablate.py): build with -DSIMVGB_ATTN_ABLATE
#define SIMVGB_DBG(p) ((p).dbg)
#else
#define SIMVGB_DBG(p) 0
#endif
#include "simvg_b200.h"

namespace simvgb {

int make_attn_maps(CUtensorMap* full, CUtensorMap* tail, CUtensorMap* text, const AttnGeom& g, const void* base_v,
                   const void* base_t, int row_elems);

constexpr int kBwdThreads = 704;        // TMA + MMA + 16 compute + 4 dQ warps
constexpr int kComputeThreads = 512;    // four warps per TMEM lane quarter, one 32-column chunk each
#ifndef SIMVGB_BWD_STAGES
#define SIMVGB_BWD_STAGES 2
#endif
constexpr int kQS = SIMVGB_BWD_STAGES;   // Q_i / dO_i ring depth (TMA latency is ~1.5 us: 2 stages cannot hide it)
constexpr int kBwdTiles = 2 + 2 * kQS + 8;   // K, V, Q[kQS], dO[kQS], P[2](2 sub-tiles), dS[2](2 sub-tiles)
constexpr int kBwdSmem = kBwdTiles * kTileBytes + 1024 + 256;
constexpr float kLog2eB = 1.4426950408889634f;

struct AttnBwdParams {
  AttnGeom g;
  const unsigned char* pad;
  const float* lse;
  const float* delta;
  bf16* dqkv_v;
  bf16* dqkv_t;
  float* dq_acc_v;
  float* dq_acc_t;
  int dbg;   // timing ablations (SIMVGB_ATTN_DEBUG); 0 in production
  long long* ts;   // optional clock64 trace of CTA (0,0,0): [pair][16]
};

static long long* g_attn_trace = nullptr;

struct AttnMaps6 {
  CUtensorMap qkv_full, qkv_tail, qkv_text, do_full, do_tail, do_text;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ AttnMaps6 maps, const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = smem + kTileBytes;
  uint8_t* sQ = smem + 2 * kTileBytes;                 // [kQS]
  uint8_t* sdO = smem + (2 + kQS) * kTileBytes;        // [kQS]
  uint8_t* sP = smem + (2 + 2 * kQS) * kTileBytes;     // [2 buffers][2 sub-tiles]
  uint8_t* sdS = smem + (6 + 2 * kQS) * kTileBytes;    // [2 buffers][2 sub-tiles]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBwdTiles * kTileBytes);
  uint64_t* kv_full = bars;
  uint64_t* qdo_full = bars + 1;           // [kQS]
  uint64_t* qdo_empty = bars + 1 + kQS;    // [kQS]
  uint64_t* s_full = bars + 1 + 2 * kQS;
  uint64_t* s_empty = s_full + 1;
  uint64_t* p_full = s_full + 2;
  uint64_t* pds_done = s_full + 3;   // [2]: one per P/dS buffer — a single barrier could advance two phases past a slow
                                     // waiter (parity aliasing) once the buffers are double-buffered
  uint64_t* dq_full = s_full + 5;
  uint64_t* dq_empty = s_full + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 7);
  uint32_t* kmask = tmem_slot + 2;  // [4]

  const AttnGeom& g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nq = g.ntiles;
  const int lse_stride = g.ntiles * kTile;

  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < kBwdTiles * kTileBytes / 16; i += kBwdThreads) z[i] = zero;
  }
  if (threadIdx.x == 0) {
    mbar_init(kv_full, 2);   // K_j and V_j arrive separately
    for (int s = 0; s < kQS; ++s) { mbar_init(&qdo_full[s], 2); mbar_init(&qdo_empty[s], 1); }  // Q_i + dO_i
    mbar_init(s_full, 1);
    mbar_init(s_empty, kComputeThreads / 32);   // one elected arrival per compute warp
    mbar_init(p_full, kComputeThreads / 32);
    mbar_init(&pds_done[0], 1);
    mbar_init(&pds_done[1], 1);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 4);
    fence_barrier_init();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 128) {   // warps 2-5
    if (kt >= g.nfull) build_tile_mask(kmask, g, p.pad, b, kt, threadIdx.x - 64);
    else if (threadIdx.x < 64 + 4) kmask[threadIdx.x - 64] = 0xffffffffu;
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // provably warp-uniform (uniform registers for UTCHMMA)
  const uint32_t tmS = tmem, tmdP = tmem + 128, tmdV = tmem + 256, tmdK = tmem + 320, tmdQ = tmem + 384;

  if (warp == 0) {
    if (lane == 0) {
      const int colq = h * kHeadDim, colk = g.D + h * kHeadDim, colv = 2 * g.D + h * kHeadDim;
      // K_j and V_j share one barrier (two expect_tx arrivals, init count 2); likewise Q_i and dO_i.
      load_virtual_tile(sK, kv_full, g, &maps.qkv_full, &maps.qkv_tail, &maps.qkv_text, kt, colk, b);
      for (int i = 0; i < nq; ++i) {
        const int s = i % kQS;
        mbar_wait(&qdo_empty[s], ((i / kQS) & 1) ^ 1);
        load_virtual_tile(sQ + s * kTileBytes, &qdo_full[s], g, &maps.qkv_full, &maps.qkv_tail, &maps.qkv_text, i,
                          colq, b);
        if (i == 0) load_virtual_tile(sV, kv_full, g, &maps.qkv_full, &maps.qkv_tail, &maps.qkv_text, kt, colv, b);
        load_virtual_tile(sdO + s * kTileBytes, &qdo_full[s], g, &maps.do_full, &maps.do_tail, &maps.do_text, i,
                          h * kHeadDim, b);
      }
    }
  } else if (warp == 1) {
    // MMA issuer: warp-uniform control flow, single-lane issue (keeps descriptors in uniform registers).
    const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
    const uint32_t idesc_dq = umma_idesc_bf16(128, kHeadDim, 0, 1);   // A = dS K-major, B = K_j MN-major
    const uint32_t idesc_dkv = umma_idesc_bf16(128, kHeadDim, 1, 1);  // A = P^T / dS^T MN-major, B = dO / Q MN-major
    const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
    const uint64_t dK_kmaj = umma_smem_desc(k_addr, 16, 1024), dV_kmaj = umma_smem_desc(v_addr, 16, 1024);
    const uint64_t dK_mn = umma_smem_desc(k_addr, 8192, 1024);
    const uint64_t dP_mn0 = umma_smem_desc(smem_u32(sP), kTileBytes, 1024), dS_mn0 = umma_smem_desc(smem_u32(sdS), kTileBytes, 1024);
    const uint64_t dS_kmaj0 = umma_smem_desc(smem_u32(sdS), 16, 1024);
    const uint64_t dQ_kmaj0 = umma_smem_desc(smem_u32(sQ), 16, 1024), dO_kmaj0 = umma_smem_desc(smem_u32(sdO), 16, 1024);
    const uint64_t dQ_mn0 = umma_smem_desc(smem_u32(sQ), 8192, 1024), dO_mn0 = umma_smem_desc(smem_u32(sdO), 8192, 1024);
    constexpr uint32_t kTileStep = kTileBytes >> 4;   // descriptor start-address units per 16 KB tile
    auto issue_sdp = [&](int i) {
      const int s = i % kQS;
      mbar_wait(&qdo_full[s], (i / kQS) & 1);
      tc_fence_after();
      const uint64_t dq = dQ_kmaj0 + s * kTileStep, dd = dO_kmaj0 + s * kTileStep;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(tmS, dq + 2 * k, dK_kmaj + 2 * k, idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(tmdP, dd + 2 * k, dV_kmaj + 2 * k, idesc_s, k > 0);
        umma_commit(s_full);
      }
      __syncwarp();
    };
    mbar_wait(kv_full, 0);
    issue_sdp(0);
    const bool trace = p.ts != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
    for (int i = 0; i < nq; ++i) {
      const int s = i % kQS;
      const int pb = i & 1;   // P / dS buffer of this pair
      mbar_wait(s_empty, i & 1);
      if (trace) p.ts[i * 16 + 0] = clock64();
      if (i + 1 < nq) issue_sdp(i + 1);
      if (trace) p.ts[i * 16 + 1] = clock64();
      mbar_wait(p_full, i & 1);
      if (trace) p.ts[i * 16 + 2] = clock64();
      if (i > 0) mbar_wait(dq_empty, (i - 1) & 1);
      if (trace) p.ts[i * 16 + 3] = clock64();
      tc_fence_after();
      const uint64_t dsk = dS_kmaj0 + pb * 2 * kTileStep, dsm = dS_mn0 + pb * 2 * kTileStep, dpm = dP_mn0 + pb * 2 * kTileStep;
      const uint64_t dqm = dQ_mn0 + s * kTileStep, dom = dO_mn0 + s * kTileStep;
      if (!(SIMVGB_DBG(p) & 16) && elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dQ_i = dS K_j        (A: dS K-major, B: K_j MN-major)
          umma_f16_ss(tmdQ, dsk + (k >> 2) * kTileStep + (k & 3) * 2, dK_mn + k * 128, idesc_dq, k > 0);
      }
      if (elect_one()) umma_commit(dq_full);
      if (!(SIMVGB_DBG(p) & 16) && elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dV_j += P^T dO_i     (A: P MN-major, B: dO_i MN-major)
          umma_f16_ss(tmdV, dpm + k * 128, dom + k * 128, idesc_dkv, (i > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dK_j += dS^T Q_i     (A: dS MN-major, B: Q_i MN-major)
          umma_f16_ss(tmdK, dsm + k * 128, dqm + k * 128, idesc_dkv, (i > 0 || k > 0) ? 1u : 0u);
      }
      if (elect_one()) {
        umma_commit(&qdo_empty[s]);
        umma_commit(&pds_done[pb]);
      }
      __syncwarp();
      if (trace) p.ts[i * 16 + 4] = clock64();
    }
  } else if (warp < 18) {
    // ------------------------------ compute: P and dS ------------------------------
    const int quarter = warp & 3;
    const int chunk = (warp - 2) >> 2;     // key columns [32*chunk, 32*chunk + 32)
    const int half = chunk;                // (chunks 0,1 also store dK columns [32*chunk, +32) at the end)
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = uint32_t(quarter * 32) << 16;
    const float* lse = p.lse + ((long long)b * g.H + h) * lse_stride;
    const float* delta = p.delta + ((long long)b * g.H + h) * lse_stride;
    const bool ktile_masked = kt >= g.nfull;
    for (int i = 0; i < nq; ++i) {
      const int qv = i * kTile + r;
      const bool row_ok = (qv < g.Lv) || (qv >= g.T0 && qv < g.T0 + g.Lt);
      const float L = row_ok ? __ldg(lse + qv) : 0.f;
      const float dl = row_ok ? __ldg(delta + qv) : 0.f;
      const bool slow = ktile_masked || (i >= g.nfull);   // warp-uniform: only tail tiles need masking
      const bool trace = p.ts != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == 2 && lane == 0;
      if (trace) p.ts[i * 16 + 8] = clock64();
      mbar_wait(s_full, i & 1);
      if (trace) p.ts[i * 16 + 9] = clock64();
      if (i > 1) mbar_wait(&pds_done[i & 1], ((i >> 1) + 1) & 1);   // MMAs of pair i-2 have finished reading this buffer
      if (trace) p.ts[i * 16 + 10] = clock64();
      const uint32_t aP = smem_u32(sP) + (i & 1) * 2 * kTileBytes;
      const uint32_t adS = smem_u32(sdS) + (i & 1) * 2 * kTileBytes;
      tc_fence_after();
      {
        const int c = chunk;
        uint32_t sv[32], dv[32];
        if (!(SIMVGB_DBG(p) & 4)) {
          tmem_ld32(tmS + lane_base + c * 32, sv);
          tmem_ld32(tmdP + lane_base + c * 32, dv);
          tmem_wait_ld();
          // this warp's share of S / dP is in registers: release it now so S_{i+1} / dP_{i+1} overlap the math below
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_empty);
        } else {
#pragma unroll
          for (int t = 0; t < 32; ++t) { sv[t] = 0; dv[t] = 0; }
        }
        float pr[32], ds[32];
        if (SIMVGB_DBG(p) & 2) {
#pragma unroll
          for (int t = 0; t < 32; ++t) { pr[t] = __uint_as_float(sv[t]); ds[t] = __uint_as_float(dv[t]); }
        } else if (!slow) {
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            const float pv = ex2_approx(fmaf(__uint_as_float(sv[t]), kLog2eB, -L));
            pr[t] = pv;
            ds[t] = pv * (__uint_as_float(dv[t]) - dl);
          }
        } else {
          const uint32_t bits = row_ok ? kmask[c] : 0u;
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            const float pv = (bits >> t) & 1u ? ex2_approx(fmaf(__uint_as_float(sv[t]), kLog2eB, -L)) : 0.f;
            pr[t] = pv;
            ds[t] = pv * (__uint_as_float(dv[t]) - dl);
          }
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const uint32_t off = swz_off(r, c * 4 + q4);
          st_shared_v4(aP + off, pack_bf16x2(pr[8 * q4], pr[8 * q4 + 1]), pack_bf16x2(pr[8 * q4 + 2], pr[8 * q4 + 3]),
                       pack_bf16x2(pr[8 * q4 + 4], pr[8 * q4 + 5]), pack_bf16x2(pr[8 * q4 + 6], pr[8 * q4 + 7]));
          st_shared_v4(adS + off, pack_bf16x2(ds[8 * q4], ds[8 * q4 + 1]), pack_bf16x2(ds[8 * q4 + 2], ds[8 * q4 + 3]),
                       pack_bf16x2(ds[8 * q4 + 4], ds[8 * q4 + 5]), pack_bf16x2(ds[8 * q4 + 6], ds[8 * q4 + 7]));
        }
      }
      if (trace) p.ts[i * 16 + 11] = clock64();
      tc_fence_before();
      fence_proxy_async();   // this thread's P/dS stores -> async proxy
      __syncwarp();
      if (lane == 0) {
        if (SIMVGB_DBG(p) & 4) mbar_arrive(s_empty);
        mbar_arrive(p_full);
      }
      if (trace) p.ts[i * 16 + 12] = clock64();
    }
    // dK_j -> dqkv[:, D + h*64 ...]   (each of the two warps of a quarter stores one 32-column half)
    if (nq > 1) mbar_wait(&pds_done[nq & 1], ((nq - 2) >> 1) & 1);          // pair nq-2 ...
    mbar_wait(&pds_done[(nq - 1) & 1], ((nq - 1) >> 1) & 1);              // ... and the last pair have retired
    tc_fence_after();
    const int kv = kt * kTile + r;
    bf16* dst = nullptr;
    if (kv < g.Lv) dst = p.dqkv_v + ((long long)b * g.Lv + kv) * (3 * g.D) + g.D + h * kHeadDim;
    else if (kv >= g.T0 && kv < g.T0 + g.Lt) dst = p.dqkv_t + ((long long)b * g.Lt + (kv - g.T0)) * (3 * g.D) + g.D + h * kHeadDim;
    if (chunk < 2) {
      uint32_t v[32];
      tmem_ld32(tmdK + lane_base + half * 32, v);
      tmem_wait_ld();
      if (dst != nullptr) {
        uint4* o = reinterpret_cast<uint4*>(dst + half * 32);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          o[q4] = make_uint4(pack_bf16x2(__uint_as_float(v[8 * q4]), __uint_as_float(v[8 * q4 + 1])),
                             pack_bf16x2(__uint_as_float(v[8 * q4 + 2]), __uint_as_float(v[8 * q4 + 3])),
                             pack_bf16x2(__uint_as_float(v[8 * q4 + 4]), __uint_as_float(v[8 * q4 + 5])),
                             pack_bf16x2(__uint_as_float(v[8 * q4 + 6]), __uint_as_float(v[8 * q4 + 7])));
      }
    }
    tc_fence_before();
  } else {
    // ------------------------------ dQ write-back (fp32 atomics) and dV ------------------------------
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = uint32_t(quarter * 32) << 16;
    for (int i = 0; i < nq; ++i) {
      const int qv = i * kTile + r;
      float* dst = nullptr;
      if (qv < g.Lv) dst = p.dq_acc_v + ((long long)b * g.Lv + qv) * g.D + h * kHeadDim;
      else if (qv >= g.T0 && qv < g.T0 + g.Lt) dst = p.dq_acc_t + ((long long)b * g.Lt + (qv - g.T0)) * g.D + h * kHeadDim;
      mbar_wait(dq_full, i & 1);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld32(tmdQ + lane_base, v0);
      tmem_ld32(tmdQ + lane_base + 32, v1);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_empty);
      // Lane pairs (2k, 2k+1) swap half of their row so that each red.v4 instruction has the two lanes of a pair writing
      // adjacent 16-byte chunks of the SAME row: every L2 atomic operation then covers a full 32-byte sector (half as
      // many L2 atomic operations as one-row-per-lane).  even lane keeps float4 #0,2,4,6 of its row and receives the same
      // of the odd lane's row; the odd lane keeps / receives float4 #1,3,5,7.
      if (!(SIMVGB_DBG(p) & 1)) {
        const bool odd = lane & 1;
        const unsigned long long my = reinterpret_cast<unsigned long long>(dst);
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, my, 1);
        float* row_e = reinterpret_cast<float*>(odd ? other : my);   // row owned by the even lane of the pair
        float* row_o = reinterpret_cast<float*>(odd ? my : other);   // row owned by the odd lane
        auto flush_half = [&](const uint32_t (&v)[32], int col0) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float recv[4], own[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              // float4 indices 2q (even) and 2q+1 (odd) of this 32-column half: send the one the partner keeps
              const uint32_t ev = v[4 * (2 * q) + e], ov = v[4 * (2 * q + 1) + e];
              recv[e] = __uint_as_float(__shfl_xor_sync(0xffffffffu, odd ? ev : ov, 1));
              own[e] = __uint_as_float(odd ? ov : ev);
            }
            const int col = col0 + 4 * (2 * q + (odd ? 1 : 0));
            // instruction A: the even lane's row; instruction B: the odd lane's row
            if (row_e != nullptr) {
              if (odd) red_add_v4(row_e + col, recv[0], recv[1], recv[2], recv[3]);
              else red_add_v4(row_e + col, own[0], own[1], own[2], own[3]);
            }
            if (row_o != nullptr) {
              if (odd) red_add_v4(row_o + col, own[0], own[1], own[2], own[3]);
              else red_add_v4(row_o + col, recv[0], recv[1], recv[2], recv[3]);
            }
          }
        };
        flush_half(v0, 0);
        flush_half(v1, 32);
      }
    }
    if (nq > 1) mbar_wait(&pds_done[nq & 1], ((nq - 2) >> 1) & 1);          // pair nq-2 ...
    mbar_wait(&pds_done[(nq - 1) & 1], ((nq - 1) >> 1) & 1);              // ... and the last pair have retired
    tc_fence_after();
    const int kv = kt * kTile + r;
    bf16* dst = nullptr;
    if (kv < g.Lv) dst = p.dqkv_v + ((long long)b * g.Lv + kv) * (3 * g.D) + 2 * g.D + h * kHeadDim;
    else if (kv >= g.T0 && kv < g.T0 + g.Lt) dst = p.dqkv_t + ((long long)b * g.Lt + (kv - g.T0)) * (3 * g.D) + 2 * g.D + h * kHeadDim;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld32(tmdV + lane_base + c * 32, v);
      tmem_wait_ld();
      if (dst != nullptr) {
        uint4* o = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          o[q4] = make_uint4(pack_bf16x2(__uint_as_float(v[8 * q4]), __uint_as_float(v[8 * q4 + 1])),
                             pack_bf16x2(__uint_as_float(v[8 * q4 + 2]), __uint_as_float(v[8 * q4 + 3])),
                             pack_bf16x2(__uint_as_float(v[8 * q4 + 4]), __uint_as_float(v[8 * q4 + 5])),
                             pack_bf16x2(__uint_as_float(v[8 * q4 + 6]), __uint_as_float(v[8 * q4 + 7])));
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// delta[b, h, virtual q] = sum_d O * dO   (one thread per (token, head))
__global__ void attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o, float* __restrict__ delta,
                                  int B, int H, int L, int D, int vbase, int lse_stride) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * L * H) return;
  const int h = idx % H;
  const long long row = idx / H;
  const int l = row % L, b = row / L;
  const uint4* po = reinterpret_cast<const uint4*>(o + row * D + h * kHeadDim);
  const uint4* pd = reinterpret_cast<const uint4*>(d_o + row * D + h * kHeadDim);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint4 a = __ldg(po + i), c = __ldg(pd + i);
    const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* c2 = reinterpret_cast<const __nv_bfloat162*>(&c);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 x = __bfloat1622float2(a2[k]), y = __bfloat1622float2(c2[k]);
      acc += x.x * y.x + x.y * y.y;
    }
  }
  delta[((long long)b * H + h) * lse_stride + vbase + l] = acc;
}

// dqkv[:, 0:D] = bf16(scale * dq_acc)
__global__ void attn_dq_convert_kernel(const float* __restrict__ acc, bf16* __restrict__ dqkv, long long rows, int D,
                                       float scale) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per 8 elements
  const int per_row = D / 8;
  if (idx >= rows * per_row) return;
  const long long row = idx / per_row;
  const int c = (idx % per_row) * 8;
  const float4 a = __ldg(reinterpret_cast<const float4*>(acc + row * D + c));
  const float4 b2 = __ldg(reinterpret_cast<const float4*>(acc + row * D + c + 4));
  *reinterpret_cast<uint4*>(dqkv + row * 3 * D + c) =
      make_uint4(pack_bf16x2(a.x * scale, a.y * scale), pack_bf16x2(a.z * scale, a.w * scale),
                 pack_bf16x2(b2.x * scale, b2.y * scale), pack_bf16x2(b2.z * scale, b2.w * scale));
}

}  // namespace simvgb

extern "C" int simvgb_attn_bwd(const simvgb_attn_args* a, void* stream) {
  using namespace simvgb;
  SIMVGB_CHECK(a != nullptr, "simvgb_attn_bwd: null args");
  SIMVGB_CHECK(a->head_dim == kHeadDim, "simvgb_attn_bwd: head_dim must be 64 (got %d)", a->head_dim);
  SIMVGB_CHECK(a->Lt <= 120, "simvgb_attn_bwd: at most 120 text tokens (got %d)", a->Lt);
  SIMVGB_CHECK(a->qkv_v && a->out_v && a->dout_v && a->dqkv_v && a->lse && a->delta && a->dq_acc_v,
               "simvgb_attn_bwd: null buffer");
  SIMVGB_CHECK(a->Lt == 0 || (a->qkv_t && a->out_t && a->dout_t && a->dqkv_t && a->dq_acc_t),
               "simvgb_attn_bwd: null text buffer");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int D = a->H * kHeadDim;
  AttnBwdParams p;
  p.g = make_attn_geom(a->B, a->H, a->Lv, a->Lt, D);
  p.pad = reinterpret_cast<const unsigned char*>(a->text_pad);
  p.lse = a->lse;
  p.delta = a->delta;
  p.dqkv_v = reinterpret_cast<bf16*>(a->dqkv_v);
  p.dqkv_t = reinterpret_cast<bf16*>(a->dqkv_t);
  p.dq_acc_v = a->dq_acc_v;
  p.dq_acc_t = a->dq_acc_t;
  {
    static const int dbg_env = [] { const char* e = getenv("SIMVGB_ATTN_DEBUG"); return e ? atoi(e) : 0; }();
    p.dbg = dbg_env;          // timing ablations for tools/attn_ablate.py; 0 in production
    p.ts = g_attn_trace;      // optional clock64 trace (tools/attn_trace.py)
  }
  const int lse_stride = p.g.ntiles * kTile;
  AttnMaps6 maps;
  if (make_attn_maps(&maps.qkv_full, &maps.qkv_tail, &maps.qkv_text, p.g, a->qkv_v, a->qkv_t, 3 * D)) return -1;
  if (make_attn_maps(&maps.do_full, &maps.do_tail, &maps.do_text, p.g, a->dout_v, a->dout_t, D)) return -1;

  {
    const long long n = (long long)a->B * a->Lv * a->H;
    attn_delta_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(reinterpret_cast<const bf16*>(a->out_v),
                                                                   reinterpret_cast<const bf16*>(a->dout_v), a->delta,
                                                                   a->B, a->H, a->Lv, D, 0, lse_stride);
    if (a->Lt > 0) {
      const long long nt = (long long)a->B * a->Lt * a->H;
      attn_delta_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, s>>>(reinterpret_cast<const bf16*>(a->out_t),
                                                                      reinterpret_cast<const bf16*>(a->dout_t), a->delta,
                                                                      a->B, a->H, a->Lt, D, p.g.T0, lse_stride);
    }
  }
  SIMVGB_CUDA(cudaMemsetAsync(a->dq_acc_v, 0, sizeof(float) * (size_t)a->B * a->Lv * D, s));
  if (a->Lt > 0) SIMVGB_CUDA(cudaMemsetAsync(a->dq_acc_t, 0, sizeof(float) * (size_t)a->B * a->Lt * D, s));

  static bool attr_set = false;
  if (!attr_set) {
    SIMVGB_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
    attr_set = true;
  }
  dim3 grid(p.g.ntiles, a->H, a->B);
  attn_bwd_kernel<<<grid, kBwdThreads, kBwdSmem, s>>>(maps, p);
  SIMVGB_CUDA(cudaGetLastError());
  {
    const long long rows = (long long)a->B * a->Lv;
    const long long n = rows * (D / 8);
    attn_dq_convert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a->dq_acc_v, p.dqkv_v, rows, D, a->q_scale);
    if (a->Lt > 0) {
      const long long rt = (long long)a->B * a->Lt;
      const long long nt = rt * (D / 8);
      attn_dq_convert_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, s>>>(a->dq_acc_t, p.dqkv_t, rt, D, a->q_scale);
    }
  }
  SIMVGB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" void simvgb_debug_attn_trace(long long* buf) { simvgb::g_attn_trace = buf; }
