// Fused multiway self-attention backward for sm_100a (flash-style recompute; nothing of size L x L is stored).
//
// Replaces the autograd backward of torchscale MultiheadAttention's bmm/softmax/bmm
// (/root/reference/simvg/models/vis_encs/beit/beit3_base.py:137-145; driven by loss.backward() at
// /root/reference/simvg/apis/train.py:80).
//
// One CTA owns one 128-key tile j of one (b, h) and loops over the query tiles i.  Scores are computed TRANSPOSED
// (keys on the TMEM lanes, queries along the columns) so that P^T and dS^T can feed the dV / dK products straight from
// tensor memory as the A operand — at head_dim 64 the kernel is bound by shared-memory bandwidth (every SS-form
// tcgen05.mma streams both operands from smem), and this removes P entirely and half of dS from shared memory:
//     S^T  = K_j Q_i^T - LSE_i[q]/log2e      P^T  = exp2(S^T*log2e)      (masked keys -> 0; non-token queries: LSE -> huge)
//     dP^T = V_j dO_i^T - delta_i[q]         dS^T = P^T o dP^T
// The per-query statistics are per-COLUMN constants in this orientation; rather than have every compute thread fetch
// them (one shared-memory wavefront per value per warp — more traffic than the operands), they are folded into the
// accumulators by one extra K=16 tcgen05.mma each:  ones[keys,16] x stat[q,16]^T with the statistic split into three bf16
// terms (hi + mid + lo, exact to fp32 rounding) in columns 0-2 of a 32-byte K-slot written by the producer warp.
//     dV_j += P^T dO_i    (A = P^T  from TMEM)
//     dK_j += dS^T Q_i    (A = dS^T from TMEM, written in place over dP^T)
//     dQ_i  = dS K_j      (A = dS^T tile in smem read MN-major, B = K_j MN-major; the fp32 tile is staged in shared memory and
//                          added to the global accumulator by ONE bulk tensor reduction per 32-column half
//                          (cp.reduce.async.bulk.tensor .add.f32): round 1 issued 8192 red.global.add.v4 per tile pair from the
//                          LSU, 820 of the ~3000 L1 data-pipe wavefronts per pair — the pipe that also feeds the MMA operands)
// TMEM columns: S^T [0,128)  dP^T/dS^T [128,256)  dV [256,320)  dK [320,384)  dQ [384,448)  P^T (bf16 pairs) [448,512).
// Tensor-pipe order per pair: S(i+1), dK(i), dP(i+1), dV(i), dQ(i) — in-order execution makes the in-place dS^T -> dP^T
// hand-over safe without a round trip through the issuing warp.
// Warp roles: 0 TMA producer (Q_i, dO_i ring) + statistic K-slots, 1 MMA issuer, 2-17 compute (four warps per TMEM lane
// quarter, 32 query columns each), 18-21 dQ write-back.
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
#ifndef SIMVGB_BWD_POLY
#define SIMVGB_BWD_POLY 0
#endif
#ifndef SIMVGB_BWD_STAGES
#define SIMVGB_BWD_STAGES 3
#endif
constexpr int kQS = SIMVGB_BWD_STAGES;   // Q_i / dO_i / LSE_i / delta_i ring depth (TMA latency is ~1.5 us)
constexpr int kStatSlots = 2 * kQS + 1;                  // 32-byte K-slots: (LSE, delta) per stage + one all-ones A slot
constexpr int kStatTiles = (kStatSlots + 3) / 4;         // four K-slots per 128B-swizzled [128 x 64] tile
constexpr int kBwdTiles = 2 + 2 * kQS + 4 + kStatTiles;  // K, V, Q[kQS], dO[kQS], dS^T (2 sub-tiles), dQ staging (2 fp32 sub-tiles), statistic slots
constexpr int kStageBytes = 2 * 2 * kTile * 4;           // raw LSE_i / delta_i rows (bulk-copied, 2 deep), converted into K-slots by warp 0
// 14 tiles + staging + barriers = 231,680 B of the 232,448 B a CTA may own: no slack for manual alignment, so the dynamic
// shared-memory window itself must be 1024-byte aligned (it is when the kernel has no static __shared__; checked at entry).
constexpr int kBwdSmem = kBwdTiles * kTileBytes + kStageBytes + 256;
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

struct AttnMaps9 {
  CUtensorMap qkv_full, qkv_tail, qkv_text, do_full, do_tail, do_text;
  CUtensorMap dq_full, dq_tail, dq_text;   // fp32 dQ accumulators, 32-column boxes (reduction targets)
};

// fp32 tile (shared memory, 128B-swizzled box of the tensor map) += into global memory: one TMA reduction, line-granular at L2
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* m, uint32_t src_smem, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src_smem), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void dq_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the four dQ write-back warps
// 1-D bulk copy global -> shared, completion on an mbarrier (bytes: multiple of 16, both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ AttnMaps9 maps, const AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();         // 128B-swizzled TMA / UMMA tiles need 1024-byte alignment
  uint8_t* sK = smem;
  uint8_t* sV = smem + kTileBytes;
  uint8_t* sQ = smem + 2 * kTileBytes;                 // [kQS]
  uint8_t* sdO = smem + (2 + kQS) * kTileBytes;        // [kQS]
  uint8_t* sdS = smem + (2 + 2 * kQS) * kTileBytes;    // [2 sub-tiles]: rows = keys, 128 queries per row (single buffer: dQ(i) has
                                                       // long retired when the compute warps reach the dS^T store of pair i+1)
  uint8_t* sdQ = smem + (4 + 2 * kQS) * kTileBytes;    // [2 sub-tiles]: 128 query rows x 32 fp32 columns each, 128B-swizzled
  uint8_t* sStat = smem + (6 + 2 * kQS) * kTileBytes;  // K-slot n: rows 128 B apart, 16-byte chunks 2n', 2n'+1 (n' = n % 4) of tile n / 4
  float* sStage = reinterpret_cast<float*>(smem + kBwdTiles * kTileBytes);   // [2][2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBwdTiles * kTileBytes + kStageBytes);
  uint64_t* kv_full = bars;
  uint64_t* qdo_full = bars + 1;           // [kQS]
  uint64_t* qdo_empty = bars + 1 + kQS;    // [kQS]
  uint64_t* s_full = bars + 1 + 2 * kQS;   // S^T(i) complete
  uint64_t* dp_full = s_full + 1;          // dP^T(i) complete
  uint64_t* s_empty = s_full + 2;          // every compute warp has S^T(i) in registers
  uint64_t* p_full = s_full + 3;           // P^T(i), dS^T(i) (TMEM) and dS^T(i) (smem) written
  uint64_t* pt_free = s_full + 4;          // dV(i) has consumed P^T(i)
  uint64_t* ds_free = s_full + 5;          // [0]: dQ(i) has consumed the smem dS^T tile ([1] unused)
  uint64_t* dq_full = s_full + 7;
  uint64_t* dq_empty = s_full + 8;
  uint64_t* mma_done = s_full + 9;
  uint64_t* st_full = s_full + 10;         // [2]: raw statistics have landed in the staging buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 12);
  uint32_t* kmask = tmem_slot + 2;  // [4]

  const AttnGeom& g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nq = g.ntiles;
  const int lse_stride = g.ntiles * kTile;

  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < kBwdTiles * kTileBytes / 16; i += kBwdThreads) z[i] = zero;   // staging tiles, statistic slots
  }
  if (threadIdx.x == 0) {
    mbar_init(kv_full, 2);   // K_j and V_j arrive separately
    for (int s = 0; s < kQS; ++s) { mbar_init(&qdo_full[s], 3); mbar_init(&qdo_empty[s], 1); }  // Q_i + dO_i + stats
    mbar_init(s_full, 1);
    mbar_init(dp_full, 1);
    mbar_init(s_empty, kComputeThreads / 32);   // one elected arrival per compute warp
    mbar_init(p_full, kComputeThreads / 32);
    mbar_init(pt_free, 1);
    mbar_init(&ds_free[0], 1);
    mbar_init(&ds_free[1], 1);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 4);
    mbar_init(mma_done, 1);
    mbar_init(&st_full[0], 1);
    mbar_init(&st_full[1], 1);
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
  const uint32_t tmS = tmem, tmdP = tmem + 128, tmdV = tmem + 256, tmdK = tmem + 320, tmdQ = tmem + 384, tmP = tmem + 448;

  if (warp == 0) {
    const int colq = h * kHeadDim, colk = g.D + h * kHeadDim, colv = 2 * g.D + h * kHeadDim;
    const float* lse = p.lse + ((long long)b * g.H + h) * lse_stride;
    const float* delta = p.delta + ((long long)b * g.H + h) * lse_stride;
    // Writes 16 bf16 (K elements 0..15 of K-slot `slot`) of row `row`: {hi, mid, lo, 0 ...} with hi + mid + lo == v to fp32 rounding.
    auto put_stat = [&](int slot, int row, float v) {
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      const float r1 = v - __bfloat162float(hi);
      const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
      const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
      const uint32_t w0 = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(mid) << 16);
      const uint32_t w1 = (uint32_t)__bfloat16_as_ushort(lo);
      const uint32_t base = smem_u32(sStat) + (slot >> 2) * kTileBytes + row * 128;
      const int c0 = 2 * (slot & 3);
      st_shared_v4(base + (((c0) ^ (row & 7)) << 4), w0, w1, 0u, 0u);
      st_shared_v4(base + (((c0 + 1) ^ (row & 7)) << 4), 0u, 0u, 0u, 0u);
    };
    // all-ones A slot (every key row contributes 1 x statistic)
    for (int row = lane; row < kTile; row += 32) {
      const uint32_t base = smem_u32(sStat) + ((2 * kQS) >> 2) * kTileBytes + row * 128;
      const int c0 = 2 * ((2 * kQS) & 3);
      st_shared_v4(base + (((c0) ^ (row & 7)) << 4), 0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
      st_shared_v4(base + (((c0 + 1) ^ (row & 7)) << 4), 0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
    }
    // K_j and V_j share one barrier (two expect_tx arrivals, init count 2); Q_i, dO_i and the statistic slots share one (3).
    // The raw LSE / delta rows are bulk-copied (TMA path: plain global loads from this warp would queue behind the dQ
    // reductions in the LSU) into a 2-deep staging ring one iteration before they are converted into K-slots.
    auto convert_stats = [&](int i) {
      const int s = i % kQS;
      mbar_wait(&st_full[i & 1], (i >> 1) & 1);
      const float* raw = sStage + (i & 1) * 2 * kTile;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int row = lane + 32 * t, qv = i * kTile + row;
        const bool ok = (qv < g.Lv) || (qv >= g.T0 && qv < g.T0 + g.Lt);
        put_stat(2 * s, row, ok ? -raw[row] * 0.6931471805599453f : -60000.f);   // natural-log units: exp2((s + v) * log2e)
        put_stat(2 * s + 1, row, ok ? -raw[kTile + row] : 0.f);
      }
      fence_proxy_async();   // statistic slots (generic proxy) -> tcgen05.mma (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&qdo_full[s]);
    };
    auto load_stats = [&](int i) {   // staging buffer i & 1 was drained by convert_stats(i - 2) (program order + its proxy fence)
      mbar_expect_tx(&st_full[i & 1], 2 * kTile * 4);
      bulk_load(sStage + (i & 1) * 2 * kTile, lse + i * kTile, kTile * 4, &st_full[i & 1]);
      bulk_load(sStage + (i & 1) * 2 * kTile + kTile, delta + i * kTile, kTile * 4, &st_full[i & 1]);
    };
    if (lane == 0) {
      load_virtual_tile(sK, kv_full, g, &maps.qkv_full, &maps.qkv_tail, &maps.qkv_text, kt, colk, b);
      load_stats(0);
    }
    for (int i = 0; i < nq; ++i) {
      const int s = i % kQS;
      if (lane == 0 && i + 1 < nq) load_stats(i + 1);   // a full iteration ahead of its conversion
      mbar_wait(&qdo_empty[s], ((i / kQS) & 1) ^ 1);
      if (lane == 0) {
        load_virtual_tile(sQ + s * kTileBytes, &qdo_full[s], g, &maps.qkv_full, &maps.qkv_tail, &maps.qkv_text, i,
                          colq, b);
        if (i == 0) load_virtual_tile(sV, kv_full, g, &maps.qkv_full, &maps.qkv_tail, &maps.qkv_text, kt, colv, b);
        load_virtual_tile(sdO + s * kTileBytes, &qdo_full[s], g, &maps.do_full, &maps.do_tail, &maps.do_text, i,
                          h * kHeadDim, b);
      }
      __syncwarp();
      convert_stats(i);   // raw rows landed long ago; the K-slots of stage s are free (qdo_empty above)
    }
  } else if (warp == 1) {
    // MMA issuer: warp-uniform control flow, single-lane issue (keeps descriptors in uniform registers).
    const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);         // S^T / dP^T: A = K_j / V_j, B = Q_i / dO_i, all K-major
    const uint32_t idesc_ts = umma_idesc_bf16(128, kHeadDim, 0, 1);   // dV / dK: A from TMEM, B = dO_i / Q_i MN-major
    const uint32_t idesc_dq = umma_idesc_bf16(128, kHeadDim, 1, 1);   // dQ: A = dS^T tile MN-major, B = K_j MN-major
    const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
    const uint64_t dK_kmaj = umma_smem_desc(k_addr, 16, 1024), dV_kmaj = umma_smem_desc(v_addr, 16, 1024);
    const uint64_t dK_mn = umma_smem_desc(k_addr, 8192, 1024);
    const uint64_t dS_mn0 = umma_smem_desc(smem_u32(sdS), kTileBytes, 1024);
    const uint64_t dQ_kmaj0 = umma_smem_desc(smem_u32(sQ), 16, 1024), dO_kmaj0 = umma_smem_desc(smem_u32(sdO), 16, 1024);
    const uint64_t dQ_mn0 = umma_smem_desc(smem_u32(sQ), 8192, 1024), dO_mn0 = umma_smem_desc(smem_u32(sdO), 8192, 1024);
    constexpr uint32_t kTileStep = kTileBytes >> 4;   // descriptor start-address units per 16 KB tile
    const uint64_t dStat0 = umma_smem_desc(smem_u32(sStat), 16, 1024);
    auto stat_desc = [&](int slot) { return dStat0 + (slot >> 2) * kTileStep + 2 * (slot & 3); };
    const uint64_t dOnes = stat_desc(2 * kQS);
    auto issue_s = [&](int i) {     // S^T(i) = K_j Q_i^T - LSE_i[q] / log2e
      const uint64_t dq = dQ_kmaj0 + (i % kQS) * kTileStep, dst = stat_desc(2 * (i % kQS));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(tmS, dK_kmaj + 2 * k, dq + 2 * k, idesc_s, k > 0);
        umma_f16_ss(tmS, dOnes, dst, idesc_s, 1u);
        umma_commit(s_full);
      }
      __syncwarp();
    };
    auto issue_dp = [&](int i) {    // dP^T(i) = V_j dO_i^T - delta_i[q]
      const uint64_t dd = dO_kmaj0 + (i % kQS) * kTileStep, dst = stat_desc(2 * (i % kQS) + 1);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(tmdP, dV_kmaj + 2 * k, dd + 2 * k, idesc_s, k > 0);
        umma_f16_ss(tmdP, dOnes, dst, idesc_s, 1u);
        umma_commit(dp_full);
      }
      __syncwarp();
    };
    const bool trace = p.ts != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
    mbar_wait(kv_full, 0);
    mbar_wait(&qdo_full[0], 0);
    tc_fence_after();
    issue_s(0);
    issue_dp(0);
    for (int i = 0; i < nq; ++i) {
      const int s = i % kQS;
      if (trace) p.ts[i * 16 + 0] = clock64();
      if (i + 1 < nq) {       // S^T(i+1) runs on the tensor pipe while the compute warps work on pair i
        mbar_wait(&qdo_full[(i + 1) % kQS], ((i + 1) / kQS) & 1);
        mbar_wait(s_empty, i & 1);
        tc_fence_after();
        issue_s(i + 1);
      }
      if (trace) p.ts[i * 16 + 1] = clock64();
      mbar_wait(p_full, i & 1);
      if (trace) p.ts[i * 16 + 2] = clock64();
      tc_fence_after();
      const uint64_t dqm = dQ_mn0 + s * kTileStep, dom = dO_mn0 + s * kTileStep;
      if (!(SIMVGB_DBG(p) & 16) && elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dK_j += dS^T Q_i     (A: dS^T in TMEM, 16 packed columns per 32-query chunk)
          umma_f16_ts(tmdK, tmdP + 32 * (k >> 1) + 8 * (k & 1), dqm + k * 128, idesc_ts, (i > 0 || k > 0) ? 1u : 0u);
      }
      __syncwarp();
      if (i + 1 < nq) issue_dp(i + 1);   // overwrites dS^T(i): the tensor pipe executes in issue order, after dK(i)
      if (!(SIMVGB_DBG(p) & 16) && elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dV_j += P^T dO_i     (A: P^T in TMEM)
          umma_f16_ts(tmdV, tmP + 8 * k, dom + k * 128, idesc_ts, (i > 0 || k > 0) ? 1u : 0u);
      }
      if (elect_one()) {
        umma_commit(pt_free);
        umma_commit(&qdo_empty[s]);   // S(i), dP(i), dK(i), dV(i) were the readers of this Q_i / dO_i / statistics stage
      }
      __syncwarp();
      if (trace) p.ts[i * 16 + 3] = clock64();
      if (i > 0) mbar_wait(dq_empty, (i - 1) & 1);
      if (trace) p.ts[i * 16 + 4] = clock64();
      tc_fence_after();
      const uint64_t dsm = dS_mn0;
      if (!(SIMVGB_DBG(p) & 16) && elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dQ_i = dS K_j        (A: smem dS^T tile read MN-major, B: K_j MN-major)
          umma_f16_ss(tmdQ, dsm + k * 128, dK_mn + k * 128, idesc_dq, k > 0);
      }
      if (elect_one()) {
        umma_commit(dq_full);
        umma_commit(&ds_free[0]);
        if (i == nq - 1) umma_commit(mma_done);
      }
      __syncwarp();
      if (trace) p.ts[i * 16 + 5] = clock64();
    }
  } else if (warp < 18) {
    // ------------------------------ compute: P^T and dS^T ------------------------------
    const int quarter = warp & 3;
    const int c = (warp - 2) >> 2;         // query columns [32c, 32c + 32) of the pair
    const int r = quarter * 32 + lane;     // key row of this thread
    const uint32_t lane_base = uint32_t(quarter * 32) << 16;
    const bool ktile_masked = kt >= g.nfull;
    const bool kvalid = (kmask[quarter] >> lane) & 1u;
    const bool trace = p.ts != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == 2 && lane == 0;
    for (int i = 0; i < nq; ++i) {
      if (trace) p.ts[i * 16 + 8] = clock64();
      mbar_wait(s_full, i & 1);
      if (trace) p.ts[i * 16 + 9] = clock64();
      tc_fence_after();
      uint32_t sv[32];
      tmem_ld32(tmS + lane_base + c * 32, sv);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);   // S^T(i) is in registers: S^T(i+1) may overwrite it
      // P^T as bf16 pairs (TMEM column t of the chunk holds queries 2t, 2t+1).  Only the packed form stays live across the
      // wait for dP^T: dS^T below uses the bf16-rounded P^T, the value the dV product sees.
      uint32_t pk[16];
#pragma unroll
      for (int t = 0; t < 16; ++t) {
#if SIMVGB_BWD_POLY
        // every other exponential on the FMA pipe (the P phase is otherwise bound by the 16/clk/SM MUFU unit)
        pk[t] = pack_bf16x2(ex2_approx(__uint_as_float(sv[2 * t]) * kLog2eB),
                            ex2_fma(fmaxf(__uint_as_float(sv[2 * t + 1]) * kLog2eB, -120.f)));
#else
        pk[t] = pack_bf16x2(ex2_approx(__uint_as_float(sv[2 * t]) * kLog2eB), ex2_approx(__uint_as_float(sv[2 * t + 1]) * kLog2eB));
#endif
      }
      if (ktile_masked && !kvalid) {
#pragma unroll
        for (int t = 0; t < 16; ++t) pk[t] = 0u;
      }
      if (i > 0) mbar_wait(pt_free, (i - 1) & 1);   // dV(i-1) has consumed P^T(i-1)
      tc_fence_after();
      tmem_st16(tmP + lane_base + 16 * c, pk);
      if (trace) p.ts[i * 16 + 10] = clock64();
      mbar_wait(dp_full, i & 1);
      if (trace) p.ts[i * 16 + 11] = clock64();
      tc_fence_after();
      uint32_t dv[32];
      tmem_ld32(tmdP + lane_base + c * 32, dv);
      tmem_wait_ld();
      uint32_t dk[16];
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const float p0 = __uint_as_float(pk[t] << 16), p1 = __uint_as_float(pk[t] & 0xffff0000u);
        dk[t] = pack_bf16x2(p0 * __uint_as_float(dv[2 * t]), p1 * __uint_as_float(dv[2 * t + 1]));
      }
      tmem_st16(tmdP + lane_base + 32 * c, dk);   // dS^T in place over this thread's own (already loaded) dP^T columns
      if (i > 0) mbar_wait(&ds_free[0], (i - 1) & 1);   // dQ(i-1) has consumed the smem dS^T tile
      const uint32_t adS = smem_u32(sdS);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4)
        st_shared_v4(adS + swz_off(r, c * 4 + q4), dk[4 * q4], dk[4 * q4 + 1], dk[4 * q4 + 2], dk[4 * q4 + 3]);
      tmem_wait_st();
      tc_fence_before();
      fence_proxy_async();   // this thread's dS^T smem stores -> async proxy (dQ MMA)
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (trace) p.ts[i * 16 + 12] = clock64();
    }
    // dK_j -> dqkv[:, D + h*64 ...]   (warps with c < 2 store one 32-column half each)
    mbar_wait(mma_done, 0);
    tc_fence_after();
    const int kv = kt * kTile + r;
    bf16* dst = nullptr;
    if (kv < g.Lv) dst = p.dqkv_v + ((long long)b * g.Lv + kv) * (3 * g.D) + g.D + h * kHeadDim;
    else if (kv >= g.T0 && kv < g.T0 + g.Lt) dst = p.dqkv_t + ((long long)b * g.Lt + (kv - g.T0)) * (3 * g.D) + g.D + h * kHeadDim;
    if (c < 2) {
      uint32_t v[32];
      tmem_ld32(tmdK + lane_base + c * 32, v);
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
  } else {
    // ------------------------------ dQ write-back (bulk tensor reductions) and dV ------------------------------
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = uint32_t(quarter * 32) << 16;
    const bool leader = warp == 18 && lane == 0;
    const uint32_t stage_row = smem_u32(sdQ) + r * 128;
    for (int i = 0; i < nq; ++i) {
      mbar_wait(dq_full, i & 1);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld32(tmdQ + lane_base, v0);
      tmem_ld32(tmdQ + lane_base + 32, v1);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_empty);
      if (!(SIMVGB_DBG(p) & 1)) {
        if (leader) bulk_wait_read0();      // the previous tile's reductions have finished READING the staging tile
        dq_sync();
        // row r of the fp32 tile: 64 columns = two 128-byte rows (one per 32-column sub-tile), 16-byte chunks XOR-swizzled with
        // the row index exactly as TMA's 128B swizzle expects (conflict-free: 8 consecutive rows cover all 32 banks)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          st_shared_v4(stage_row + ((c ^ (r & 7)) << 4), v0[4 * c], v0[4 * c + 1], v0[4 * c + 2], v0[4 * c + 3]);
          st_shared_v4(stage_row + kTileBytes + ((c ^ (r & 7)) << 4), v1[4 * c], v1[4 * c + 1], v1[4 * c + 2], v1[4 * c + 3]);
        }
        fence_proxy_async();                // generic-proxy stores -> visible to the bulk (async-proxy) reads
        dq_sync();
        if (leader) {
          const bool full = i < g.nfull;
          const bool tail = (i == g.nfull) && g.tail_rows > 0;
          const bool text = (i == g.text_tile) && g.Lt > 0;
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            const uint32_t src = smem_u32(sdQ) + sub * kTileBytes;
            const int col = h * kHeadDim + 32 * sub;
            if (full) tma_reduce_add_3d(&maps.dq_full, src, col, i * kTile, b);
            else {
              if (tail) tma_reduce_add_3d(&maps.dq_tail, src, col, i * kTile, b);
              if (text) tma_reduce_add_3d(&maps.dq_text, src + g.text_row * 128, col, 0, b);
            }
          }
          bulk_commit();
        }
      }
    }
    if (leader) bulk_wait_all0();           // shared memory must outlive the last bulk read
    mbar_wait(mma_done, 0);
    tc_fence_after();
    const int kv = kt * kTile + r;
    bf16* dst = nullptr;
    if (kv < g.Lv) dst = p.dqkv_v + ((long long)b * g.Lv + kv) * (3 * g.D) + 2 * g.D + h * kHeadDim;
    else if (kv >= g.T0 && kv < g.T0 + g.Lt) dst = p.dqkv_t + ((long long)b * g.Lt + (kv - g.T0)) * (3 * g.D) + 2 * g.D + h * kHeadDim;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      uint32_t v[32];
      tmem_ld32(tmdV + lane_base + cc * 32, v);
      tmem_wait_ld();
      if (dst != nullptr) {
        uint4* o = reinterpret_cast<uint4*>(dst + cc * 32);
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
  AttnMaps9 maps;
  if (make_attn_maps(&maps.qkv_full, &maps.qkv_tail, &maps.qkv_text, p.g, a->qkv_v, a->qkv_t, 3 * D)) return -1;
  if (make_attn_maps(&maps.do_full, &maps.do_tail, &maps.do_text, p.g, a->dout_v, a->dout_t, D)) return -1;
  {   // fp32 accumulators [B*L, D]: dims (D, tokens-per-sample, B), box 32 columns (= one 128-byte swizzle row) x tile rows
    const AttnGeom& g = p.g;
    uint64_t dims[3] = {(uint64_t)D, (uint64_t)g.Lv, (uint64_t)g.B};
    uint64_t strides[2] = {(uint64_t)D * 4, (uint64_t)D * 4 * g.Lv};
    uint32_t box[3] = {32, kTile, 1};
    if (make_tmap(&maps.dq_full, a->dq_acc_v, 4, 3, dims, strides, box, 1)) return -1;
    box[1] = g.tail_rows > 0 ? g.tail_rows : 8;
    if (make_tmap(&maps.dq_tail, a->dq_acc_v, 4, 3, dims, strides, box, 1)) return -1;
    if (g.Lt > 0) {
      dims[1] = g.Lt;
      strides[1] = (uint64_t)D * 4 * g.Lt;
      box[1] = g.Ltp;
      if (make_tmap(&maps.dq_text, a->dq_acc_t, 4, 3, dims, strides, box, 1)) return -1;
    } else {
      maps.dq_text = maps.dq_tail;
    }
  }

  if (!a->delta_ready) {
    // positions of the virtual axis that belong to neither token range are never written by the delta kernel: keep them 0
    SIMVGB_CUDA(cudaMemsetAsync(a->delta, 0, sizeof(float) * (size_t)a->B * a->H * lse_stride, s));
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

  if (ensure_dynamic_smem(reinterpret_cast<const void*>(attn_bwd_kernel), kBwdSmem)) return -2;
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
