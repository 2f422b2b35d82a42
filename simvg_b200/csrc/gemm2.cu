// 2-CTA (cta_group::2) variant of the tcgen05 GEMM in gemm.cu: a cluster of two CTAs (two SMs of one TPC) owns a 256 x 256
// output tile; each CTA stages its 128 rows of A and HALF of the B tile (so TMA writes and UMMA operand reads per SM drop
// from ~192 B/clk to ~128 B/clk, the shared-memory limit that caps the 1-CTA kernel), the leader issues M=256 MMAs.
// tcgen05 GEMM for sm_100a:  C[M,N] = epilogue(A(M,K) · B(N,K)^T), bf16 x bf16 -> fp32 (TMEM accumulators).
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer   global -> 128B-swizzled smem ring (kStages deep)
//   warp 1      MMA issuer     one elected lane issues tcgen05.mma (M=128, N=BN, K=16) x4 per k-block
//   warp 2      TMEM allocator (2 x BN fp32 columns: accumulators are double-buffered so the epilogue of
//               tile i overlaps the main loop of tile i+1)
//   warps 4-11  epilogue       tcgen05.ld -> registers -> fused bias / GELU / residual / scale -> global
// Operands may be K-major ([rows,K] row-major) or MN-major ([K,rows] row-major); the latter is what the
// weight-gradient GEMMs (dW = dY^T X) need, so no activation transposes are ever materialised.
//
// Replaces (reference = PyTorch eager -> cuBLAS): every nn.Linear on SimVG's hot path
// (torchscale q/k/v/out_proj, fc1, fc2 — built at /root/reference/simvg/models/vis_encs/beit/beit3_base.py:57-63,112-121)
// and its autograd backward.
#include "common.cuh"
#include "simvg_b200.h"

namespace simvgb {

#ifndef SIMVGB_GEMM_STAGED_EPI
#define SIMVGB_GEMM_STAGED_EPI 1
#endif
constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 128 + kEpiWarps * 32;

struct Gemm2Cfg {
  static constexpr int kStages = 6;
  static constexpr int kABytes = BM * BK * 2;        // 16 KB: this CTA's 128 rows of A
  static constexpr int kBBytes = 128 * BK * 2;       // 16 KB: this CTA's half (128 rows) of the 256-wide B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 512;              // 2 x 256 fp32 columns (double-buffered accumulators)
  static constexpr int kEpiStageBytes = 32 * 32 * 4; // per epilogue warp: one 32 x 32 fp32 chunk, transposed through smem so that
                                                     // global accesses cover whole 64/128-byte row segments
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiWarps * kEpiStageBytes + 1024 + 256;
};

struct GemmParams {
  int M, N, K;
  int a_mn, b_mn;
  int m_tiles, n_tiles, k_splits, kb_per_split, kb_total;
  int epilogue;
  const float* bias;
  bf16* out_bf16;
  bf16* out2_bf16;
  float* out_f32;
  const float* res_f32;
  long long ldo;
  float scale;
  int scale_cols;
  const float* row_scale;
  int rows_per_scale;
  int accumulate;
};

__device__ __forceinline__ float gelu_erf(float x) { return gelu_fwd(x); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
             const __grid_constant__ GemmParams p0, const __grid_constant__ CUtensorMap tmA1,
             const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ GemmParams p1, const int tiles1) {
  // Two independent problems may share one persistent launch (SimVG's multiway experts: the vision-token GEMM and the
  // 80x smaller text-token GEMM of the same layer).  Tiles of problem 0 come first; the tiles of problem 1 fill the tail.
  using Cfg = Gemm2Cfg;
  constexpr int BN = 256;
  const uint32_t cta = cluster_ctarank();   // 0 = leader (issues the MMAs), 1 = peer
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* smemA = smem;
  uint8_t* smemB = smem + Cfg::kStages * Cfg::kABytes;
  uint8_t* smemEpi = smem + Cfg::kStages * Cfg::kStageBytes;   // [kEpiWarps][4 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smemEpi + kEpiWarps * Cfg::kEpiStageBytes);
  uint64_t* full = bars;                       // [kStages]
  uint64_t* empty = bars + Cfg::kStages;       // [kStages]
  uint64_t* acc_full = bars + 2 * Cfg::kStages;   // [2]
  uint64_t* acc_empty = acc_full + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (tiles1 > 0) {
      tma_prefetch_desc(&tmA1);
      tma_prefetch_desc(&tmB1);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full[s], 2);    // leader: expect_tx arrive + the peer's remote arrive
      mbar_init(&empty[s], 1);   // multicast tcgen05.commit from the leader
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 2 * kEpiWarps);   // epilogue warps of both CTAs arrive on the leader's barrier
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2cta(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();   // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // shfl from a fixed lane => provably warp-uniform

  const int tiles0 = p0.m_tiles * p0.n_tiles * p0.k_splits;   // m_tiles counts 256-row tiles
  const int total_tiles = tiles0 + tiles1;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tt = cluster_id; tt < total_tiles; tt += num_clusters) {
        const bool second = tt >= tiles0;
        const GemmParams& p = second ? p1 : p0;
        const CUtensorMap& tmA = second ? tmA1 : tmA0;
        const CUtensorMap& tmB = second ? tmB1 : tmB0;
        const int t = second ? tt - tiles0 : tt;
        const int split = t % p.k_splits;
        const int mn = t / p.k_splits;
        const int m0 = (mn / p.n_tiles) * 256 + cta * BM;       // this CTA's 128 rows of the 256-row tile
        const int n0 = (mn % p.n_tiles) * BN + cta * (BN / 2);  // this CTA's half of the B tile
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (cta == 0) mbar_expect_tx(&full[stage], 2 * Cfg::kStageBytes);   // both CTAs' bytes land on the leader's barrier
          uint8_t* a = smemA + stage * Cfg::kABytes;
          uint8_t* b = smemB + stage * Cfg::kBBytes;
          const int k0 = kb * BK;
          if (!p.a_mn) {
            tma_load_2d_2cta(a, &tmA, &full[stage], k0, m0);  // box (64 k, 128 m)
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i)            // boxes (64 m, 64 k)
              tma_load_2d_2cta(a + i * 8192, &tmA, &full[stage], m0 + i * 64, k0);
          }
          if (!p.b_mn) {
            tma_load_2d_2cta(b, &tmB, &full[stage], k0, n0);  // box (64 k, 128 n)
          } else {
#pragma unroll
            for (int i = 0; i < BN / 128; ++i)
              tma_load_2d_2cta(b + i * 8192, &tmB, &full[stage], n0 + i * 64, k0);
          }
          if (cta != 0) mbar_arrive_remote(&full[stage], 0);   // "my loads are in flight" -> leader's full barrier
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && cta == 0) {
    // ===================== MMA issuer =====================
    // The whole warp runs this loop with warp-uniform values (descriptors stay in uniform registers); only the
    // tcgen05.mma / tcgen05.commit instructions themselves are issued by one lane.
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tt = cluster_id; tt < total_tiles; tt += num_clusters) {
      const bool second = tt >= tiles0;
      const GemmParams& p = second ? p1 : p0;
      const int t = second ? tt - tiles0 : tt;
      const uint32_t idesc = umma_idesc_bf16(256, BN, p.a_mn, p.b_mn);   // M = 256 across the CTA pair
      // K-major: 8-row groups 1024 B apart; UMMA_K=16 -> +32 B.  MN-major: next 64-wide atom 8192 B
      // away (LBO), 8 k-row groups 1024 B apart (SBO); UMMA_K=16 -> +16 rows = 2048 B.
      const uint32_t a_lbo = p.a_mn ? 8192u : 16u, b_lbo = p.b_mn ? 8192u : 16u;
      const uint32_t a_kstep = p.a_mn ? (2048u >> 4) : (32u >> 4), b_kstep = p.b_mn ? (2048u >> 4) : (32u >> 4);
      const int split = t % p.k_splits;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t adesc = umma_smem_desc(smem_u32(smemA + stage * Cfg::kABytes), a_lbo, 1024);
        const uint64_t bdesc = umma_smem_desc(smem_u32(smemB + stage * Cfg::kBBytes), b_lbo, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_f16_ss_2cta(d_tmem, adesc + k * a_kstep, bdesc + k * b_kstep, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit_2cta(&empty[stage]);  // frees the smem slot in both CTAs once these MMAs retire
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit_2cta(&acc_full[acc]);   // accumulator ready for both CTAs' epilogues
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp - 4;
    const int quarter = warp & 3;         // TMEM lane quarter this warp may access
    const int half = ew >> 2;             // which half of the BN columns
    constexpr int kChunksPerHalf = BN / 64;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tt = cluster_id; tt < total_tiles; tt += num_clusters) {
      const bool second = tt >= tiles0;
      const GemmParams& p = second ? p1 : p0;
      const int t = second ? tt - tiles0 : tt;
      const int mn = t / p.k_splits;
      const int m0 = (mn / p.n_tiles) * 256 + cta * BM;
      const int n0 = (mn % p.n_tiles) * BN;
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      float rs = 1.0f;
      if (p.epilogue == SIMVGB_EPI_RESID && p.row_scale != nullptr && row_ok)
        rs = __ldg(p.row_scale + row / p.rows_per_scale);
#pragma unroll 1
      for (int cc = 0; cc < kChunksPerHalf; ++cc) {
        const int c = half * kChunksPerHalf + cc;
        const int col0 = n0 + c * 32;
        uint32_t v[32];
        tmem_ld32(tmem_base + (uint32_t(quarter * 32) << 16) + acc * BN + c * 32, v);
        tmem_wait_ld();
        const bool chunk_rows_ok = (m0 + quarter * 32 + 32) <= p.M;   // warp-uniform: all 32 rows of this warp exist
        if (col0 >= p.N || !row_ok) continue;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        const bool full_chunk = (col0 + 32 <= p.N) && ((p.ldo & 7) == 0);
        if (p.bias != nullptr && p.epilogue != SIMVGB_EPI_ATOMIC) {
          if (col0 + 32 <= p.N) {   // 8 vector loads (all lanes read the same addresses: one broadcast transaction each)
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bv = __ldg(b4 + j);
              f[4 * j] += bv.x; f[4 * j + 1] += bv.y; f[4 * j + 2] += bv.z; f[4 * j + 3] += bv.w;
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) f[j] += __ldg(p.bias + col0 + j);
          }
        }
        const long long off = (long long)row * p.ldo + col0;
#if SIMVGB_GEMM_STAGED_EPI
        // ---- coalesced path (full 32-column chunks): with one output row per lane every 16-byte access of a warp lands in
        // a different 128-byte line (32 line requests per instruction — the residual epilogue of the K=768 GEMMs was LSU-
        // bound at 2.7x the tile's MMA time).  The chunk is transposed through a private 4 KB smem slab (XOR-swizzled,
        // conflict-free both ways) so that 8 (fp32) / 4 (bf16) lanes cover one contiguous row segment.
        if (full_chunk && chunk_rows_ok && (p.epilogue == SIMVGB_EPI_RESID || p.epilogue == SIMVGB_EPI_BF16 ||
                                            (p.epilogue == SIMVGB_EPI_F32 && !p.accumulate))) {
          const uint32_t slab = smem_u32(smemEpi + ew * Cfg::kEpiStageBytes);
          const long long row0 = m0 + quarter * 32;
          if (p.epilogue == SIMVGB_EPI_BF16) {
            if (col0 + 32 <= p.scale_cols) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] *= p.scale;
            } else if (col0 < p.scale_cols) {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.scale_cols) f[j] *= p.scale;
            }
            // rows of 64 B; 16-byte slot s of row r lives at r*64 + ((s ^ ((r >> 1) & 3)) << 4)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              st_shared_v4(slab + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4), pack_bf16x2(f[8 * j], f[8 * j + 1]),
                           pack_bf16x2(f[8 * j + 2], f[8 * j + 3]), pack_bf16x2(f[8 * j + 4], f[8 * j + 5]),
                           pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = 8 * i + (lane >> 2), sl = lane & 3;
              const uint4 v4 = ld_shared_v4(slab + r * 64 + ((sl ^ ((r >> 1) & 3)) << 4));
              *reinterpret_cast<uint4*>(p.out_bf16 + (row0 + r) * p.ldo + col0 + 8 * sl) = v4;
            }
          } else {
            if (p.epilogue == SIMVGB_EPI_RESID) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] *= rs;
            }
            // rows of 128 B; 16-byte slot s of row r lives at r*128 + ((s ^ (r & 7)) << 4)
#pragma unroll
            for (int j = 0; j < 8; ++j)
              st_shared_v4(slab + lane * 128 + ((j ^ (lane & 7)) << 4), __float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]),
                           __float_as_uint(f[4 * j + 2]), __float_as_uint(f[4 * j + 3]));
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = 4 * i + (lane >> 3), sl = lane & 7;
              const uint4 v4 = ld_shared_v4(slab + r * 128 + ((sl ^ (r & 7)) << 4));
              float4 o4 = make_float4(__uint_as_float(v4.x), __uint_as_float(v4.y), __uint_as_float(v4.z), __uint_as_float(v4.w));
              const long long o = (row0 + r) * p.ldo + col0 + 4 * sl;
              if (p.epilogue == SIMVGB_EPI_RESID) {
                const float4 rv = __ldg(reinterpret_cast<const float4*>(p.res_f32 + o));
                o4.x += rv.x; o4.y += rv.y; o4.z += rv.z; o4.w += rv.w;
              }
              *reinterpret_cast<float4*>(p.out_f32 + o) = o4;
            }
          }
          __syncwarp();   // the slab is reused by the next chunk
          continue;
        }
#endif
        switch (p.epilogue) {
          case SIMVGB_EPI_BF16: {
            if (col0 + 32 <= p.scale_cols) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] *= p.scale;
            } else if (col0 < p.scale_cols) {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.scale_cols) f[j] *= p.scale;
            }
            if (full_chunk) {
              uint4* o = reinterpret_cast<uint4*>(p.out_bf16 + off);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                o[j] = make_uint4(pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                                  pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
            } else {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) p.out_bf16[off + j] = __float2bfloat16(f[j]);
            }
          } break;
          case SIMVGB_EPI_GELU: {
            float g[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) g[j] = gelu_erf(f[j]);
            if (full_chunk) {
              uint4* o = reinterpret_cast<uint4*>(p.out_bf16 + off);
              uint4* o2 = reinterpret_cast<uint4*>(p.out2_bf16 + off);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                o[j] = make_uint4(pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                                  pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
                o2[j] = make_uint4(pack_bf16x2(g[8 * j], g[8 * j + 1]), pack_bf16x2(g[8 * j + 2], g[8 * j + 3]),
                                   pack_bf16x2(g[8 * j + 4], g[8 * j + 5]), pack_bf16x2(g[8 * j + 6], g[8 * j + 7]));
              }
            } else {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) {
                  p.out_bf16[off + j] = __float2bfloat16(f[j]);
                  p.out2_bf16[off + j] = __float2bfloat16(g[j]);
                }
            }
          } break;
          case SIMVGB_EPI_RESID: {
            if (full_chunk) {
              const float4* r = reinterpret_cast<const float4*>(p.res_f32 + off);
              float4* o = reinterpret_cast<float4*>(p.out_f32 + off);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 rv = __ldg(r + j);
                o[j] = make_float4(rv.x + rs * f[4 * j], rv.y + rs * f[4 * j + 1], rv.z + rs * f[4 * j + 2],
                                   rv.w + rs * f[4 * j + 3]);
              }
            } else {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) p.out_f32[off + j] = p.res_f32[off + j] + rs * f[j];
            }
          } break;
          case SIMVGB_EPI_F32: {
            if (full_chunk) {
              float4* o = reinterpret_cast<float4*>(p.out_f32 + off);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 ov = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                if (p.accumulate) {
                  float4 old = o[j];
                  ov.x += old.x; ov.y += old.y; ov.z += old.z; ov.w += old.w;
                }
                o[j] = ov;
              }
            } else {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) p.out_f32[off + j] = (p.accumulate ? p.out_f32[off + j] : 0.f) + f[j];
            }
          } break;
          default: {  // SIMVGB_EPI_ATOMIC
            if (full_chunk) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.out_f32 + off + 4 * j), "f"(f[4 * j]),
                             "f"(f[4 * j + 1]), "f"(f[4 * j + 2]), "f"(f[4 * j + 3])
                             : "memory");
            } else {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) atomicAdd(p.out_f32 + off + j, f[j]);
            }
          } break;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&acc_empty[acc], 0);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();   // the peer may still be reading TMEM written by the leader's MMAs / smem being multicast-signalled
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, Cfg::kTmemCols);
  }
}


// Host side.
static int prepare_2cta(const simvgb_gemm_args* a, CUtensorMap* tmA, CUtensorMap* tmB, GemmParams* pp) {
  constexpr int BN = 256;
  {
    uint64_t dims[2], strides[1];
    uint32_t box[2];
    if (!a->a_mn_major) { dims[0] = a->K; dims[1] = a->M; box[0] = BK; box[1] = BM; }
    else                { dims[0] = a->M; dims[1] = a->K; box[0] = 64; box[1] = BK; }
    strides[0] = (uint64_t)a->lda * 2;
    if (make_tmap(tmA, a->A, 2, 2, dims, strides, box, 1)) return -1;
    if (!a->b_mn_major) { dims[0] = a->K; dims[1] = a->N; box[0] = BK; box[1] = BN / 2; }
    else                { dims[0] = a->N; dims[1] = a->K; box[0] = 64; box[1] = BK; }
    strides[0] = (uint64_t)a->ldb * 2;
    if (make_tmap(tmB, a->B, 2, 2, dims, strides, box, 1)) return -1;
  }
  GemmParams& p = *pp;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.a_mn = a->a_mn_major; p.b_mn = a->b_mn_major;
  p.m_tiles = (a->M + 255) / 256;
  p.n_tiles = (a->N + BN - 1) / BN;
  p.kb_total = (a->K + BK - 1) / BK;
  int ks = a->k_splits < 1 ? 1 : a->k_splits;
  if (ks > p.kb_total) ks = p.kb_total;
  p.kb_per_split = (p.kb_total + ks - 1) / ks;
  p.k_splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.epilogue = a->epilogue;
  p.bias = a->bias;
  p.out_bf16 = reinterpret_cast<bf16*>(a->out_bf16);
  p.out2_bf16 = reinterpret_cast<bf16*>(a->out2_bf16);
  p.out_f32 = a->out_f32;
  p.res_f32 = a->res_f32;
  p.ldo = a->ldo;
  p.scale = a->scale;
  p.scale_cols = a->scale_cols;
  p.row_scale = a->row_scale;
  p.rows_per_scale = a->rows_per_scale > 0 ? a->rows_per_scale : 1;
  p.accumulate = a->accumulate;
  return 0;
}

// Launches problem `a` and, when b != nullptr, problem `b` in the same persistent grid.
int launch_gemm_2cta(const simvgb_gemm_args* a, const simvgb_gemm_args* b, cudaStream_t stream) {
  CUtensorMap tmA0, tmB0, tmA1, tmB1;
  GemmParams p0, p1;
  if (prepare_2cta(a, &tmA0, &tmB0, &p0)) return -1;
  int tiles1 = 0;
  if (b != nullptr) {
    if (prepare_2cta(b, &tmA1, &tmB1, &p1)) return -1;
    tiles1 = p1.m_tiles * p1.n_tiles * p1.k_splits;
  } else {
    tmA1 = tmA0; tmB1 = tmB0; p1 = p0;
  }
  if (ensure_dynamic_smem(reinterpret_cast<const void*>(gemm2_kernel), Gemm2Cfg::kSmemBytes)) return -2;
  const int total = p0.m_tiles * p0.n_tiles * p0.k_splits + tiles1;
  int clusters = sm_count() / 2;
  if (clusters > total) clusters = total;
  gemm2_kernel<<<2 * clusters, kThreads, Gemm2Cfg::kSmemBytes, stream>>>(tmA0, tmB0, p0, tmA1, tmB1, p1, tiles1);
  SIMVGB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace simvgb
