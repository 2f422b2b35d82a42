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
#include <stdlib.h>

#include "common.cuh"
#include "simvg_b200.h"

namespace simvgb {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int kEpiWarps = 8;
constexpr int kThreads = 128 + kEpiWarps * 32;

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;  // 512 or 256
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
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

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* smemA = smem;
  uint8_t* smemB = smem + Cfg::kStages * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full = bars;                       // [kStages]
  uint64_t* empty = bars + Cfg::kStages;       // [kStages]
  uint64_t* acc_full = bars + 2 * Cfg::kStages;   // [2]
  uint64_t* acc_empty = acc_full + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // shfl from a fixed lane => provably warp-uniform

  const int total_tiles = p.m_tiles * p.n_tiles * p.k_splits;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int split = t % p.k_splits;
        const int mn = t / p.k_splits;
        const int m0 = (mn / p.n_tiles) * BM;
        const int n0 = (mn % p.n_tiles) * BN;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], Cfg::kStageBytes);
          uint8_t* a = smemA + stage * Cfg::kABytes;
          uint8_t* b = smemB + stage * Cfg::kBBytes;
          const int k0 = kb * BK;
          if (!p.a_mn) {
            tma_load_2d(a, &tmA, &full[stage], k0, m0);  // box (64 k, 128 m)
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i)            // boxes (64 m, 64 k)
              tma_load_2d(a + i * 8192, &tmA, &full[stage], m0 + i * 64, k0);
          }
          if (!p.b_mn) {
            tma_load_2d(b, &tmB, &full[stage], k0, n0);  // box (64 k, BN n)
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              tma_load_2d(b + i * 8192, &tmB, &full[stage], n0 + i * 64, k0);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs this loop with warp-uniform values (descriptors stay in uniform registers); only the
    // tcgen05.mma / tcgen05.commit instructions themselves are issued by one lane.
    const uint32_t idesc = umma_idesc_bf16(BM, BN, p.a_mn, p.b_mn);
    // K-major: 8-row groups 1024 B apart; UMMA_K=16 -> +32 B.  MN-major: next 64-wide atom 8192 B
    // away (LBO), 8 k-row groups 1024 B apart (SBO); UMMA_K=16 -> +16 rows = 2048 B.
    const uint32_t a_lbo = p.a_mn ? 8192u : 16u, b_lbo = p.b_mn ? 8192u : 16u;
    const uint32_t a_kstep = p.a_mn ? (2048u >> 4) : (32u >> 4), b_kstep = p.b_mn ? (2048u >> 4) : (32u >> 4);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
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
            umma_f16_ss(d_tmem, adesc + k * a_kstep, bdesc + k * b_kstep, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit(&empty[stage]);  // frees the smem slot once these MMAs retire
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&acc_full[acc]);   // accumulator ready for the epilogue
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
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int mn = t / p.k_splits;
      const int m0 = (mn / p.n_tiles) * BM;
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
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN>
static int launch_gemm(const simvgb_gemm_args* a, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap tmA, tmB;
  {
    // A: K-major -> dims (K, M), box (64, 128).  MN-major -> dims (M, K), box (64, 64).
    uint64_t dims[2], strides[1];
    uint32_t box[2];
    if (!a->a_mn_major) { dims[0] = a->K; dims[1] = a->M; box[0] = BK; box[1] = BM; }
    else                { dims[0] = a->M; dims[1] = a->K; box[0] = 64; box[1] = BK; }
    strides[0] = (uint64_t)a->lda * 2;
    if (make_tmap(&tmA, a->A, 2, 2, dims, strides, box, 1)) return -1;
    if (!a->b_mn_major) { dims[0] = a->K; dims[1] = a->N; box[0] = BK; box[1] = BN; }
    else                { dims[0] = a->N; dims[1] = a->K; box[0] = 64; box[1] = BK; }
    strides[0] = (uint64_t)a->ldb * 2;
    if (make_tmap(&tmB, a->B, 2, 2, dims, strides, box, 1)) return -1;
  }
  GemmParams p;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.a_mn = a->a_mn_major; p.b_mn = a->b_mn_major;
  p.m_tiles = (a->M + BM - 1) / BM;
  p.n_tiles = (a->N + BN - 1) / BN;
  p.kb_total = (a->K + BK - 1) / BK;
  int ks = a->k_splits < 1 ? 1 : a->k_splits;
  if (ks > p.kb_total) ks = p.kb_total;
  p.kb_per_split = (p.kb_total + ks - 1) / ks;
  p.k_splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;  // no empty splits
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

  if (ensure_dynamic_smem(reinterpret_cast<const void*>(gemm_kernel<BN>), Cfg::kSmemBytes)) return -2;
  const int total = p.m_tiles * p.n_tiles * p.k_splits;
  int grid = sm_count();
  if (grid > total) grid = total;
  gemm_kernel<BN><<<grid, kThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, p);
  SIMVGB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace simvgb

namespace simvgb {
int launch_gemm_2cta(const simvgb_gemm_args* a, const simvgb_gemm_args* b, cudaStream_t stream);
}

static int simvgb_gemm_validate(const simvgb_gemm_args* a) {
  using namespace simvgb;
  SIMVGB_CHECK(a != nullptr, "simvgb_gemm: null args");
  SIMVGB_CHECK(a->M > 0 && a->N > 0 && a->K > 0, "simvgb_gemm: bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
  SIMVGB_CHECK(a->A && a->B, "simvgb_gemm: null operand");
  SIMVGB_CHECK((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->B) & 15) == 0,
               "simvgb_gemm: operands must be 16-byte aligned");
  SIMVGB_CHECK(a->bias == nullptr || (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0, "simvgb_gemm: bias must be 16-byte aligned");
  SIMVGB_CHECK((a->lda % 8) == 0 && (a->ldb % 8) == 0, "simvgb_gemm: lda/ldb must be multiples of 8 (TMA 16-byte strides)");
  SIMVGB_CHECK(a->epilogue >= 0 && a->epilogue <= SIMVGB_EPI_ATOMIC, "simvgb_gemm: bad epilogue %d", a->epilogue);
  SIMVGB_CHECK(a->k_splits <= 1 || a->epilogue == SIMVGB_EPI_ATOMIC, "simvgb_gemm: k_splits > 1 needs the atomic epilogue");
  switch (a->epilogue) {
    case SIMVGB_EPI_BF16: SIMVGB_CHECK(a->out_bf16, "simvgb_gemm: out_bf16 is null"); break;
    case SIMVGB_EPI_GELU: SIMVGB_CHECK(a->out_bf16 && a->out2_bf16, "simvgb_gemm: GELU needs out_bf16 and out2_bf16"); break;
    case SIMVGB_EPI_RESID: SIMVGB_CHECK(a->out_f32 && a->res_f32, "simvgb_gemm: RESID needs out_f32 and res_f32"); break;
    default: SIMVGB_CHECK(a->out_f32, "simvgb_gemm: out_f32 is null"); break;
  }
  return 0;
}

static bool simvgb_use_2cta(const simvgb_gemm_args* a) {
  static const int use_2cta = [] { const char* e = getenv("SIMVGB_GEMM_2CTA"); return e ? atoi(e) : 1; }();
  return use_2cta && a->N > 128 && a->M >= 256;
}

extern "C" int simvgb_gemm(const simvgb_gemm_args* a, void* stream) {
  using namespace simvgb;
  if (int rc = simvgb_gemm_validate(a)) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // Large problems: 2-CTA (cta_group::2) 256x256 tiles.  SIMVGB_GEMM_2CTA=0 forces the 1-CTA kernel (A/B testing).
  if (simvgb_use_2cta(a)) return launch_gemm_2cta(a, nullptr, s);
  // BN=256 tiles unless N is small (head projections, N<=128) where half the tile would be masked.
  if (a->N > 128) return launch_gemm<256>(a, s);
  return launch_gemm<128>(a, s);
}

// Two independent GEMMs in one persistent launch (the vision-expert and text-expert problems of one multiway layer).
// Falls back to two launches when either problem is not eligible for the 2-CTA kernel.
extern "C" int simvgb_gemm_pair(const simvgb_gemm_args* a, const simvgb_gemm_args* b, void* stream) {
  using namespace simvgb;
  if (int rc = simvgb_gemm_validate(a)) return rc;
  if (int rc = simvgb_gemm_validate(b)) return rc;
  static const int pair_ok = [] { const char* e = getenv("SIMVGB_GEMM_PAIR"); return e ? atoi(e) : 1; }();
  if (pair_ok && simvgb_use_2cta(a) && simvgb_use_2cta(b))
    return launch_gemm_2cta(a, b, reinterpret_cast<cudaStream_t>(stream));
  if (int rc = simvgb_gemm(a, stream)) return rc;
  return simvgb_gemm(b, stream);
}
