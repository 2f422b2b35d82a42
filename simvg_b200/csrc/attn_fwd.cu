// Fused multiway self-attention forward for sm_100a:  O = softmax(Q K^T + key_padding_mask) V  per (sample, head).
//
// Replaces torchscale MultiheadAttention's bmm -> masked_fill(-inf) -> softmax(fp32) -> bmm sequence
// (called at /root/reference/simvg/models/vis_encs/beit/beit3_base.py:137-145; semantics SURVEY Appendix A.4) without
// ever materialising the [B*H, L, L] score tensor.  q arrives pre-scaled by head_dim^-0.5 (fused into the QKV GEMM
// epilogue, matching `q *= self.scaling` after the bias add).
//
// One CTA = one 128-query tile of one (b, h); two CTAs are co-resident per SM so one CTA's softmax overlaps the other's
// MMAs.  Warp roles:  0 = TMA producer, 1 = tcgen05.mma issuer, 2-5 = softmax (one thread per query row).
//   S = Q K_j^T   : tcgen05.mma M=128 N=128 K=64, both operands K-major, S in TMEM columns [0,128)
//   O += P V_j    : P (bf16) written by the softmax threads into 128B-swizzled smem, V_j read MN-major straight
//                   from its token-major TMA tile; O in TMEM columns [128,192)
// Online softmax with lazy rescaling (O is only rescaled when a row max grows by more than 2^8).
#include "attn_common.cuh"
#include "simvg_b200.h"

namespace simvgb {

constexpr int kFwdThreads = 192;
constexpr int kSlots = 3;  // K/V ring: K_j, V_j, K_{j+1} ...
constexpr int kFwdSmem = kTileBytes /*Q*/ + 2 * kTileBytes /*P*/ + kSlots * kTileBytes + 1024 + 256;
constexpr float kLog2e = 1.4426950408889634f;

struct AttnFwdParams {
  AttnGeom g;
  const unsigned char* pad;  // [B, Lt] 1 = padded text token, or null
  bf16* out_v;               // [B*Lv, D]
  bf16* out_t;               // [B*Lt, D]
  float* lse;                // [B, H, ntiles*128]  log2-domain logsumexp of each query row
};

__global__ void __launch_bounds__(kFwdThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap map_full, const __grid_constant__ CUtensorMap map_tail,
                const __grid_constant__ CUtensorMap map_text, const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sP = smem + kTileBytes;
  uint8_t* sKV = smem + 3 * kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (3 + kSlots) * kTileBytes);
  uint64_t* q_full = bars;
  uint64_t* slot_full = bars + 1;             // [kSlots]
  uint64_t* slot_empty = bars + 1 + kSlots;   // [kSlots]
  uint64_t* s_full = bars + 1 + 2 * kSlots;
  uint64_t* s_empty = s_full + 1;
  uint64_t* p_full = s_full + 2;
  uint64_t* pv_done = s_full + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 4);
  uint32_t* masks = tmem_slot + 2;            // [2][4] validity bits of the (at most two) partial tiles

  const AttnGeom& g = p.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nk = g.ntiles;

  // Zero the staging tiles: partial (tail/text) tiles only overwrite some rows, the rest must stay finite.
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < (3 + kSlots) * kTileBytes / 16; i += kFwdThreads) z[i] = zero;
  }
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < kSlots; ++s) { mbar_init(&slot_full[s], 1); mbar_init(&slot_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 128);
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 2) {
    const int t = g.nfull + (threadIdx.x - 64);
    if (t < nk) build_tile_mask(masks + 4 * (threadIdx.x - 64), g, p.pad, b, t);
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  fence_proxy_async();  // generic-proxy zero fill -> visible before TMA (async proxy) writes
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmS = tmem, tmO = tmem + 128;

  if (warp == 0) {
    if (lane == 0) {
      const int colq = h * kHeadDim, colk = g.D + h * kHeadDim, colv = 2 * g.D + h * kHeadDim;
      load_virtual_tile(sQ, q_full, g, &map_full, &map_tail, &map_text, qt, colq, b);
      for (int n = 0; n < 2 * nk; ++n) {
        const int slot = n % kSlots;
        const uint32_t ph = (n / kSlots) & 1;
        mbar_wait(&slot_empty[slot], ph ^ 1);
        load_virtual_tile(sKV + slot * kTileBytes, &slot_full[slot], g, &map_full, &map_tail, &map_text, n >> 1,
                          (n & 1) ? colv : colk, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(128, kHeadDim, 0, 1);  // A = P (K-major), B = V (MN-major)
      const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
      auto issue_s = [&](int j) {
        const int n = 2 * j, slot = n % kSlots;
        mbar_wait(&slot_full[slot], (n / kSlots) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sKV + slot * kTileBytes);
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k)
          umma_f16_ss(tmS, umma_smem_desc(q_addr + k * 32, 16, 1024), umma_smem_desc(k_addr + k * 32, 16, 1024),
                      idesc_s, k > 0);
        umma_commit(&slot_empty[slot]);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < nk; ++j) {
        mbar_wait(s_empty, j & 1);  // softmax has consumed S_j
        if (j + 1 < nk) issue_s(j + 1);
        const int n = 2 * j + 1, slot = n % kSlots;
        mbar_wait(p_full, j & 1);
        mbar_wait(&slot_full[slot], (n / kSlots) & 1);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(sKV + slot * kTileBytes);
#pragma unroll
        for (int k = 0; k < kTile / 16; ++k) {
          const uint64_t adesc = umma_smem_desc(p_addr + (k >> 2) * kTileBytes + (k & 3) * 32, 16, 1024);
          const uint64_t bdesc = umma_smem_desc(v_addr + k * 2048, 8192, 1024);
          umma_f16_ss(tmO, adesc, bdesc, idesc_o, (j > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&slot_empty[slot]);
        umma_commit(pv_done);
      }
    }
  } else {
    // ------------------------------ softmax: one thread per query row ------------------------------
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_base = uint32_t(quarter * 32) << 16;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < nk; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const bool partial = j >= g.nfull;
      const uint32_t* mk = masks + 4 * (j - g.nfull);
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tmS + lane_base + c * 32, v);
        tmem_wait_ld();
        const uint32_t bits = partial ? mk[c] : 0xffffffffu;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float s = (bits >> i) & 1u ? __uint_as_float(v[i]) : -INFINITY;
          mx = fmaxf(mx, s);
        }
      }
      mx *= kLog2e;
      const bool need = mx > m + 8.0f;       // lazy rescale threshold (log2 units); true on the first tile
      const float m_use = need ? mx : m;
      const float alpha = need ? exp2f(m - m_use) : 1.0f;
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);     // P buffer free, O up to tile j-1 complete
        if (__any_sync(0xffffffffu, need)) {
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld32(tmO + lane_base + c * 32, v);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st32(tmO + lane_base + c * 32, v);
          }
          tmem_wait_st();
        }
      }
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tmS + lane_base + c * 32, v);
        tmem_wait_ld();
        const uint32_t bits = partial ? mk[c] : 0xffffffffu;
        float pr[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float s = (bits >> i) & 1u ? __uint_as_float(v[i]) : -INFINITY;
          pr[i] = exp2f(fmaf(s, kLog2e, -m_use));
          sum += pr[i];
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const uint4 w = make_uint4(pack_bf16x2(pr[8 * q4], pr[8 * q4 + 1]), pack_bf16x2(pr[8 * q4 + 2], pr[8 * q4 + 3]),
                                     pack_bf16x2(pr[8 * q4 + 4], pr[8 * q4 + 5]), pack_bf16x2(pr[8 * q4 + 6], pr[8 * q4 + 7]));
          *reinterpret_cast<uint4*>(sP + swz_off(r, c * 4 + q4)) = w;
        }
      }
      l = l * alpha + sum;
      m = m_use;
      tc_fence_before();
      mbar_arrive(s_empty);
      fence_proxy_async();   // P stores (generic proxy) -> visible to tcgen05.mma (async proxy)
      mbar_arrive(p_full);
    }
    // ------------------------------ epilogue ------------------------------
    mbar_wait(pv_done, (nk - 1) & 1);
    tc_fence_after();
    const int qv = qt * kTile + r;
    bf16* dst = nullptr;
    if (qv < g.Lv) dst = p.out_v + ((long long)b * g.Lv + qv) * g.D + h * kHeadDim;
    else if (qv >= g.T0 && qv < g.T0 + g.Lt) dst = p.out_t + ((long long)b * g.Lt + (qv - g.T0)) * g.D + h * kHeadDim;
    const float inv = 1.0f / l;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld32(tmO + lane_base + c * 32, v);
      tmem_wait_ld();
      if (dst != nullptr) {
        uint4* o = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          o[q4] = make_uint4(pack_bf16x2(__uint_as_float(v[8 * q4]) * inv, __uint_as_float(v[8 * q4 + 1]) * inv),
                             pack_bf16x2(__uint_as_float(v[8 * q4 + 2]) * inv, __uint_as_float(v[8 * q4 + 3]) * inv),
                             pack_bf16x2(__uint_as_float(v[8 * q4 + 4]) * inv, __uint_as_float(v[8 * q4 + 5]) * inv),
                             pack_bf16x2(__uint_as_float(v[8 * q4 + 6]) * inv, __uint_as_float(v[8 * q4 + 7]) * inv));
      }
    }
    if (p.lse != nullptr) p.lse[((long long)b * g.H + h) * (g.ntiles * kTile) + qv] = m + log2f(l);
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
  CUtensorMap full, tail, text;
  if (make_attn_maps(&full, &tail, &text, p.g, a->qkv_v, a->qkv_t, 3 * D)) return -1;
  static bool attr_set = false;
  if (!attr_set) {
    SIMVGB_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem));
    attr_set = true;
  }
  dim3 grid(p.g.ntiles, a->H, a->B);
  attn_fwd_kernel<<<grid, kFwdThreads, kFwdSmem, reinterpret_cast<cudaStream_t>(stream)>>>(full, tail, text, p);
  SIMVGB_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int simvgb_attn_lse_stride(int Lv, int Lt) {
  simvgb::AttnGeom g = simvgb::make_attn_geom(1, 1, Lv, Lt, 64);
  return g.ntiles * simvgb::kTile;
}
