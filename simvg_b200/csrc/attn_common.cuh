// Shared pieces of the fused multiway attention kernels (forward + backward).
//
// Token layout.  The encoder keeps vision tokens and text tokens in two contiguous token-major buffers
// (expert A rows / expert B rows — SURVEY §7.1), so the per-token GEMMs are plain 2-D problems.  Only attention is
// joint across modalities (beit3_base.py:137-145 -> torchscale MultiheadAttention, A.4): per sample the key/query
// sequence is [Lv vision tokens | Lt text tokens].  The kernels address it through a *virtual* sequence axis:
//     vision token i  -> virtual position i            (0 <= i < Lv)
//     text token  i   -> virtual position T0 + i       (T0 = first multiple of 8 >= Lv that keeps the text block
//                                                        inside one 128-row tile)
// so that a 128-row tile is assembled from at most two TMA boxes (vision rows, text rows) that both land on
// 1024-byte (8-row) swizzle-atom boundaries.  Positions in neither range — and padded text tokens — are masked.
#pragma once
#include "common.cuh"

namespace simvgb {

constexpr int kTile = 128;     // q rows / k rows per tile
constexpr int kHeadDim = 64;   // BEiT-3 base and large both use 64
constexpr int kTileBytes = kTile * kHeadDim * 2;  // 16 KB: one [128 x 64] bf16 tile, 128-byte rows

struct AttnGeom {
  int B, H, Lv, Lt, D;
  int Ltp;        // text box rows (Lt rounded up to 8)
  int T0;         // virtual position of text token 0
  int nfull;      // number of full 128-row vision tiles
  int tail_rows;  // rows of the vision tail box (multiple of 8, 0 if Lv % 128 == 0)
  int text_tile;  // T0 / 128
  int text_row;   // T0 % 128
  int ntiles;     // tiles along the virtual axis
};

inline AttnGeom make_attn_geom(int B, int H, int Lv, int Lt, int D) {
  AttnGeom g;
  g.B = B; g.H = H; g.Lv = Lv; g.Lt = Lt; g.D = D;
  g.Ltp = (Lt + 7) & ~7;
  const int Lv8 = (Lv + 7) & ~7;
  g.nfull = Lv / kTile;
  g.tail_rows = Lv8 - g.nfull * kTile;
  g.T0 = ((Lv8 % kTile) + g.Ltp <= kTile) ? Lv8 : ((Lv8 + kTile - 1) / kTile) * kTile;
  if (Lt == 0) g.T0 = Lv8;
  g.text_tile = g.T0 / kTile;
  g.text_row = g.T0 % kTile;
  g.ntiles = (g.T0 + (Lt > 0 ? Lt : 0) + kTile - 1) / kTile;
  if (g.ntiles * kTile < Lv) g.ntiles = (Lv + kTile - 1) / kTile;
  return g;
}

#ifdef __CUDACC__
// Loads virtual tile `t` (columns [col, col+64) of the token-major q/k/v buffers) into a 16 KB smem tile.
__device__ __forceinline__ void load_virtual_tile(uint8_t* dst, uint64_t* bar, const AttnGeom& g,
                                                  const CUtensorMap* map_full, const CUtensorMap* map_tail,
                                                  const CUtensorMap* map_text, int t, int col, int b) {
  uint32_t bytes = 0;
  const bool full = t < g.nfull;
  const bool tail = (t == g.nfull) && g.tail_rows > 0;
  const bool text = (t == g.text_tile) && g.Lt > 0;
  if (full) bytes = kTileBytes;
  else bytes = (tail ? g.tail_rows * 128 : 0) + (text ? g.Ltp * 128 : 0);
  mbar_expect_tx(bar, bytes);
  if (full) {
    tma_load_3d(dst, map_full, bar, col, t * kTile, b);
  } else {
    if (tail) tma_load_3d(dst, map_tail, bar, col, t * kTile, b);
    if (text) tma_load_3d(dst + g.text_row * 128, map_text, bar, col, 0, b);
  }
}

// 128-bit validity mask of virtual tile t (bit j = position t*128+j is a real, unpadded token).  Called by 128
// consecutive threads (4 full warps): thread j tests position j, one ballot per warp writes one mask word.
__device__ __forceinline__ void build_tile_mask(uint32_t* mask4, const AttnGeom& g, const unsigned char* pad,
                                                int b, int t, int j /*0..127*/) {
  const int v = t * kTile + j;
  bool ok = v < g.Lv;
  if (!ok && v >= g.T0 && v < g.T0 + g.Lt) ok = (pad == nullptr) || (pad[b * g.Lt + (v - g.T0)] == 0);
  const uint32_t bits = __ballot_sync(0xffffffffu, ok);
  if ((j & 31) == 0) mask4[j >> 5] = bits;
}

// Byte offset of (row r, 16-byte chunk c) inside a [128 x 128] bf16 tile stored as two [128 x 64] 128B-swizzled
// sub-tiles (the layout both the K-major and the MN-major UMMA descriptors read).
__device__ __forceinline__ uint32_t swz_off(int r, int chunk16 /*0..15*/) {
  const int sub = chunk16 >> 3, c = chunk16 & 7;
  return sub * kTileBytes + r * 128 + ((c ^ (r & 7)) << 4);
}
#endif

}  // namespace simvgb
