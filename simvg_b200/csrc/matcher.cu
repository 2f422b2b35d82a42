// Device-side Hungarian matching for the DETR-style set criterion (nq <= 32 queries, <= 32 targets per sample).
//
// Replaces detrex HungarianMatcher's `C.cpu()` + per-sample scipy.optimize.linear_sum_assignment loop
// (/root/reference/simvg/core/criterion/criterion.py:226-271 calls self.matcher(outputs, targets); matcher semantics in SURVEY
// Appendix A.12) — a blocking device->host copy and a Python loop per criterion call (main + every auxiliary decoder layer +
// teacher targets: >= 6 round trips per train step in the reference).  One thread per sample runs the O(n^2 m) shortest
// augmenting path algorithm (Kuhn-Munkres with potentials, the classic u/v/p/way formulation) in double precision on that
// sample's [nq x n_i] block of the batched cost matrix; the problems are tiny (<= 32 x 32), the point is staying on the device.
#include "common.cuh"
#include "simvg_b200.h"

namespace simvgb {

constexpr int kMaxDim = 32;

// cost: [B, nq, ttot] fp32 (sample b's targets are columns offsets[b] .. offsets[b+1]-1 of ITS rows, exactly the layout
// HungarianMatcher builds: C.view(B, nq, -1) then split by target counts).  out_q / out_t: [B, kmax] int64, assignments of
// sample b in ascending query order (scipy's row_ind / col_ind), padded with -1.
__global__ void hungarian_kernel(const float* __restrict__ cost, int B, int nq, int ttot, const int* __restrict__ offsets,
                                 long long* __restrict__ out_q, long long* __restrict__ out_t, int kmax) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int t0 = offsets[b], nt = offsets[b + 1] - t0;
  for (int k = 0; k < kmax; ++k) { out_q[(long long)b * kmax + k] = -1; out_t[(long long)b * kmax + k] = -1; }
  if (nt <= 0 || nq <= 0) return;
  const float* c = cost + ((long long)b * nq) * ttot + t0;   // c[q * ttot + t]
  // rows of the assignment problem = the smaller side (n <= m), as scipy does by transposing
  const bool transposed = nq > nt;
  const int n = transposed ? nt : nq, m = transposed ? nq : nt;
  auto a = [&](int i, int j) -> double {   // 1-based (row i, column j) of the n x m problem
    return transposed ? (double)c[(j - 1) * ttot + (i - 1)] : (double)c[(i - 1) * ttot + (j - 1)];
  };
  double u[kMaxDim + 1], v[kMaxDim + 1], minv[kMaxDim + 1];
  int p[kMaxDim + 1], way[kMaxDim + 1];
  bool used[kMaxDim + 1];
  for (int j = 0; j <= m; ++j) { v[j] = 0.0; p[j] = 0; way[j] = 0; }
  for (int i = 0; i <= n; ++i) u[i] = 0.0;
  for (int i = 1; i <= n; ++i) {
    p[0] = i;
    int j0 = 0;
    for (int j = 0; j <= m; ++j) { minv[j] = 1e300; used[j] = false; }
    do {
      used[j0] = true;
      const int i0 = p[j0];
      double delta = 1e300;
      int j1 = 0;
      for (int j = 1; j <= m; ++j) {
        if (used[j]) continue;
        const double cur = a(i0, j) - u[i0] - v[j];
        if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
        if (minv[j] < delta) { delta = minv[j]; j1 = j; }
      }
      for (int j = 0; j <= m; ++j) {
        if (used[j]) { u[p[j]] += delta; v[j] -= delta; }
        else minv[j] -= delta;
      }
      j0 = j1;
    } while (p[j0] != 0);
    do {
      const int j1 = way[j0];
      p[j0] = p[j1];
      j0 = j1;
    } while (j0 != 0);
  }
  // p[j] = row assigned to column j.  Emit (query, target) pairs in ascending query order.
  int qt[kMaxDim];   // target assigned to query q, or -1
  for (int q = 0; q < nq; ++q) qt[q] = -1;
  for (int j = 1; j <= m; ++j) {
    if (p[j] == 0) continue;
    if (transposed) qt[j - 1] = p[j] - 1;   // column = query, row = target
    else qt[p[j] - 1] = j - 1;              // row = query, column = target
  }
  int k = 0;
  for (int q = 0; q < nq && k < kmax; ++q) {
    if (qt[q] < 0) continue;
    out_q[(long long)b * kmax + k] = q;
    out_t[(long long)b * kmax + k] = qt[q];
    ++k;
  }
}

}  // namespace simvgb

extern "C" int simvgb_hungarian(const float* cost, int B, int nq, int ttot, const int32_t* offsets, int64_t* out_q,
                                int64_t* out_t, int kmax, void* stream) {
  using namespace simvgb;
  SIMVGB_CHECK(cost && offsets && out_q && out_t, "simvgb_hungarian: null pointer");
  SIMVGB_CHECK(B > 0 && nq > 0 && nq <= kMaxDim && ttot >= 0 && kmax > 0, "simvgb_hungarian: bad shape (B=%d nq=%d ttot=%d kmax=%d; nq <= %d)",
               B, nq, ttot, kmax, kMaxDim);
  hungarian_kernel<<<(B + 63) / 64, 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      cost, B, nq, ttot, offsets, reinterpret_cast<long long*>(out_q), reinterpret_cast<long long*>(out_t), kmax);
  SIMVGB_CUDA(cudaGetLastError());
  return 0;
}
