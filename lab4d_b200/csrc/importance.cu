// Eval-mode importance sampling after the coarse pass (NeRF.importance_sampling, lab4d/nnutils/nerf.py:686-738;
// sample_pdf with det=True, lab4d/utils/render_utils.py:187-233): per ray, the piecewise-constant pdf
// weights[1:-1] + 1e-5 over the Dc-1 mid-points of the coarse depths is inverted at u_j = j / (Dc - 1), and the Dc new
// depths are merged with the coarse ones.  One warp per ray, everything in shared memory; a few KB of traffic per ray.
#include <cuda_runtime.h>

#include "kernels.h"

namespace b200r {

constexpr int kImpWarps = 4;

__global__ void __launch_bounds__(kImpWarps * 32) importance_fwd_kernel(const b200r_importance_args a) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Dc = a.Dc, n = Dc - 2;                 // n bins between the Dc - 1 mid-points
  const int r = blockIdx.x * kImpWarps + warp;
  float* base = sm + (size_t)warp * (4 * Dc + 4);
  float* dc = base;                                // [Dc]   coarse depths
  float* mid = dc + Dc;                            // [Dc-1] bin edges
  float* cdf = mid + Dc;                           // [n+1]
  float* fine = cdf + Dc;                          // [Dc]   new samples
  if (r >= a.R) return;
  const float* dsrc = a.depth_c + (size_t)r * Dc;
  const float* wsrc = a.weights + (size_t)r * Dc;
  for (int i = lane; i < Dc; i += 32) dc[i] = dsrc[i];
  __syncwarp();
  for (int i = lane; i < Dc - 1; i += 32) mid[i] = 0.5f * (dc[i] + dc[i + 1]);
  // pdf = (w + eps) / sum; cdf = [0, cumsum(pdf)] -- sequential sums in the order torch.cumsum uses on one row
  if (lane == 0) {
    float tot = 0.f;
    for (int i = 0; i < n; ++i) tot += wsrc[1 + i] + 1e-5f;
    float c = 0.f;
    cdf[0] = 0.f;
    for (int i = 0; i < n; ++i) {
      c += (wsrc[1 + i] + 1e-5f) / tot;
      cdf[i + 1] = c;
    }
  }
  __syncwarp();
  for (int j = lane; j < Dc; j += 32) {
    const float stepu = 1.0f / (float)(Dc - 1);      // torch.linspace(0, 1, Dc): symmetric evaluation from both ends
    const float u = j < Dc / 2 ? stepu * (float)j : 1.0f - stepu * (float)(Dc - 1 - j);
    int lo = 0, hi = n + 1;                          // searchsorted(cdf, u, right=True)
    while (lo < hi) { const int m = (lo + hi) >> 1; if (cdf[m] <= u) lo = m + 1; else hi = m; }
    const int below = lo - 1 < 0 ? 0 : lo - 1, above = lo > n ? n : lo;
    float denom = cdf[above] - cdf[below];
    if (denom < 1e-5f) denom = 1.f;
    fine[j] = mid[below] + (u - cdf[below]) / denom * (mid[above] - mid[below]);
  }
  __syncwarp();
  // fp32 rounding at a bin boundary (or the denom < eps branch) can put a sample 1 ulp below its predecessor; the merge
  // below ranks by binary search and needs a non-decreasing list (the reference sorts instead): running maximum
  if (lane == 0)
    for (int j = 1; j < Dc; ++j) fine[j] = fmaxf(fine[j], fine[j - 1]);
  __syncwarp();
  // merge the two ascending lists (coarse first on ties, like a stable sort of cat([coarse, fine]))
  float* out = a.depth_out + (size_t)r * 2 * Dc;
  for (int t = lane; t < 2 * Dc; t += 32) {
    int lo = 0, hi = Dc, pos;
    float key;
    if (t < Dc) {
      key = dc[t];
      while (lo < hi) { const int m = (lo + hi) >> 1; if (fine[m] < key) lo = m + 1; else hi = m; }
      pos = t + lo;
    } else {
      key = fine[t - Dc];
      while (lo < hi) { const int m = (lo + hi) >> 1; if (dc[m] <= key) lo = m + 1; else hi = m; }
      pos = (t - Dc) + lo;
    }
    out[pos] = key;
  }
}

cudaError_t launch_importance_fwd(const b200r_importance_args& a, cudaStream_t stream) {
  const size_t smem = (size_t)kImpWarps * (4 * a.Dc + 4) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(importance_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  importance_fwd_kernel<<<(a.R + kImpWarps - 1) / kImpWarps, kImpWarps * 32, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace b200r
