// Per-ray feature matching of FeatureNeRF.global_match (lab4d/nnutils/feature.py:152-205) and its hand-derived backward.
//
//   score[r,k] = exp(logsigma) * <feat_px[r], feat_can[idx[k]]>,  prob = softmax_k(score),  xyz_matched[r] = sum_k prob[r,k] xyz_can[idx[k]]
// over K <= 2048 candidates drawn by the caller (torch.randperm, like the reference, so the random stream is the reference's).
// Latency-bound SIMT work (R x K x 19 MACs): forward = one warp per ray, the K candidates staged once per block in shared memory
// (row stride 17 floats: conflict-free), scores recomputed instead of kept; backward = one thread per candidate looping over a slice
// of the rays (no atomics: deterministic), partial sums per ray slice, then a reduce + scatter kernel that adds candidate k's
// gradient to row idx[k] of the dense per-sample gradients (idx has no duplicates).
#include <cuda_runtime.h>
#include <math.h>

#include "kernels.h"

namespace b200r {

constexpr int kMatchC = 16;          // feature channels
constexpr int kMatchStride = 17;     // shared-memory row stride of a candidate's features
constexpr int kMatchWarps = 8;
constexpr int kMatchSplit = 128;     // rays per backward block

__device__ __forceinline__ float mwarp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float mwarp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// candidates -> shared memory: cf[k][17] features, cx[k][3] points
__device__ __forceinline__ void stage_candidates(const b200r_match_args& a, float* cf, float* cx) {
  for (int e = threadIdx.x; e < a.K * kMatchC; e += blockDim.x) {
    const int k = e / kMatchC, c = e - k * kMatchC;
    cf[k * kMatchStride + c] = __ldg(a.feat_can + (size_t)a.idx[k] * kMatchC + c);
  }
  for (int e = threadIdx.x; e < a.K * 3; e += blockDim.x) {
    const int k = e / 3, c = e - k * 3;
    cx[e] = __ldg(a.xyz_can + (size_t)a.idx[k] * 3 + c);
  }
}

__global__ void __launch_bounds__(kMatchWarps * 32) match_fwd_kernel(const b200r_match_args a) {
  extern __shared__ float sm[];
  float* cf = sm;
  float* cx = sm + (size_t)a.K * kMatchStride;
  stage_candidates(a, cf, cx);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float sigma = expf(__ldg(a.logsigma));
  for (int r = blockIdx.x * kMatchWarps + warp; r < a.R; r += gridDim.x * kMatchWarps) {
    float fp[kMatchC];
#pragma unroll
    for (int c = 0; c < kMatchC; ++c) fp[c] = __ldg(a.feat_px + (size_t)r * kMatchC + c) * sigma;
    float m = -INFINITY;
    for (int k = lane; k < a.K; k += 32) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < kMatchC; ++c) s += fp[c] * cf[k * kMatchStride + c];
      m = fmaxf(m, s);
    }
    m = mwarp_max(m);
    float se = 0.f, x0 = 0.f, x1 = 0.f, x2 = 0.f;
    for (int k = lane; k < a.K; k += 32) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < kMatchC; ++c) s += fp[c] * cf[k * kMatchStride + c];
      const float pe = expf(s - m);
      se += pe;
      x0 += pe * cx[3 * k]; x1 += pe * cx[3 * k + 1]; x2 += pe * cx[3 * k + 2];
    }
    se = mwarp_sum(se); x0 = mwarp_sum(x0); x1 = mwarp_sum(x1); x2 = mwarp_sum(x2);
    if (lane == 0) {
      const float inv = 1.0f / se;
      a.xyz_matched[(size_t)r * 3] = x0 * inv; a.xyz_matched[(size_t)r * 3 + 1] = x1 * inv; a.xyz_matched[(size_t)r * 3 + 2] = x2 * inv;
      if (a.lse) a.lse[r] = m + logf(se);
    }
  }
}

// ---- backward, stage 1: block (ks, rs) = candidates [128 ks, 128 ks + 128) x rays [kMatchSplit rs, ...): thread = candidate.
// With p = prob[r,k], t = <g_out[r], xyz_k - xyz_matched[r]>, ds = p t (= dL/dscore):
//   g_xyz_k += p g_out[r],   g_feat_k += sigma ds feat_px[r],   g_logsigma += ds score[r,k].
// partial[rs][k][0:16] features, [16:19] point, [19] logsigma term.
constexpr int kPartW = 20;
__global__ void __launch_bounds__(128) match_bwd_kernel(const b200r_match_bwd_args b, float* __restrict__ partial) {
  const b200r_match_args& a = b.fwd;
  __shared__ float rs_fp[kMatchSplit][kMatchC];  // sigma * feat_px
  __shared__ float rs_aux[kMatchSplit][8];       // g_out[3], xyz_matched[3], lse, pad
  const int k = blockIdx.x * 128 + threadIdx.x;
  const int r0 = blockIdx.y * kMatchSplit;
  const int nr = min(kMatchSplit, a.R - r0);
  const float sigma = expf(__ldg(a.logsigma));
  for (int e = threadIdx.x; e < nr * kMatchC; e += 128) rs_fp[e / kMatchC][e % kMatchC] = __ldg(a.feat_px + (size_t)r0 * kMatchC + e) * sigma;
  for (int i = threadIdx.x; i < nr; i += 128) {
    const size_t r = (size_t)(r0 + i);
    rs_aux[i][0] = __ldg(b.g_out + r * 3); rs_aux[i][1] = __ldg(b.g_out + r * 3 + 1); rs_aux[i][2] = __ldg(b.g_out + r * 3 + 2);
    rs_aux[i][3] = __ldg(a.xyz_matched + r * 3); rs_aux[i][4] = __ldg(a.xyz_matched + r * 3 + 1); rs_aux[i][5] = __ldg(a.xyz_matched + r * 3 + 2);
    rs_aux[i][6] = __ldg(a.lse + r);
  }
  __syncthreads();
  if (k >= a.K) return;
  float cf[kMatchC], acc[kPartW];
  const size_t src = (size_t)a.idx[k];
#pragma unroll
  for (int c = 0; c < kMatchC; ++c) cf[c] = __ldg(a.feat_can + src * kMatchC + c);
  const float cx0 = __ldg(a.xyz_can + src * 3), cx1 = __ldg(a.xyz_can + src * 3 + 1), cx2 = __ldg(a.xyz_can + src * 3 + 2);
#pragma unroll
  for (int j = 0; j < kPartW; ++j) acc[j] = 0.f;
  for (int i = 0; i < nr; ++i) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kMatchC; ++c) s += rs_fp[i][c] * cf[c];
    const float p = expf(s - rs_aux[i][6]);
    const float g0 = rs_aux[i][0], g1 = rs_aux[i][1], g2 = rs_aux[i][2];
    const float t = g0 * (cx0 - rs_aux[i][3]) + g1 * (cx1 - rs_aux[i][4]) + g2 * (cx2 - rs_aux[i][5]);
    const float ds = p * t;
#pragma unroll
    for (int c = 0; c < kMatchC; ++c) acc[c] += ds * rs_fp[i][c];   // sigma already inside rs_fp
    acc[16] += p * g0; acc[17] += p * g1; acc[18] += p * g2;
    acc[19] += ds * s;
  }
  float* dst = partial + ((size_t)blockIdx.y * a.K + k) * kPartW;
#pragma unroll
  for (int j = 0; j < kPartW; ++j) dst[j] = acc[j];
}

// ---- backward, stage 2: sum the ray slices in a fixed order, add candidate k's gradient to row idx[k] of the dense gradients;
// block 0 also reduces the logsigma terms (fixed tree) and adds them to g_logsigma
__global__ void __launch_bounds__(256) match_scatter_kernel(const b200r_match_bwd_args b, const float* __restrict__ partial, int n_split) {
  const b200r_match_args& a = b.fwd;
  __shared__ float red[256];
  const int e = blockIdx.x * 256 + threadIdx.x;  // (k, j) with j < 19
  if (e < a.K * 19) {
    const int k = e / 19, j = e - k * 19;
    float v = 0.f;
    for (int s = 0; s < n_split; ++s) v += partial[((size_t)s * a.K + k) * kPartW + j];
    const size_t row = (size_t)a.idx[k];
    if (j < kMatchC) { if (b.g_feat_can) b.g_feat_can[row * kMatchC + j] += v; }
    else if (b.g_xyz_can) b.g_xyz_can[row * 3 + (j - kMatchC)] += v;
  }
  if (blockIdx.x == 0 && b.g_logsigma) {
    float v = 0.f;
    for (int i = threadIdx.x; i < n_split * a.K; i += 256) v += partial[(size_t)i * kPartW + 19];
    red[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) b.g_logsigma[0] += red[0];
  }
}

size_t match_partial_floats(int R, int K) { return (size_t)((R + kMatchSplit - 1) / kMatchSplit) * K * kPartW; }

cudaError_t launch_match_fwd(const b200r_match_args& a, int n_sm, cudaStream_t stream) {
  const size_t smem = (size_t)a.K * (kMatchStride + 3) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(match_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int grid = (a.R + kMatchWarps - 1) / kMatchWarps;
  if (grid > 2 * n_sm) grid = 2 * n_sm;
  match_fwd_kernel<<<grid, kMatchWarps * 32, smem, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_match_bwd(const b200r_match_bwd_args& b, float* partial, cudaStream_t stream) {
  const b200r_match_args& a = b.fwd;
  const int n_split = (a.R + kMatchSplit - 1) / kMatchSplit;
  dim3 grid((a.K + 127) / 128, n_split);
  match_bwd_kernel<<<grid, 128, 0, stream>>>(b, partial);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  match_scatter_kernel<<<(a.K * 19 + 255) / 256, 256, 0, stream>>>(b, partial, n_split);
  return cudaGetLastError();
}

}  // namespace b200r
