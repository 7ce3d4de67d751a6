// Alpha compositing along rays (render_pixel / compute_weights / integrate,
// lab4d/utils/render_utils.py:59-184) and its hand-derived backward.
//
// HBM-bound: one warp per ray, lanes across samples, warp-shuffle inclusive scan of the optical
// depth with a running carry, per-ray weights kept in shared memory while the value channels are
// reduced.  Algorithmic traffic = 4 B x (2 + sum of channel widths) per sample read + O(c) per ray.
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/b200r.h"

namespace b200r {

constexpr int kWarpsPerBlock = 8;
constexpr int kWarpsPerRayFwd = 1;  // (tried 4 warps per ray with the channels spread out: no gain on B200)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_incl_scan(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ float log_sigmoid(float x) { return fminf(x, 0.f) - log1pf(expf(-fabsf(x))); }

// weights w_k = (1-exp(-tau_k)) exp(-sum_{j<k} tau_j),  T_k = exp(-sum_{j<=k} tau_j); returns sum_k w_k
__device__ __forceinline__ float ray_weights(const float* __restrict__ dens, const float* __restrict__ dl, int D, int lane,
                                             float* w_s, float* T_s) {
  float carry = 0.f, msum = 0.f;
  for (int c0 = 0; c0 < D; c0 += 32) {
    const int k = c0 + lane;
    const float tau = k < D ? dens[k] * dl[k] : 0.f;
    const float incl = warp_incl_scan(tau, lane) + carry;
    float prev = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) prev = carry;
    const float w = (1.f - expf(-tau)) * expf(-prev);
    if (k < D) {
      w_s[k] = w;
      if (T_s) T_s[k] = expf(-incl);
      msum += w;
    }
    carry = __shfl_sync(0xffffffffu, incl, 31);
  }
  return warp_sum(msum);
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) composite_fwd_kernel(const b200r_composite_args a) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = a.D;
  float* w_s = sm + (size_t)warp * 3 * D;
  float* T_s = w_s + D;
  float* g_s = T_s + D;  // scratch weights (gauss)
  const int r = blockIdx.x * (kWarpsPerBlock / kWarpsPerRayFwd) + warp / kWarpsPerRayFwd;
  const int cw = warp % kWarpsPerRayFwd;  // this warp's share of the channels; the weights are recomputed per warp
  if (r >= a.R) return;
  const size_t base = (size_t)r * D;
  const float mask = ray_weights(a.density + base, a.deltas + base, D, lane, w_s, T_s);
  __syncwarp();
  if (cw == 0) {
    if (lane == 0 && a.mask) a.mask[r] = mask;
    if (a.weights)
      for (int k = lane; k < D; k += 32) a.weights[base + k] = w_s[k];
    if (a.transmit)
      for (int k = lane; k < D; k += 32) a.transmit[base + k] = T_s[k];
  }
  const float inv = 1.0f / (mask + 1e-6f);

  for (int c = cw; c < a.n_channels; c += kWarpsPerRayFwd) {
    const int nch = a.nch[c], mode = a.mode[c];
    const float* __restrict__ src = a.src[c] + base * nch;
    float* dst = a.dst[c];
    if (mode == B200R_CH_NORM || mode == B200R_CH_NORM_FROZEN) {
      for (int j0 = 0; j0 < nch; j0 += 4) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = lane; k < D; k += 32) {
          const float wn = w_s[k] * inv;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j0 + j < nch) acc[j] += wn * src[(size_t)k * nch + j0 + j];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float t = warp_sum(acc[j]);
          if (lane == 0 && j0 + j < nch) dst[(size_t)r * nch + j0 + j] = t;
        }
      }
    } else if (mode == B200R_CH_MEAN) {
      float acc = 0.f;
      for (int e = lane; e < D * nch; e += 32) acc += src[e];
      acc = warp_sum(acc);
      if (lane == 0) dst[r] = acc / (float)(D * nch);
    } else if (mode == B200R_CH_FLOW) {
      float sw = 0.f, sx = 0.f, sy = 0.f;
      for (int k = lane; k < D; k += 32) {
        const float wf = w_s[k] * src[(size_t)k * 3 + 2];
        sw += wf;
      }
      sw = warp_sum(sw);
      const float invf = 1.0f / (sw + 1e-6f);
      for (int k = lane; k < D; k += 32) {
        const float wf = w_s[k] * src[(size_t)k * 3 + 2] * invf;
        sx += wf * src[(size_t)k * 3];
        sy += wf * src[(size_t)k * 3 + 1];
      }
      sx = warp_sum(sx);
      sy = warp_sum(sy);
      if (lane == 0) { dst[(size_t)r * 2] = sx; dst[(size_t)r * 2 + 1] = sy; }
    } else if (mode == B200R_CH_WEIGHTSUM) {
      const float m2 = ray_weights(src, a.deltas + base, D, lane, g_s, nullptr);
      if (lane == 0) dst[r] = m2;
    } else if (mode == B200R_CH_VIS) {
      float s0 = 0.f, s1 = 0.f;
      for (int k = lane; k < D; k += 32) {
        s0 += log_sigmoid(src[k]) * T_s[k];
        s1 += T_s[k];
      }
      s0 = warp_sum(s0);
      s1 = warp_sum(s1);
      if (lane == 0) { dst[(size_t)r * 2] = s0; dst[(size_t)r * 2 + 1] = s1; }
    }
  }
}

// ------------------------------------------------------------------------------------ backward
// With tau_k = sigma_k delta_k, c_k = sum_{j<=k} tau_j, w_k = (1-e^{-tau_k}) e^{-c_{k-1}}:
//   dL/dtau_k = gw_k e^{-c_k}  -  sum_{j>k} gw_j w_j        (gw = dL/dw)
// and for a normalised channel out = sum_k w_k v_k / (m+eps), m = sum w:
//   dL/dw_k += g . (v_k - out) / (m+eps),   dL/dv_k = g w_k/(m+eps).
// A density-type channel (WEIGHTSUM) has its own tau and gw = g.
__device__ __forceinline__ void tau_backward(const float* gw_s, const float* w_s, const float* __restrict__ dens,
                                             const float* __restrict__ dl, int D, int lane, float* __restrict__ g_out,
                                             bool accumulate) {
  // suffix sum of gw_j w_j, walking the ray backwards in 32-sample chunks
  float carry = 0.f;  // sum over samples after the current chunk
  // need c_k (inclusive cumulative tau) again: recompute forward, store e^{-c_k} in place of gw? keep simple:
  // first pass forward to get c_k into registers chunk by chunk is not possible backwards -> two passes.
  float total_tau = 0.f;
  for (int c0 = 0; c0 < D; c0 += 32) {
    const int k = c0 + lane;
    total_tau += k < D ? dens[k] * dl[k] : 0.f;
  }
  total_tau = warp_sum(total_tau);
  float tail_tau = 0.f;  // sum of tau over samples after the current chunk
  const int nchunk = (D + 31) / 32;
  for (int ci = nchunk - 1; ci >= 0; --ci) {
    const int k = ci * 32 + lane;
    const bool in = k < D;
    const float tau = in ? dens[k] * dl[k] : 0.f;
    const float gww = in ? gw_s[k] * w_s[k] : 0.f;
    // exclusive suffix within the chunk: S_k = sum_{j>k, j in chunk} gww_j
    const float incl_tau = warp_incl_scan(tau, lane);
    const float incl_g = warp_incl_scan(gww, lane);
    const float chunk_tau = __shfl_sync(0xffffffffu, incl_tau, 31);
    const float chunk_g = __shfl_sync(0xffffffffu, incl_g, 31);
    const float suffix_g = (chunk_g - incl_g) + carry;
    const float c_k = (total_tau - tail_tau - chunk_tau) + incl_tau;  // inclusive cumulative tau at k
    if (in) {
      const float gtau = gw_s[k] * expf(-c_k) - suffix_g;
      const float gd = gtau * dl[k];
      if (accumulate) g_out[k] += gd; else g_out[k] = gd;
    }
    carry += chunk_g;
    tail_tau += chunk_tau;
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) composite_bwd_kernel(const b200r_composite_bwd_args b) {
  extern __shared__ float sm[];
  const b200r_composite_args& a = b.fwd;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = a.D;
  float* w_s = sm + (size_t)warp * 3 * D;
  float* T_s = w_s + D;
  float* gw_s = T_s + D;
  const int r = blockIdx.x * kWarpsPerBlock + warp;
  if (r >= a.R) return;
  const size_t base = (size_t)r * D;
  const float mask = ray_weights(a.density + base, a.deltas + base, D, lane, w_s, T_s);
  __syncwarp();
  const float inv = 1.0f / (mask + 1e-6f);
  const float gm = b.g_mask ? b.g_mask[r] : 0.f;
  for (int k = lane; k < D; k += 32) gw_s[k] = gm;
  __syncwarp();

  for (int c = 0; c < a.n_channels; ++c) {
    const int nch = a.nch[c], mode = a.mode[c];
    const float* __restrict__ src = a.src[c] + base * nch;
    const float* g = b.g_dst[c];
    float* gs = b.g_src[c] ? b.g_src[c] + base * nch : nullptr;
    if (mode == B200R_CH_NORM || mode == B200R_CH_NORM_FROZEN) {
      if (!g) { if (gs) for (int e = lane; e < D * nch; e += 32) gs[e] = 0.f; continue; }
      // out_j recomputed: sum_k wn_k v_kj
      for (int j = 0; j < nch; ++j) {
        const float gj = g[(size_t)r * nch + j];
        float o = 0.f;
        if (mode == B200R_CH_NORM) {
          for (int k = lane; k < D; k += 32) o += w_s[k] * inv * src[(size_t)k * nch + j];
          o = warp_sum(o);
        }
        for (int k = lane; k < D; k += 32) {
          if (gs) gs[(size_t)k * nch + j] = gj * w_s[k] * inv;
          if (mode == B200R_CH_NORM) gw_s[k] += gj * (src[(size_t)k * nch + j] - o) * inv;
        }
      }
    } else if (mode == B200R_CH_MEAN) {
      if (gs) {
        const float gv = g ? g[r] / (float)(D * nch) : 0.f;
        for (int e = lane; e < D * nch; e += 32) gs[e] = gv;
      }
    } else if (mode == B200R_CH_FLOW) {
      if (!g) { if (gs) for (int e = lane; e < D * 3; e += 32) gs[e] = 0.f; continue; }
      float sw = 0.f, sx = 0.f, sy = 0.f;
      for (int k = lane; k < D; k += 32) {
        const float wf = w_s[k] * src[(size_t)k * 3 + 2];
        sw += wf; sx += wf * src[(size_t)k * 3]; sy += wf * src[(size_t)k * 3 + 1];
      }
      sw = warp_sum(sw); sx = warp_sum(sx); sy = warp_sum(sy);
      const float invf = 1.0f / (sw + 1e-6f);
      const float ox = sx * invf, oy = sy * invf;
      const float gx = g[(size_t)r * 2], gy = g[(size_t)r * 2 + 1];
      for (int k = lane; k < D; k += 32) {
        const float val = src[(size_t)k * 3 + 2];
        const float vx = src[(size_t)k * 3], vy = src[(size_t)k * 3 + 1];
        gw_s[k] += (gx * (vx - ox) + gy * (vy - oy)) * invf * val;
        if (gs) {
          gs[(size_t)k * 3] = gx * w_s[k] * val * invf;
          gs[(size_t)k * 3 + 1] = gy * w_s[k] * val * invf;
          gs[(size_t)k * 3 + 2] = 0.f;  // validity flag is piecewise constant
        }
      }
    } else if (mode == B200R_CH_VIS) {
      // transmittance is detached (render_utils.py:83): only the logit receives gradient
      if (gs) {
        const float g0 = g ? g[(size_t)r * 2] : 0.f;
        for (int k = lane; k < D; k += 32) {
          const float x = src[k];
          gs[k] = g0 * T_s[k] * (1.f / (1.f + expf(x)));  // d logsigmoid = sigmoid(-x)
        }
      }
    }
  }
  __syncwarp();
  tau_backward(gw_s, w_s, a.density + base, a.deltas + base, D, lane, b.g_density + base, false);
  __syncwarp();
  // density-type channels: their own weights
  for (int c = 0; c < a.n_channels; ++c) {
    if (a.mode[c] != B200R_CH_WEIGHTSUM || !b.g_src[c]) continue;
    const float* src = a.src[c] + base;
    const float gg = b.g_dst[c] ? b.g_dst[c][r] : 0.f;
    ray_weights(src, a.deltas + base, D, lane, w_s, nullptr);
    for (int k = lane; k < D; k += 32) gw_s[k] = gg;
    __syncwarp();
    tau_backward(gw_s, w_s, src, a.deltas + base, D, lane, b.g_src[c] + base, false);
    __syncwarp();
  }
}

cudaError_t launch_composite_fwd(const b200r_composite_args& a, cudaStream_t stream) {
  const size_t smem = (size_t)kWarpsPerBlock * 3 * a.D * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(composite_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int rays_per_block = kWarpsPerBlock / kWarpsPerRayFwd;
  const int blocks = (a.R + rays_per_block - 1) / rays_per_block;
  composite_fwd_kernel<<<blocks, kWarpsPerBlock * 32, smem, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_composite_bwd(const b200r_composite_bwd_args& b, cudaStream_t stream) {
  const size_t smem = (size_t)kWarpsPerBlock * 3 * b.fwd.D * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(composite_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int blocks = (b.fwd.R + kWarpsPerBlock - 1) / kWarpsPerBlock;
  composite_bwd_kernel<<<blocks, kWarpsPerBlock * 32, smem, stream>>>(b);
  return cudaGetLastError();
}

}  // namespace b200r
