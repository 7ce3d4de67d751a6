// Alpha compositing along rays (render_pixel / compute_weights / integrate,
// lab4d/utils/render_utils.py:59-184) and its hand-derived backward.
//
// HBM/L2-bound.  Forward and backward: one 128-thread block per ray (block scan of the optical depth, flat coalesced walks
// over the value / gradient arrays; the backward's reverse scan runs on one warp with a running carry).  Per-ray weights stay in shared memory while the value channels are reduced.  Algorithmic traffic = 4 B x (2 + sum of channel widths) per sample read + O(c) per ray.
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/b200r.h"

namespace b200r {


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

// ---- forward: one 128-thread block per ray.  Every per-sample array of the ray is walked as one flat, fully
// coalesced run of D * nch floats (thread t takes elements t, t + 128, ...), so a ray's 37 floats per sample are in
// flight at once instead of 4 strided channels per pass.
constexpr int kFwdThreads = 128;
constexpr int kFwdWarps = kFwdThreads / 32;

struct FwdSmem {
  float* w;    // [D] weights
  float* T;    // [D] transmittance after each sample
  float* red;  // [kFwdWarps * 32] reduction scratch
};

// sum over the block; every thread gets the result.  `red` must hold kFwdWarps floats per value.
__device__ __forceinline__ float block_sum(float v, float* red, int warp, int lane) {
  v = warp_sum(v);
  __syncthreads();  // previous users of red are done
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < kFwdWarps; ++i) t += red[i];
  return t;
}

// weights of one ray with the block: sample k = c0 + tid; returns sum_k w_k (all threads)
__device__ __forceinline__ float block_ray_weights(const float* __restrict__ dens, const float* __restrict__ dl, int D, int warp,
                                                   int lane, float* w_s, float* T_s, float* red) {
  float carry = 0.f, msum = 0.f;
  for (int c0 = 0; c0 < D; c0 += kFwdThreads) {
    const int k = c0 + (int)threadIdx.x;
    const float tau = k < D ? dens[k] * dl[k] : 0.f;
    const float incl_w = warp_incl_scan(tau, lane);
    float prev_w = __shfl_up_sync(0xffffffffu, incl_w, 1);
    if (lane == 0) prev_w = 0.f;
    __syncthreads();
    if (lane == 31) red[warp] = incl_w;  // warp totals
    __syncthreads();
    float off = carry, tot = 0.f;
#pragma unroll
    for (int i = 0; i < kFwdWarps; ++i) {
      if (i < warp) off += red[i];
      tot += red[i];
    }
    const float w = (1.f - expf(-tau)) * expf(-(prev_w + off));
    if (k < D) {
      w_s[k] = w;
      if (T_s) T_s[k] = expf(-(incl_w + off));
      msum += w;
    }
    carry += tot;
  }
  return block_sum(msum, red, warp, lane);
}

__global__ void __launch_bounds__(kFwdThreads) composite_fwd_kernel(const b200r_composite_args a) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int D = a.D;
  float* w_s = sm;
  float* T_s = sm + D;
  float* g_s = sm + 2 * D;                 // scratch weights (gauss)
  float* red = sm + 3 * D;                 // kFwdWarps * 32 floats
  const int r = blockIdx.x;
  const size_t base = (size_t)r * D;
  const float mask = block_ray_weights(a.density + base, a.deltas + base, D, warp, lane, w_s, T_s, red);
  __syncthreads();  // w_s / T_s visible to every thread
  if (tid == 0 && a.mask) a.mask[r] = mask;
  if (a.weights)
    for (int k = tid; k < D; k += kFwdThreads) a.weights[base + k] = w_s[k];
  if (a.transmit)
    for (int k = tid; k < D; k += kFwdThreads) a.transmit[base + k] = T_s[k];
  const float inv = 1.0f / (mask + 1e-6f);

  for (int c = 0; c < a.n_channels; ++c) {
    const int nch = a.nch[c], mode = a.mode[c];
    const float* __restrict__ src = a.src[c] + base * nch;
    float* dst = a.dst[c];
    const int n_el = D * nch;
    if (mode == B200R_CH_NORM || mode == B200R_CH_NORM_FROZEN) {
      if (nch <= 32 && (32 % nch) == 0) {
        // channel of a thread is fixed: j = tid % nch
        float acc = 0.f;
        const int sh = __ffs(nch) - 1;  // nch is a power of two here
        for (int e = tid; e < n_el; e += kFwdThreads) acc += w_s[e >> sh] * src[e];
        for (int o = nch; o < 32; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        __syncthreads();
        if (lane < nch) red[warp * 32 + lane] = acc;
        __syncthreads();
        if (tid < nch) {
          float t = 0.f;
#pragma unroll
          for (int i = 0; i < kFwdWarps; ++i) t += red[i * 32 + tid];
          dst[(size_t)r * nch + tid] = t * inv;
        }
      } else if (nch == 3) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int e = tid; e < n_el; e += kFwdThreads) {
          const int k = e / 3, j = e - 3 * k;
          const float v = w_s[k] * src[e];
          if (j == 0) a0 += v; else if (j == 1) a1 += v; else a2 += v;
        }
        a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
        __syncthreads();
        if (lane == 0) { red[warp * 32] = a0; red[warp * 32 + 1] = a1; red[warp * 32 + 2] = a2; }
        __syncthreads();
        if (tid < 3) {
          float t = 0.f;
#pragma unroll
          for (int i = 0; i < kFwdWarps; ++i) t += red[i * 32 + tid];
          dst[(size_t)r * 3 + tid] = t * inv;
        }
      } else {  // any other width: one channel at a time
        for (int j = 0; j < nch; ++j) {
          float acc = 0.f;
          for (int k = tid; k < D; k += kFwdThreads) acc += w_s[k] * src[(size_t)k * nch + j];
          const float t = block_sum(acc, red, warp, lane);
          if (tid == 0) dst[(size_t)r * nch + j] = t * inv;
        }
      }
    } else if (mode == B200R_CH_MEAN) {
      float acc = 0.f;
      for (int e = tid; e < n_el; e += kFwdThreads) acc += src[e];
      const float t = block_sum(acc, red, warp, lane);
      if (tid == 0) dst[r] = t / (float)n_el;
    } else if (mode == B200R_CH_FLOW) {
      float sw = 0.f;
      for (int k = tid; k < D; k += kFwdThreads) sw += w_s[k] * src[(size_t)k * 3 + 2];
      sw = block_sum(sw, red, warp, lane);
      const float invf = 1.0f / (sw + 1e-6f);
      float sx = 0.f, sy = 0.f;
      for (int k = tid; k < D; k += kFwdThreads) {
        const float wf = w_s[k] * src[(size_t)k * 3 + 2] * invf;
        sx += wf * src[(size_t)k * 3];
        sy += wf * src[(size_t)k * 3 + 1];
      }
      sx = block_sum(sx, red, warp, lane);
      sy = block_sum(sy, red, warp, lane);
      if (tid == 0) { dst[(size_t)r * 2] = sx; dst[(size_t)r * 2 + 1] = sy; }
    } else if (mode == B200R_CH_WEIGHTSUM) {
      const float m2 = block_ray_weights(src, a.deltas + base, D, warp, lane, g_s, nullptr, red);
      if (tid == 0) dst[r] = m2;
    } else if (mode == B200R_CH_VIS) {
      float s0 = 0.f, s1 = 0.f;
      for (int k = tid; k < D; k += kFwdThreads) {
        s0 += log_sigmoid(src[k]) * T_s[k];
        s1 += T_s[k];
      }
      s0 = block_sum(s0, red, warp, lane);
      s1 = block_sum(s1, red, warp, lane);
      if (tid == 0) { dst[(size_t)r * 2] = s0; dst[(size_t)r * 2 + 1] = s1; }
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

// One 128-thread block per ray, like the forward: every per-sample array of the ray (values in, gradients out) is walked as
// one flat, fully coalesced run of D * nch floats; the per-sample weight gradient gw_s[k] collects the channels' contributions
// (lanes that share a sample reduce by shuffle, one writer per sample: no atomics); warp 0 then runs the reverse scan.
__global__ void __launch_bounds__(kFwdThreads) composite_bwd_kernel(const b200r_composite_bwd_args b) {
  extern __shared__ float sm[];
  const b200r_composite_args& a = b.fwd;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int D = a.D;
  float* w_s = sm;
  float* T_s = sm + D;
  float* gw_s = sm + 2 * D;
  float* red = sm + 3 * D;  // kFwdWarps * 32 floats
  const int r = blockIdx.x;
  const size_t base = (size_t)r * D;
  const float mask = block_ray_weights(a.density + base, a.deltas + base, D, warp, lane, w_s, T_s, red);
  const float inv = 1.0f / (mask + 1e-6f);
  const float gm = b.g_mask ? b.g_mask[r] : 0.f;
  for (int k = tid; k < D; k += kFwdThreads) gw_s[k] = gm;
  __syncthreads();

  for (int c = 0; c < a.n_channels; ++c) {
    const int nch = a.nch[c], mode = a.mode[c];
    const float* __restrict__ src = a.src[c] + base * nch;
    const float* g = b.g_dst[c];
    float* gs = b.g_src[c] ? b.g_src[c] + base * nch : nullptr;
    const int n_el = D * nch;
    if (mode == B200R_CH_NORM || mode == B200R_CH_NORM_FROZEN) {
      if (!g) { if (gs) for (int e = tid; e < n_el; e += kFwdThreads) gs[e] = 0.f; continue; }
      const bool live_w = mode == B200R_CH_NORM;  // frozen channels: weights detached
      if (nch <= 32 && (32 % nch) == 0) {
        const int sh = __ffs(nch) - 1, j = tid & (nch - 1);  // the channel of a thread is fixed
        const float gj = g[(size_t)r * nch + j];
        float oj = 0.f;
        if (live_w) {  // out_j = sum_k w_k inv v_kj
          float acc = 0.f;
          for (int e = tid; e < n_el; e += kFwdThreads) acc += w_s[e >> sh] * src[e];
          for (int o = nch; o < 32; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
          __syncthreads();
          if (lane < nch) red[warp * 32 + lane] = acc;
          __syncthreads();
#pragma unroll
          for (int i = 0; i < kFwdWarps; ++i) oj += red[i * 32 + j];
          oj *= inv;
        }
        for (int e0 = 0; e0 < n_el; e0 += kFwdThreads) {  // uniform trip count: the shuffles below need whole warps
          const int e = e0 + tid;
          const bool in = e < n_el;
          const int k = in ? e >> sh : 0;
          if (in && gs) gs[e] = gj * w_s[k] * inv;
          if (live_w) {
            float cg = in ? gj * (src[e] - oj) * inv : 0.f;
            for (int o = 1; o < nch; o <<= 1) cg += __shfl_xor_sync(0xffffffffu, cg, o);
            if (in && j == 0) gw_s[k] += cg;
          }
        }
      } else {  // any other width (3: rgb, xyz): one thread per sample
        float* o_s = red;  // out_j, j < nch <= 16
        if (live_w) {
          for (int j = 0; j < nch; ++j) {
            float acc = 0.f;
            for (int k = tid; k < D; k += kFwdThreads) acc += w_s[k] * src[(size_t)k * nch + j];
            const float t = block_sum(acc, red + 32, warp, lane) * inv;
            if (tid == 0) o_s[j] = t;
          }
          __syncthreads();
        }
        for (int k = tid; k < D; k += kFwdThreads) {
          float cg = 0.f;
          for (int j = 0; j < nch; ++j) {
            const float gj = g[(size_t)r * nch + j];
            if (gs) gs[(size_t)k * nch + j] = gj * w_s[k] * inv;
            if (live_w) cg += gj * (src[(size_t)k * nch + j] - o_s[j]) * inv;
          }
          if (live_w) gw_s[k] += cg;
        }
      }
      __syncthreads();
    } else if (mode == B200R_CH_MEAN) {
      if (gs) {
        const float gv = g ? g[r] / (float)n_el : 0.f;
        for (int e = tid; e < n_el; e += kFwdThreads) gs[e] = gv;
      }
    } else if (mode == B200R_CH_FLOW) {
      if (!g) { if (gs) for (int e = tid; e < D * 3; e += kFwdThreads) gs[e] = 0.f; continue; }
      float sw = 0.f, sx = 0.f, sy = 0.f;
      for (int k = tid; k < D; k += kFwdThreads) {
        const float wf = w_s[k] * src[(size_t)k * 3 + 2];
        sw += wf; sx += wf * src[(size_t)k * 3]; sy += wf * src[(size_t)k * 3 + 1];
      }
      sw = block_sum(sw, red, warp, lane); sx = block_sum(sx, red, warp, lane); sy = block_sum(sy, red, warp, lane);
      const float invf = 1.0f / (sw + 1e-6f);
      const float ox = sx * invf, oy = sy * invf;
      const float gx = g[(size_t)r * 2], gy = g[(size_t)r * 2 + 1];
      for (int k = tid; k < D; k += kFwdThreads) {
        const float val = src[(size_t)k * 3 + 2];
        const float vx = src[(size_t)k * 3], vy = src[(size_t)k * 3 + 1];
        gw_s[k] += (gx * (vx - ox) + gy * (vy - oy)) * invf * val;
        if (gs) {
          gs[(size_t)k * 3] = gx * w_s[k] * val * invf;
          gs[(size_t)k * 3 + 1] = gy * w_s[k] * val * invf;
          gs[(size_t)k * 3 + 2] = 0.f;  // validity flag is piecewise constant
        }
      }
      __syncthreads();
    } else if (mode == B200R_CH_VIS) {
      // transmittance is detached (render_utils.py:83): only the logit receives gradient
      if (gs) {
        const float g0 = g ? g[(size_t)r * 2] : 0.f;
        for (int k = tid; k < D; k += kFwdThreads) {
          const float x = src[k];
          gs[k] = g0 * T_s[k] * (1.f / (1.f + expf(x)));  // d logsigmoid = sigmoid(-x)
        }
      }
    }
  }
  __syncthreads();
  if (warp != 0) return;  // the reverse scans are O(D): one warp
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
  const size_t smem = ((size_t)3 * a.D + kFwdWarps * 32) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(composite_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  composite_fwd_kernel<<<a.R, kFwdThreads, smem, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_composite_bwd(const b200r_composite_bwd_args& b, cudaStream_t stream) {
  const size_t smem = ((size_t)3 * b.fwd.D + kFwdWarps * 32 + 32) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(composite_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  composite_bwd_kernel<<<b.fwd.R, kFwdThreads, smem, stream>>>(b);
  return cudaGetLastError();
}

}  // namespace b200r
