// Per-pixel reconstruction losses of dvr_model (lab4d/engine/model.py): get_mask_balance_wt (:386-412), compute_recon_loss
// (:415-498), mask_losses (:520-574) and apply_loss_weights (:576-611) for the terms compute_recon_loss creates - about forty
// small torch kernels per step in the reference (plus their backward) - as ONE single-block forward kernel (R = M*N <= a few
// thousand rays: three passes over ~50 floats per ray, fixed-order reductions, deterministic) and one elementwise backward kernel.
//
// Term k (order of the reference's loss_dict): 0 mask, 1 feature, 2 feat_reproj, 3 rgb, 4 depth, 5 flow, 6 vis, 7 reg_gauss_mask.
// Every term ends as  mean over the entries with value > 0  [/ train_res for pixel units]  * weight.
#include <cuda_runtime.h>
#include <math.h>

#include "kernels.h"

namespace b200r {

constexpr int kLossThreads = 1024;
constexpr int kLossTerms = B200R_LOSS_TERMS;
enum : int { T_MASK = 0, T_FEAT, T_REPROJ, T_RGB, T_DEPTH, T_FLOW, T_VIS, T_GAUSS };
enum : int { FT_FG = 0, FT_BG = 1, FT_COMP = 2 };
// stats layout: [0,8) sum of positive values, [8,16) their count, 16 pos_wt, 17 neg_wt, 18 balanced (1 / 0)

// values of one ray's terms; v[T_RGB..] rgb has three entries (rgbv)
struct RayTerms {
  float v[kLossTerms];
  float rgbv[3];
};

__device__ __forceinline__ float ldz(const float* p, size_t i) { return p ? __ldg(p + i) : 0.f; }

__device__ __forceinline__ void ray_terms(const b200r_loss_args& a, int r, float pos_wt, float neg_wt, bool balanced, RayTerms& t) {
  const int f = r / a.N;
  const float m = __ldg(a.b_mask + r), v2 = __ldg(a.b_vis2d + r), det = __ldg(a.b_is_detected + f);
  const float wbal = balanced ? 0.5f * pos_wt * m + 0.5f * neg_wt * (1.f - m) : 1.f;
  const float mtype = a.field_type == FT_BG ? (1.f - m) * v2 : (a.field_type == FT_FG ? m * v2 : v2);
  const float rm = __ldg(a.r_mask + r);
  const float rfg = a.field_type == FT_COMP ? __ldg(a.r_mask_fg + r) : rm;
  float lm;
  if (a.field_type == FT_BG) lm = (rm - 1.f) * (rm - 1.f);
  else if (a.field_type == FT_FG) lm = (rfg - m) * (rfg - m) * wbal;
  else lm = (rfg - m) * (rfg - m) * wbal + (rm - 1.f) * (rm - 1.f);
  t.v[T_MASK] = lm * v2 * det;
  t.v[T_FEAT] = 0.f; t.v[T_REPROJ] = 0.f; t.v[T_GAUSS] = 0.f;
  if (a.field_type != FT_BG) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) { const float d = __ldg(a.a_feature + (size_t)r * 16 + c) - __ldg(a.b_feature + (size_t)r * 16 + c); s += d * d; }
    t.v[T_FEAT] = sqrtf(s) * m * det;
    const float dx = __ldg(a.a_xy_reproj + (size_t)r * 2) - __ldg(a.b_hxy + (size_t)r * 3), dy = __ldg(a.a_xy_reproj + (size_t)r * 2 + 1) - __ldg(a.b_hxy + (size_t)r * 3 + 1);
    t.v[T_REPROJ] = sqrtf(dx * dx + dy * dy) * m * det;
    if (a.a_gauss_mask) { const float d = __ldg(a.a_gauss_mask + r) - rfg; t.v[T_GAUSS] = d * d; }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) { const float d = __ldg(a.r_rgb + (size_t)r * 3 + c) - __ldg(a.b_rgb + (size_t)r * 3 + c); t.rgbv[c] = d * d * mtype; }
  t.v[T_RGB] = 0.f;
  t.v[T_DEPTH] = fabsf(__ldg(a.r_depth + r) - __ldg(a.b_depth + r)) * mtype;
  {
    const float dx = __ldg(a.r_flow + (size_t)r * 2) - __ldg(a.b_flow + (size_t)r * 2), dy = __ldg(a.r_flow + (size_t)r * 2 + 1) - __ldg(a.b_flow + (size_t)r * 2 + 1);
    t.v[T_FLOW] = sqrtf(dx * dx + dy * dy) * (__ldg(a.b_flow_uct + r) > 0.f ? 1.f : 0.f) * mtype;
  }
  t.v[T_VIS] = (ldz(a.vis_fg, r) + 0.01f * ldz(a.vis_bg, r)) * mtype;
}

// block-wide sums of n (<= 16) values; every thread gets the totals.  red: [32][16] floats
template <int NV>
__device__ __forceinline__ void block_sums(float (&v)[NV], float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < NV; ++i) red[warp * 16 + i] = v[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float t = 0.f;
    for (int w = 0; w < kLossThreads / 32; ++w) t += red[w * 16 + i];
    v[i] = t;
  }
}

__global__ void __launch_bounds__(kLossThreads) loss_fwd_kernel(const b200r_loss_args a) {
  __shared__ float red[32 * 16];
  const int R = a.M * a.N;
  // ---- pass A: mask balance (model.py:386-412)
  float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};  // sum mask, sum (1-mask), sum vis', sum mask[vis'>0], sum (1-mask)[vis'>0]
  for (int r = threadIdx.x; r < R; r += kLossThreads) {
    const float m = __ldg(a.b_mask + r), v2 = __ldg(a.b_vis2d + r) * __ldg(a.b_is_detected + r / a.N);
    s[0] += m; s[1] += 1.f - m; s[2] += v2;
    if (v2 > 0.f) { s[3] += m; s[4] += 1.f - m; }
  }
  block_sums<5>(s, red);
  const bool balanced = s[0] > 0.f && s[1] > 0.f;
  const float pos_wt = s[2] / s[3], neg_wt = s[2] / s[4];
  // ---- pass B: sums and counts of the positive entries of every term
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int r = threadIdx.x; r < R; r += kLossThreads) {
    RayTerms t;
    ray_terms(a, r, pos_wt, neg_wt, balanced, t);
#pragma unroll
    for (int k = 0; k < kLossTerms; ++k)
      if (k != T_RGB && t.v[k] > 0.f) { acc[k] += t.v[k]; acc[8 + k] += 1.f; }
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (t.rgbv[c] > 0.f) { acc[T_RGB] += t.rgbv[c]; acc[8 + T_RGB] += 1.f; }
  }
  block_sums<16>(acc, red);
  if (threadIdx.x < kLossTerms) {
    const int k = threadIdx.x;
    float v = acc[k] / acc[8 + k];  // mean of an empty selection is NaN, like torch's
    if (k == T_FLOW || k == T_REPROJ) v /= a.train_res;
    a.loss[k] = v * a.wt[k];
  }
  if (threadIdx.x < 16) a.stats[threadIdx.x] = acc[threadIdx.x];
  if (threadIdx.x == 0) { a.stats[16] = pos_wt; a.stats[17] = neg_wt; a.stats[18] = balanced ? 1.f : 0.f; }
}

// ---- backward: dL/d(every rendered input) = sum_k g_loss[k] wt[k] [1/train_res] / count_k * d v_k / d input, entries with v_k > 0
__global__ void __launch_bounds__(256) loss_bwd_kernel(const b200r_loss_bwd_args b) {
  const b200r_loss_args& a = b.fwd;
  const int R = a.M * a.N;
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= R) return;
  const float pos_wt = a.stats[16], neg_wt = a.stats[17];
  const bool balanced = a.stats[18] != 0.f;
  RayTerms t;
  ray_terms(a, r, pos_wt, neg_wt, balanced, t);
  float sc[kLossTerms];
#pragma unroll
  for (int k = 0; k < kLossTerms; ++k) {
    float c = __ldg(b.g_loss + k) * a.wt[k] / a.stats[8 + k];
    if (k == T_FLOW || k == T_REPROJ) c /= a.train_res;
    sc[k] = c;
  }
  const int f = r / a.N;
  const float m = __ldg(a.b_mask + r), v2 = __ldg(a.b_vis2d + r), det = __ldg(a.b_is_detected + f);
  const float wbal = balanced ? 0.5f * pos_wt * m + 0.5f * neg_wt * (1.f - m) : 1.f;
  const float mtype = a.field_type == FT_BG ? (1.f - m) * v2 : (a.field_type == FT_FG ? m * v2 : v2);
  const float rm = __ldg(a.r_mask + r);
  const float rfg = a.field_type == FT_COMP ? __ldg(a.r_mask_fg + r) : rm;
  // mask term
  float g_rm = 0.f, g_rfg = 0.f;
  if (t.v[T_MASK] > 0.f) {
    const float w = sc[T_MASK] * v2 * det;
    if (a.field_type != FT_BG) g_rfg = w * 2.f * (rfg - m) * wbal;
    if (a.field_type != FT_FG) g_rm = w * 2.f * (rm - 1.f);
  }
  if (a.field_type == FT_COMP) {
    if (b.g_mask_fg) b.g_mask_fg[r] = g_rfg;
    if (b.g_mask) b.g_mask[r] = g_rm;
  } else if (b.g_mask) {
    b.g_mask[r] = g_rm + g_rfg;
  }
  if (a.field_type != FT_BG) {
    if (b.g_feature) {
      float d[16], s = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) { d[c] = __ldg(a.a_feature + (size_t)r * 16 + c) - __ldg(a.b_feature + (size_t)r * 16 + c); s += d[c] * d[c]; }
      const float w = t.v[T_FEAT] > 0.f ? sc[T_FEAT] * m * det * rsqrtf(s) : 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) b.g_feature[(size_t)r * 16 + c] = w * d[c];
    }
    if (b.g_xy_reproj) {
      const float dx = __ldg(a.a_xy_reproj + (size_t)r * 2) - __ldg(a.b_hxy + (size_t)r * 3), dy = __ldg(a.a_xy_reproj + (size_t)r * 2 + 1) - __ldg(a.b_hxy + (size_t)r * 3 + 1);
      const float w = t.v[T_REPROJ] > 0.f ? sc[T_REPROJ] * m * det * rsqrtf(dx * dx + dy * dy) : 0.f;
      b.g_xy_reproj[(size_t)r * 2] = w * dx; b.g_xy_reproj[(size_t)r * 2 + 1] = w * dy;
    }
    if (b.g_gauss_mask && a.a_gauss_mask) b.g_gauss_mask[r] = t.v[T_GAUSS] > 0.f ? sc[T_GAUSS] * 2.f * (__ldg(a.a_gauss_mask + r) - rfg) : 0.f;
  }
  if (b.g_rgb) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      b.g_rgb[(size_t)r * 3 + c] = t.rgbv[c] > 0.f ? sc[T_RGB] * 2.f * (__ldg(a.r_rgb + (size_t)r * 3 + c) - __ldg(a.b_rgb + (size_t)r * 3 + c)) * mtype : 0.f;
  }
  if (b.g_depth) {
    const float d = __ldg(a.r_depth + r) - __ldg(a.b_depth + r);
    b.g_depth[r] = t.v[T_DEPTH] > 0.f ? sc[T_DEPTH] * (d > 0.f ? 1.f : -1.f) * mtype : 0.f;
  }
  if (b.g_flow) {
    const float dx = __ldg(a.r_flow + (size_t)r * 2) - __ldg(a.b_flow + (size_t)r * 2), dy = __ldg(a.r_flow + (size_t)r * 2 + 1) - __ldg(a.b_flow + (size_t)r * 2 + 1);
    const float w = t.v[T_FLOW] > 0.f ? sc[T_FLOW] * mtype * rsqrtf(dx * dx + dy * dy) : 0.f;
    b.g_flow[(size_t)r * 2] = w * dx; b.g_flow[(size_t)r * 2 + 1] = w * dy;
  }
  const float gv = t.v[T_VIS] > 0.f ? sc[T_VIS] * mtype : 0.f;
  if (b.g_vis_fg) b.g_vis_fg[r] = gv;
  if (b.g_vis_bg) b.g_vis_bg[r] = 0.01f * gv;
}

cudaError_t launch_loss_fwd(const b200r_loss_args& a, cudaStream_t stream) {
  loss_fwd_kernel<<<1, kLossThreads, 0, stream>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_loss_bwd(const b200r_loss_bwd_args& b, cudaStream_t stream) {
  const int R = b.fwd.M * b.fwd.N;
  loss_bwd_kernel<<<(R + 255) / 256, 256, 0, stream>>>(b);
  return cudaGetLastError();
}

}  // namespace b200r
