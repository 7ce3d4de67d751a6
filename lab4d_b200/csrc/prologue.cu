// Per-frame prologue: one launch builds the scratch the fused field kernel stages into shared memory.
//
//   blocks 0..4M-1 : the frame block of frame f = block / 4 (program.h FrameLayout), four blocks per frame -
//       cameras of the frame and of its flip_pair partner (nnutils/nerf.py:929-946),
//       bias rows b + W[:, code columns] @ code[f] of the layers that see a per-frame code
//       (instance / time / appearance codes are constant per frame: nnutils/base.py:140-146,
//       nerf.py:200-204, skinning.py:109-116),
//       bone tables: inverse bone transforms (utils/transforms.py:9-25) and the per-bone blend
//       transforms rest (x) t^-1 / t (x) rest^-1 as dual quaternions (nnutils/warping.py:304-314).
//   block 4M       : the constant block - plain bias rows, head weights, rest bone centres (utils/transforms.py:28-40), scalars.
// M x B rows of quaternion algebra and a few (N x 32) mat-vecs: ~0.1 % of the step's FLOPs.
#include <cuda_runtime.h>
#include <math.h>

#include "kernels.h"

namespace b200r {

struct Q4 { float w, x, y, z; };
__device__ __forceinline__ Q4 qmul(const Q4& a, const Q4& b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
}
__device__ __forceinline__ Q4 qconj(const Q4& a) { return {a.w, -a.x, -a.y, -a.z}; }
__device__ __forceinline__ Q4 qadd(const Q4& a, const Q4& b) { return {a.w + b.w, a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ Q4 ld4(const float* p) { return {p[0], p[1], p[2], p[3]}; }
__device__ __forceinline__ void st4(float* p, const Q4& q) { p[0] = q.w; p[1] = q.x; p[2] = q.y; p[3] = q.z; }

__device__ void write_cam(float* dst, const PrologueParams& p, int f) {
  for (int i = threadIdx.x; i < 24; i += blockDim.x) {
    float v = 0.f;
    if (i < 9) v = p.fr.Kinv[(size_t)f * 9 + i];
    else if (i < 11) v = p.fr.near_far[(size_t)f * 2 + (i - 9)];
    else if (i < 15) v = p.fr.field2cam_q[(size_t)f * 4 + (i - 11)];
    else if (i < 18) v = p.fr.field2cam_t[(size_t)f * 3 + (i - 15)];
    dst[i] = v;
  }
}

// Inverse of a bone transform, pre-scaled by the Gaussian bone scale, as three rows
// (R'_i0, R'_i1, R'_i2, t'_i) with R' = diag(1/gauss) R(q*), t' = (1/gauss) * 2 (qd* q)_xyz, so that the
// Gaussian bone coordinate of x is R' x + t' (utils/transforms.py:9-25, nnutils/skinning.py:122-139).
__device__ void write_binv(float* dst, const float* qr_, const float* qd_, const float* cblock_inv_gauss, int B) {
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const Q4 q = qconj(ld4(qr_ + b * 4)), qd = qconj(ld4(qd_ + b * 4));
    const Q4 t = qmul(qd, qconj(q));
    const float tx = 2.f * t.x, ty = 2.f * t.y, tz = 2.f * t.z;
    const float gx = cblock_inv_gauss[b * 4], gy = cblock_inv_gauss[b * 4 + 1], gz = cblock_inv_gauss[b * 4 + 2];
    // rotation matrix of q (the quaternion sandwich q (0,p) q* expanded; |q| = 1 up to rounding)
    const float ww = q.w * q.w, xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z;
    const float xy = q.x * q.y, xz = q.x * q.z, yz = q.y * q.z, wx = q.w * q.x, wy = q.w * q.y, wz = q.w * q.z;
    float* o = dst + b * 12;
    o[0] = gx * (ww + xx - yy - zz); o[1] = gx * 2.f * (xy - wz); o[2] = gx * 2.f * (xz + wy); o[3] = gx * tx;
    o[4] = gy * 2.f * (xy + wz); o[5] = gy * (ww - xx + yy - zz); o[6] = gy * 2.f * (yz - wx); o[7] = gy * ty;
    o[8] = gz * 2.f * (xz - wy); o[9] = gz * 2.f * (yz + wx); o[10] = gz * (ww - xx - yy + zz); o[11] = gz * tz;
  }
}
// a (x) b^-1 as (real, dual)
__device__ void write_se3(float* dst, const float* ar_, const float* ad_, const float* br_, const float* bd_, int B) {
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const Q4 ar = ld4(ar_ + b * 4), ad = ld4(ad_ + b * 4);
    const Q4 bir = qconj(ld4(br_ + b * 4)), bid = qconj(ld4(bd_ + b * 4));
    st4(dst + b * 8, qmul(ar, bir));
    st4(dst + b * 8 + 4, qadd(qmul(ar, bid), qmul(ad, bir)));
  }
}

constexpr int kFrameParts = 4;  // blocks per frame: the bias rows are dealt round-robin, the bone tables by part

// bias rows b + sum_seg W[:, col0:col0+C] @ code_seg of one frame, rows dealt round-robin over the frame's blocks
template <bool FILTER>
__device__ __forceinline__ void frame_rows(const PrologueParams& p, const FrameLayout& F, const float* const* codes, float* fb, int part) {
  int rows_total = 0;
  for (int ci = 0; ci < F.n_cond; ++ci) rows_total += F.cond[ci].n;
  for (int r = part * blockDim.x + threadIdx.x; r < rows_total; r += kFrameParts * blockDim.x) {
    int ci = 0, n = r;
    while (n >= F.cond[ci].n) { n -= F.cond[ci].n; ++ci; }
    const CondRow& c = F.cond[ci];
    if (FILTER && (!codes[c.code[0]] || (c.n_seg > 1 && !codes[c.code[1]]))) continue;
    float acc = p.par.bias[c.layer][n];
    for (int sgi = 0; sgi < c.n_seg; ++sgi) {
      const float* code = codes[c.code[sgi]];
      const float* wr = p.par.weight[c.layer] + (size_t)n * c.in_dim + c.col0[sgi];
      const int wdt = c.width[sgi];
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int k = 0;
      for (; k + 8 <= wdt; k += 8) {
        a0 += wr[k] * code[k] + wr[k + 4] * code[k + 4];
        a1 += wr[k + 1] * code[k + 1] + wr[k + 5] * code[k + 5];
        a2 += wr[k + 2] * code[k + 2] + wr[k + 6] * code[k + 6];
        a3 += wr[k + 3] * code[k + 3] + wr[k + 7] * code[k + 7];
      }
      for (; k < wdt; ++k) a0 += wr[k] * code[k];
      acc += (a0 + a1) + (a2 + a3);
    }
    fb[c.frame_off + n] = acc;
  }
}


// PT = point entries (b200r_points_fwd / b200r_warp_fwd): optional inputs may be absent.  The field entry compiles
// every such check out (with them in, this latency-bound kernel takes 47 instead of 29 us).
template <bool PT>
__global__ void __launch_bounds__(256) prologue_kernel(const __grid_constant__ PrologueParams p) {
  // let the field kernel (launched with programmatic stream serialization) start its set-up while this grid runs;
  // it waits (griddepcontrol.wait) before it reads the workspace
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int M = p.fr.M, B = p.desc.n_bones;
  float* cblock = p.workspace;
  if ((int)blockIdx.x == kFrameParts * M) {
    // ---------------------------------------------------------------- constant block
    const ConstLayout& C = p.cl;
    for (int i = threadIdx.x; i < C.n_floats; i += blockDim.x) cblock[i] = 0.f;
    __syncthreads();
    for (int l = 0; l < p.n_layers; ++l) {
      if (C.plain_off[l] < 0) continue;
      for (int i = threadIdx.x; i < p.layer_out[l]; i += blockDim.x) cblock[C.plain_off[l] + i] = p.par.bias[l][i];
    }
    const int W = p.desc.W, H = W / 2;
    for (int i = threadIdx.x; i < W; i += blockDim.x) cblock[C.sdf_w + i] = p.par.sdf_w[i];
    for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) cblock[C.rgb2_w + i] = p.par.rgb2_w[i];
    if (!PT || p.par.vis_final_w)
      for (int i = threadIdx.x; i < 64; i += blockDim.x) cblock[C.vis_w + i] = p.par.vis_final_w[i];
    if (p.desc.L_dir == 0) {
      const float* w0 = p.par.weight[p.rgb0_layer];
      const int in_dim = p.layer_in[p.rgb0_layer];
      for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) cblock[C.dir_w + i] = w0[(size_t)(i / 3) * in_dim + W + (i % 3)];
    }
    for (int b = threadIdx.x; b < B && !(PT && p.skip_bones); b += blockDim.x) {
      const Q4 qr = ld4(p.fr.rest_art_qr + b * 4), qd = ld4(p.fr.rest_art_qd + b * 4);  // frame 0
      const Q4 t = qmul(qd, qconj(qr));
      cblock[C.center + b * 4 + 0] = 2.f * t.x;
      cblock[C.center + b * 4 + 1] = 2.f * t.y;
      cblock[C.center + b * 4 + 2] = 2.f * t.z;
    }
    if (threadIdx.x == 0) {
      float* s = cblock + C.scalars;
      s[SC_IBETA] = expf(p.par.logibeta[0]);
      s[SC_INV_SCALE] = 1.0f / expf(p.par.logscale[0]);
      s[SC_WARP_IBETA] = (B > 0 && (!PT || p.par.warp_logibeta)) ? expf(p.par.warp_logibeta[0]) : 0.f;
      s[SC_SDF_B] = p.par.sdf_b[0];
      s[SC_RGB2_B0] = p.par.rgb2_b[0]; s[SC_RGB2_B1] = p.par.rgb2_b[1]; s[SC_RGB2_B2] = p.par.rgb2_b[2];
      s[SC_VIS_B] = (!PT || p.par.vis_final_b) ? p.par.vis_final_b[0] : 0.f;
    }
    return;
  }
  // ------------------------------------------------------------------ frame block, split over kFrameParts blocks
  const FrameLayout& F = p.fl;
  const int f = blockIdx.x / kFrameParts, part = blockIdx.x % kFrameParts;
  const int fn = (M >= 2) ? (f ^ 1) : f;
  float* fb = p.workspace + p.cl.n_floats + (size_t)f * F.n_floats;
  if (part == 0 && !(PT && p.skip_cams)) {
    write_cam(fb + F.cam, p, f);
    write_cam(fb + F.cam_partner, p, fn);
  }
  const float* codes[kNumCodes];
  codes[CODE_INST_BASE] = p.fr.inst_base ? p.fr.inst_base + (size_t)f * 32 : nullptr;
  codes[CODE_INST_COLOR] = p.fr.inst_color ? p.fr.inst_color + (size_t)f * 32 : nullptr;
  codes[CODE_INST_VIS] = p.fr.inst_vis ? p.fr.inst_vis + (size_t)f * 32 : nullptr;
  codes[CODE_APPR] = p.fr.appr_code ? p.fr.appr_code + (size_t)f * p.desc.appr_channels : nullptr;
  codes[CODE_INST_SKIN] = p.fr.inst_skin ? p.fr.inst_skin + (size_t)f * 32 : nullptr;
  codes[CODE_T_EMBED] = p.fr.skin_t_embed ? p.fr.skin_t_embed + (size_t)f * 128 : nullptr;
  codes[CODE_T_EMBED_MEAN] = p.fr.skin_t_embed_mean;
  codes[CODE_DENSE_T] = p.fr.dense_t_embed ? p.fr.dense_t_embed + (size_t)f * 128 : nullptr;
  codes[CODE_DENSE_T_PARTNER] = p.fr.dense_t_embed ? p.fr.dense_t_embed + (size_t)fn * 128 : nullptr;
  codes[CODE_INST_DENSE_FWD] = p.fr.inst_dense_fwd ? p.fr.inst_dense_fwd + (size_t)f * 32 : nullptr;
  codes[CODE_INST_DENSE_BWD] = p.fr.inst_dense_bwd ? p.fr.inst_dense_bwd + (size_t)f * 32 : nullptr;
  // all conditioned bias rows of the frame form one index space; this block takes every kFrameParts-th chunk of 256.
  // Point entries pass only the codes of the layers they evaluate: their rows are filtered.
  frame_rows<PT>(p, F, codes, fb, part);
  if (B > 0 && part >= 1 && !(PT && p.skip_bones)) {
    __shared__ float ig[32 * 4];
    for (int b = threadIdx.x; b < B; b += blockDim.x)
      for (int c = 0; c < 3; ++c) {
        float lg = p.par.log_gauss[b * 3 + c];
        if (p.par.symm_idx) lg = 0.5f * (p.par.log_gauss[p.par.symm_idx[b] * 3 + c] + lg);
        ig[b * 4 + c] = expf(-lg);
      }
    __syncthreads();
    const size_t o = (size_t)f * B * 4, on = (size_t)fn * B * 4;
    if (part == 1) {
      write_binv(fb + F.binv_t, p.fr.t_art_qr + o, p.fr.t_art_qd + o, ig, B);
      write_se3(fb + F.se3_bwd, p.fr.rest_art_qr + o, p.fr.rest_art_qd + o, p.fr.t_art_qr + o, p.fr.t_art_qd + o, B);
    } else if (part == 2) {
      write_binv(fb + F.binv_rest, p.fr.rest_art_qr + o, p.fr.rest_art_qd + o, ig, B);
      write_se3(fb + F.se3_fwd, p.fr.t_art_qr + o, p.fr.t_art_qd + o, p.fr.rest_art_qr + o, p.fr.rest_art_qd + o, B);
    } else {
      write_binv(fb + F.binv_rest_partner, p.fr.rest_art_qr + on, p.fr.rest_art_qd + on, ig, B);
      write_se3(fb + F.se3_fwd_partner, p.fr.t_art_qr + on, p.fr.t_art_qd + on, p.fr.rest_art_qr + on, p.fr.rest_art_qd + on, B);
    }
  }
}

cudaError_t launch_prologue(const PrologueParams& p, cudaStream_t stream) {
  if (p.skip_cams) prologue_kernel<true><<<kFrameParts * p.fr.M + 1, 256, 0, stream>>>(p);
  else prologue_kernel<false><<<kFrameParts * p.fr.M + 1, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace b200r
