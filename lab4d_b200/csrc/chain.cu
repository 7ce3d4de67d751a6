// Backward of the per-frame prologue (csrc/prologue.cu): chains the gradients of the kernel-level blocks - the constant
// block and the M frame blocks that csrc/field_bwd.cu and csrc/wgrad.cu fill - to
//   * the parameters: bias rows, the weight columns that multiply per-frame codes, head weights, scalars, Gaussian bone
//     scales (accumulated into the flat gradient buffer at the offsets of b200r_param_grads), and
//   * the per-frame inputs of query_field: codes, cameras, articulations (b200r_frame_grads).
// M x B rows of quaternion calculus and a few (M x 32) mat-vecs, hand-derived from the table formulas of prologue.cu
// (utils/transforms.py:9-25, nnutils/warping.py:304-314, nnutils/base.py:140-146); <0.1 % of the step, one launch.
// tests/test_gpu_backward.py checks it against autograd of a torch restatement of the tables (oracle/chain_torch.py).
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

// rows (R'_i0 R'_i1 R'_i2 t'_i) = ig_i * [R(q) | t], q = conj(qr), t = 2 vec(conj(qd) qr)  (prologue.cu write_binv):
// cotangent G (12 floats) -> g_qr, g_qd (added), g_ig (3, added)
__device__ void binv_bwd(const Q4& qr, const Q4& qd, const float* ig, const float* G, Q4& g_qr, Q4& g_qd, float* g_ig) {
  const Q4 q = qconj(qr);
  const Q4 tq = qmul(qconj(qd), qr);
  const float t[3] = {2.f * tq.x, 2.f * tq.y, 2.f * tq.z};
  const float ww = q.w * q.w, xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z;
  const float xy = q.x * q.y, xz = q.x * q.z, yz = q.y * q.z, wx = q.w * q.x, wy = q.w * q.y, wz = q.w * q.z;
  const float R[3][3] = {{ww + xx - yy - zz, 2.f * (xy - wz), 2.f * (xz + wy)},
                         {2.f * (xy + wz), ww - xx + yy - zz, 2.f * (yz - wx)},
                         {2.f * (xz - wy), 2.f * (yz + wx), ww - xx - yy + zz}};
  float gR[3][3], gt[3];
  for (int i = 0; i < 3; ++i) {
    float s = G[4 * i + 3] * t[i];
    for (int j = 0; j < 3; ++j) { gR[i][j] = ig[i] * G[4 * i + j]; s += G[4 * i + j] * R[i][j]; }
    gt[i] = ig[i] * G[4 * i + 3];
    g_ig[i] += s;
  }
  const float gw = 2.f * (q.w * (gR[0][0] + gR[1][1] + gR[2][2]) + q.z * (gR[1][0] - gR[0][1]) + q.y * (gR[0][2] - gR[2][0]) + q.x * (gR[2][1] - gR[1][2]));
  const float gx = 2.f * (q.x * (gR[0][0] - gR[1][1] - gR[2][2]) + q.y * (gR[0][1] + gR[1][0]) + q.z * (gR[0][2] + gR[2][0]) + q.w * (gR[2][1] - gR[1][2]));
  const float gy = 2.f * (q.y * (-gR[0][0] + gR[1][1] - gR[2][2]) + q.x * (gR[0][1] + gR[1][0]) + q.z * (gR[1][2] + gR[2][1]) + q.w * (gR[0][2] - gR[2][0]));
  const float gz = 2.f * (q.z * (-gR[0][0] - gR[1][1] + gR[2][2]) + q.x * (gR[0][2] + gR[2][0]) + q.y * (gR[1][2] + gR[2][1]) + q.w * (gR[1][0] - gR[0][1]));
  g_qr = qadd(g_qr, Q4{gw, -gx, -gy, -gz});  // q = conj(qr)
  // t = 2 vec(A qr), A = conj(qd):  g_A = g_r qr*,  g_qr += A* g_r
  const Q4 g_r = {0.f, 2.f * gt[0], 2.f * gt[1], 2.f * gt[2]};
  g_qd = qadd(g_qd, qconj(qmul(g_r, qconj(qr))));
  g_qr = qadd(g_qr, qmul(qd, g_r));  // conj(conj(qd)) = qd
}
// (real, dual) = a (x) b^-1: real = ar br*, dual = ar bd* + ad br*  (prologue.cu write_se3); cotangents added
__device__ void se3_bwd(const Q4& ar, const Q4& ad, const Q4& br, const Q4& bd, const Q4& g_re, const Q4& g_du, Q4& g_ar, Q4& g_ad, Q4& g_br, Q4& g_bd) {
  g_ar = qadd(g_ar, qadd(qmul(g_re, br), qmul(g_du, bd)));
  g_ad = qadd(g_ad, qmul(g_du, br));
  g_br = qadd(g_br, qconj(qadd(qmul(qconj(ar), g_re), qmul(qconj(ad), g_du))));
  g_bd = qadd(g_bd, qconj(qmul(qconj(ar), g_du)));
}
// r = vec(q (0,p) q*): cotangent g -> (g_q, g_p)
__device__ void qrot_bwd(const Q4& q, const float* p, const float* g, Q4& g_q, float* g_p) {
  const Q4 G = {0.f, g[0], g[1], g[2]}, Pq = {0.f, p[0], p[1], p[2]};
  const Q4 u = qmul(q, Pq), g_u = qmul(G, q);
  g_q = qadd(qmul(qconj(G), u), qmul(g_u, qconj(Pq)));
  const Q4 t = qmul(qconj(q), g_u);
  g_p[0] = t.x; g_p[1] = t.y; g_p[2] = t.z;
}

__device__ __forceinline__ const float* code_ptr(const ChainParams& p, int cid, int f, int fn) {
  switch (cid) {
    case CODE_INST_BASE: return p.fr.inst_base ? p.fr.inst_base + (size_t)f * 32 : nullptr;
    case CODE_INST_COLOR: return p.fr.inst_color ? p.fr.inst_color + (size_t)f * 32 : nullptr;
    case CODE_INST_VIS: return p.fr.inst_vis ? p.fr.inst_vis + (size_t)f * 32 : nullptr;
    case CODE_APPR: return p.fr.appr_code ? p.fr.appr_code + (size_t)f * p.desc.appr_channels : nullptr;
    case CODE_INST_SKIN: return p.fr.inst_skin ? p.fr.inst_skin + (size_t)f * 32 : nullptr;
    case CODE_T_EMBED: return p.fr.skin_t_embed ? p.fr.skin_t_embed + (size_t)f * 128 : nullptr;
    case CODE_T_EMBED_MEAN: return p.fr.skin_t_embed_mean;
    case CODE_DENSE_T: return p.fr.dense_t_embed ? p.fr.dense_t_embed + (size_t)f * 128 : nullptr;
    case CODE_DENSE_T_PARTNER: return p.fr.dense_t_embed ? p.fr.dense_t_embed + (size_t)fn * 128 : nullptr;
    case CODE_INST_DENSE_FWD: return p.fr.inst_dense_fwd ? p.fr.inst_dense_fwd + (size_t)f * 32 : nullptr;
    case CODE_INST_DENSE_BWD: return p.fr.inst_dense_bwd ? p.fr.inst_dense_bwd + (size_t)f * 32 : nullptr;
  }
  return nullptr;
}
__device__ __forceinline__ float* gcode_ptr(const ChainParams& p, int cid, int f, int fn) {
  switch (cid) {
    case CODE_INST_BASE: return p.gf.inst_base ? p.gf.inst_base + (size_t)f * 32 : nullptr;
    case CODE_INST_COLOR: return p.gf.inst_color ? p.gf.inst_color + (size_t)f * 32 : nullptr;
    case CODE_INST_VIS: return p.gf.inst_vis ? p.gf.inst_vis + (size_t)f * 32 : nullptr;
    case CODE_APPR: return p.gf.appr_code ? p.gf.appr_code + (size_t)f * p.desc.appr_channels : nullptr;
    case CODE_INST_SKIN: return p.gf.inst_skin ? p.gf.inst_skin + (size_t)f * 32 : nullptr;
    case CODE_T_EMBED: return p.gf.skin_t_embed ? p.gf.skin_t_embed + (size_t)f * 128 : nullptr;
    case CODE_T_EMBED_MEAN: return p.gf.skin_t_embed_mean;
    case CODE_DENSE_T: return p.gf.dense_t_embed ? p.gf.dense_t_embed + (size_t)f * 128 : nullptr;
    case CODE_DENSE_T_PARTNER: return p.gf.dense_t_embed ? p.gf.dense_t_embed + (size_t)fn * 128 : nullptr;
    case CODE_INST_DENSE_FWD: return p.gf.inst_dense_fwd ? p.gf.inst_dense_fwd + (size_t)f * 32 : nullptr;
    case CODE_INST_DENSE_BWD: return p.gf.inst_dense_bwd ? p.gf.inst_dense_bwd + (size_t)f * 32 : nullptr;
  }
  return nullptr;
}

// Block roles, by blockIdx.x:
//   [0, M * n_cond)                 (frame f, bias row c): dL/d code[f] = G[f] W[:, code columns]
//   next sum_c n_c blocks           (bias row c, output n): dL/d bias[n] and dL/d W[n, code columns] = sum_f G[f][n] code[f]
//   next 1 block                    constant block -> flat buffer (plain biases, heads, scalars)
//   next ceil(M / 128) blocks       cameras, one thread per frame
//   next M blocks                   bone tables, one thread per (frame, bone)
__global__ void __launch_bounds__(256) chain_kernel(const __grid_constant__ ChainParams p) {
  const FrameLayout& F = p.fl;
  const ConstLayout& C = p.cl;
  const int M = p.fr.M, B = p.desc.n_bones;
  int bid = blockIdx.x;
  const size_t FF = (size_t)F.n_floats;
  // ------------------------------------------------------------------ dL/d code rows
  if (bid < M * F.n_cond) {
    const int f = bid / F.n_cond, ci = bid % F.n_cond, fn = M >= 2 ? (f ^ 1) : f;
    const CondRow& c = F.cond[ci];
    __shared__ float G[256];
    for (int n = threadIdx.x; n < c.n; n += blockDim.x) G[n] = p.g_fblk[f * FF + c.frame_off + n];
    __syncthreads();
    const float* W = p.par.weight[c.layer];
    for (int sgi = 0; sgi < c.n_seg; ++sgi) {
      float* gc = gcode_ptr(p, c.code[sgi], f, fn);
      if (!gc) continue;
      for (int j = threadIdx.x; j < c.width[sgi]; j += blockDim.x) {
        float a = 0.f;
        const float* wc = W + c.col0[sgi] + j;
        for (int n = 0; n < c.n; ++n) a += G[n] * wc[(size_t)n * c.in_dim];
        atomicAdd(gc + j, a);  // several rows / frames share a code (mean and partner codes, both delta_field.linear_1 rows)
      }
    }
    return;
  }
  bid -= M * F.n_cond;
  // ------------------------------------------------------------------ dL/d bias and dL/d W[:, code columns]
  int rows_total = 0;
  for (int ci = 0; ci < F.n_cond; ++ci) rows_total += F.cond[ci].n;
  if (bid < rows_total) {
    int ci = 0, n = bid;
    while (n >= F.cond[ci].n) { n -= F.cond[ci].n; ++ci; }
    const CondRow& c = F.cond[ci];
    __shared__ float Gf[1024];  // G[f][n] over frames (M <= 1024 per pass)
    float bsum = 0.f;
    for (int f0 = 0; f0 < M; f0 += 1024) {
      const int mf = M - f0 < 1024 ? M - f0 : 1024;
      __syncthreads();
      for (int f = threadIdx.x; f < mf; f += blockDim.x) Gf[f] = p.g_fblk[(f0 + f) * FF + c.frame_off + n];
      __syncthreads();
      if (threadIdx.x == 0)
        for (int f = 0; f < mf; ++f) bsum += Gf[f];
      for (int sgi = 0; sgi < c.n_seg; ++sgi) {
        for (int j = threadIdx.x; j < c.width[sgi]; j += blockDim.x) {
          float a = 0.f;
          for (int f = 0; f < mf; ++f) {
            const int ff = f0 + f, fn = M >= 2 ? (ff ^ 1) : ff;
            const float* code = code_ptr(p, c.code[sgi], ff, fn);
            if (code) a += Gf[f] * code[j];
          }
          atomicAdd(p.grad + p.off.weight_off[c.layer] + (size_t)n * c.in_dim + c.col0[sgi] + j, a);
        }
      }
    }
    if (threadIdx.x == 0) atomicAdd(p.grad + p.off.bias_off[c.layer] + n, bsum);
    return;
  }
  bid -= rows_total;
  // ------------------------------------------------------------------ constant block -> flat buffer
  if (bid == 0) {
    const int W = p.desc.W, H = W / 2;
    for (int l = 0; l < p.n_layers; ++l) {
      if (C.plain_off[l] < 0 || p.off.bias_off[l] < 0) continue;
      for (int i = threadIdx.x; i < p.layer_out[l]; i += blockDim.x) p.grad[p.off.bias_off[l] + i] += p.g_cblk[C.plain_off[l] + i];
    }
    for (int i = threadIdx.x; i < W; i += blockDim.x) p.grad[p.off.sdf_w + i] += p.g_cblk[C.sdf_w + i];
    for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) p.grad[p.off.rgb2_w + i] += p.g_cblk[C.rgb2_w + i];
    if (p.off.vis_final_w >= 0)
      for (int i = threadIdx.x; i < 64; i += blockDim.x) p.grad[p.off.vis_final_w + i] += p.g_cblk[C.vis_w + i];
    if (threadIdx.x == 0) {
      const float* s = p.g_cblk + C.scalars;
      p.grad[p.off.logibeta] += s[SC_IBETA];
      p.grad[p.off.logscale] += s[SC_INV_SCALE];
      if (p.off.warp_logibeta >= 0) p.grad[p.off.warp_logibeta] += s[SC_WARP_IBETA];
      p.grad[p.off.sdf_b] += s[SC_SDF_B];
      for (int i = 0; i < 3; ++i) p.grad[p.off.rgb2_b + i] += s[SC_RGB2_B0 + i];
      if (p.off.vis_final_b >= 0) p.grad[p.off.vis_final_b] += s[SC_VIS_B];
    }
    return;
  }
  bid -= 1;
  // ------------------------------------------------------------------ cameras
  const int cam_blocks = (M + 127) / 128;
  if (bid < cam_blocks) {
    const int f = bid * 128 + threadIdx.x;
    if (threadIdx.x >= 128 || f >= M) return;
    const int fn = M >= 2 ? (f ^ 1) : f;
    const float* gc = p.g_fblk + f * FF + F.cam;
    const float* gp = p.g_fblk + fn * FF + F.cam_partner;  // the partner's block holds this frame's camera
    if (p.gf.Kinv)
      for (int i = 0; i < 9; ++i) p.gf.Kinv[(size_t)f * 9 + i] = gc[i] + gp[i];
    if (!p.fr.field2cam_q || !p.fr.field2cam_t) return;  // point entries (b200r_warp_bwd) have no cameras
    // qi = conj(q), ti = R(qi) (-t): cotangents (g_qi, g_ti) -> (g_q, g_t)
    const Q4 q = ld4(p.fr.field2cam_q + (size_t)f * 4), qi = qconj(q);
    const float mt[3] = {-p.fr.field2cam_t[f * 3], -p.fr.field2cam_t[f * 3 + 1], -p.fr.field2cam_t[f * 3 + 2]};
    Q4 g_qi2;
    float g_mt[3];
    qrot_bwd(qi, mt, gc + 15, g_qi2, g_mt);
    const Q4 g_qi = qadd(ld4(gc + 11), g_qi2);
    if (p.gf.field2cam_q) st4(p.gf.field2cam_q + (size_t)f * 4, qadd(qconj(g_qi), ld4(gp + 11)));
    if (p.gf.field2cam_t)
      for (int i = 0; i < 3; ++i) p.gf.field2cam_t[(size_t)f * 3 + i] = -g_mt[i] + gp[15 + i];
    return;
  }
  bid -= cam_blocks;
  // ------------------------------------------------------------------ bone tables
  if (bid < M && B > 0) {
    const int f = bid, fn = M >= 2 ? (f ^ 1) : f, b = threadIdx.x;
    if (b >= B) return;
    const size_t o = ((size_t)f * B + b) * 4;
    const Q4 tqr = ld4(p.fr.t_art_qr + o), tqd = ld4(p.fr.t_art_qd + o), rqr = ld4(p.fr.rest_art_qr + o), rqd = ld4(p.fr.rest_art_qd + o);
    float ig[3], g_ig[3] = {0.f, 0.f, 0.f};
    for (int c = 0; c < 3; ++c) {
      float lg = p.par.log_gauss[b * 3 + c];
      if (p.par.symm_idx) lg = 0.5f * (p.par.log_gauss[p.par.symm_idx[b] * 3 + c] + lg);
      ig[c] = expf(-lg);
    }
    Q4 g_tqr = {0, 0, 0, 0}, g_tqd = {0, 0, 0, 0}, g_rqr = {0, 0, 0, 0}, g_rqd = {0, 0, 0, 0};
    const float* gf = p.g_fblk + f * FF;
    const float* gfn = p.g_fblk + fn * FF;  // this frame's rest tables also live in its partner's block
    binv_bwd(tqr, tqd, ig, gf + F.binv_t + 12 * b, g_tqr, g_tqd, g_ig);
    float Gr[12];
    for (int i = 0; i < 12; ++i) Gr[i] = gf[F.binv_rest + 12 * b + i] + gfn[F.binv_rest_partner + 12 * b + i];
    binv_bwd(rqr, rqd, ig, Gr, g_rqr, g_rqd, g_ig);
    // se3_bwd = rest (x) t^-1 ; se3_fwd = t (x) rest^-1 (own block + partner's block)
    se3_bwd(rqr, rqd, tqr, tqd, ld4(gf + F.se3_bwd + 8 * b), ld4(gf + F.se3_bwd + 8 * b + 4), g_rqr, g_rqd, g_tqr, g_tqd);
    const Q4 g_re = qadd(ld4(gf + F.se3_fwd + 8 * b), ld4(gfn + F.se3_fwd_partner + 8 * b));
    const Q4 g_du = qadd(ld4(gf + F.se3_fwd + 8 * b + 4), ld4(gfn + F.se3_fwd_partner + 8 * b + 4));
    se3_bwd(tqr, tqd, rqr, rqd, g_re, g_du, g_tqr, g_tqd, g_rqr, g_rqd);
    if (f == 0) {  // rest bone centres of the Gaussian bone density: 2 vec(rqd rqr*)
      const float* gc = p.g_cblk + C.center + 4 * b;
      const Q4 g_r = {0.f, 2.f * gc[0], 2.f * gc[1], 2.f * gc[2]};
      g_rqd = qadd(g_rqd, qmul(g_r, rqr));
      g_rqr = qadd(g_rqr, qconj(qmul(qconj(rqd), g_r)));
    }
    if (p.gf.t_art_qr) st4(p.gf.t_art_qr + o, g_tqr);
    if (p.gf.t_art_qd) st4(p.gf.t_art_qd + o, g_tqd);
    if (p.gf.rest_art_qr) st4(p.gf.rest_art_qr + o, g_rqr);
    if (p.gf.rest_art_qd) st4(p.gf.rest_art_qd + o, g_rqd);
    if (p.off.log_gauss >= 0) {  // ig = exp(-lgs), lgs = (lg[symm] + lg) / 2
      for (int c = 0; c < 3; ++c) {
        const float g_lgs = -ig[c] * g_ig[c];
        if (p.par.symm_idx) {
          atomicAdd(p.grad + p.off.log_gauss + b * 3 + c, 0.5f * g_lgs);
          atomicAdd(p.grad + p.off.log_gauss + p.par.symm_idx[b] * 3 + c, 0.5f * g_lgs);
        } else {
          atomicAdd(p.grad + p.off.log_gauss + b * 3 + c, g_lgs);
        }
      }
    }
  }
}

cudaError_t launch_chain(const ChainParams& p, cudaStream_t stream) {
  int rows_total = 0;
  for (int ci = 0; ci < p.fl.n_cond; ++ci) rows_total += p.fl.cond[ci].n;
  const int M = p.fr.M;
  const int grid = M * p.fl.n_cond + rows_total + 1 + (M + 127) / 128 + (p.desc.n_bones > 0 ? M : 0);
  chain_kernel<<<grid, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace b200r
