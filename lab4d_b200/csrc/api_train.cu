// C ABI of the training path: tape sizes, block layouts, b200r_field_bwd (declarations: include/b200r.h).
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "api_util.h"

namespace b200r {

// ---------------------------------------------------------------- gradient scale
// max |g| over the cotangent arrays (the density's is weighted by |d density / d sdf| <= ibeta^2 / 2, the factor its
// gradient picks up first) -> power-of-two scale that puts the largest entry at ~1024 in the 16-bit gradient operands
// (entries 1e-7 of the largest still land in fp16's normal range).
struct ScaleParams {
  const float* ptr[12];
  long long n[12];
  float weight[12];
  const float* logibeta;
  float* scale;  // [0] scale, [1] 1 / scale
  unsigned int* amax_bits;
  int bf16;
  float inv_extra;  // scale[1] = inv_extra / scale (0 = 1): the eikonal backward folds its reverse chain's fixed scale in
};
__global__ void absmax_kernel(const ScaleParams p) {
  float m = 0.f;
  for (int a = 0; a < 12; ++a) {
    if (!p.ptr[a]) continue;
    float w = p.weight[a];
    if (w < 0.f) { const float ib = expf(p.logibeta[0]); w = 0.5f * ib * ib; }
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.n[a]; i += (long long)gridDim.x * blockDim.x) {
      const float v = fabsf(p.ptr[a][i]) * w;
      if (v < INFINITY) m = fmaxf(m, v);
    }
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(p.amax_bits, __float_as_uint(m));
}
__global__ void scale_kernel(const ScaleParams p) {
  const float amax = __uint_as_float(*p.amax_bits);
  float s = 1.0f;
  if (!p.bf16 && amax > 0.f) s = exp2f(floorf(log2f(1024.0f / amax)));  // fp16 tops out at 65504: 64x headroom, saturating packs beyond
  s = fminf(fmaxf(s, 1.0f / 16777216.0f), 1099511627776.0f);
  p.scale[0] = s;
  p.scale[1] = (p.inv_extra != 0.f ? p.inv_extra : 1.0f) / s;
}

// ---------------------------------------------------------------- weight-gradient job list
enum : int { DST_WEIGHTS = 0, DST_FRAME = 1, DST_CONST = 2 };

// only_w >= 0: the jobs of that skinning warp alone (b200r_warp_bwd)
static std::vector<WgradJob> build_jobs(const b200r_field_desc& d, const BuiltProgram& bp, const TapeLayout& T, const int64_t* woff, int only_w = -1) {
  std::vector<WgradJob> jobs;
  const Program& P = bp.prog;
  const LayerIds L = layer_ids(d);
  const int B = d.n_bones, W = d.W, KC = W / 64, HN = W / 2;
  const int pe_b = pe_dim(d.L_xyz), pe_c = pe_dim(d.L_xyz + 2), pe_v = pe_dim(10), pe_f = pe_dim(6);
  std::vector<char> summed(512, 0);  // G operand (by first chunk id) whose columns were already summed
  auto view = [](int row0, int col0, int rows, int cols, int ld, int kind, int64_t off) {
    WgradView v;
    v.row0 = row0; v.col0 = col0; v.rows = rows; v.cols = cols; v.ld = ld; v.per_frame = kind; v.dst_off = off;
    return v;
  };
  // G (gradient tape) x A (operand from tape `a_src`): one view
  auto job = [&](int g_chunk, int n_g, int a_chunk, int n_a, int a_src, const WgradView& v0) -> WgradJob& {
    WgradJob j;
    memset(&j, 0, sizeof(j));
    j.g_chunk = (int16_t)g_chunk; j.n_g = (int16_t)n_g; j.a_chunk = (int16_t)a_chunk; j.n_a = (int16_t)n_a;
    j.g_src = 1; j.a_src = (uint8_t)a_src;
    j.n_views = 1; j.v[0] = v0;
    j.per_frame = v0.per_frame == DST_FRAME;
    jobs.push_back(j);
    return jobs.back();
  };
  // weight gradient of `layer`, in-columns [c0, c0 + cols), and (once per G operand) its bias gradient
  auto layer_job = [&](int layer, int g_chunk, int a_chunk, int n_a, int c0, int cols, int bias_frame_off = -1, int a_src = 0) {
    const int n_out = bp.layer_out[layer], n_g = (n_out + 63) / 64;
    WgradJob& j = job(g_chunk, n_g, a_chunk, n_a, a_src, view(0, 0, n_out, cols, bp.layer_in[layer], DST_WEIGHTS, woff[layer] + c0));
    if (!summed[g_chunk]) {
      summed[g_chunk] = 1;
      const bool frame = bias_frame_off >= 0 || P.bias[layer].frame;
      j.colsum = frame ? 2 : 1;
      j.colsum_off = bias_frame_off >= 0 ? bias_frame_off : P.bias[layer].off;
      j.colsum_n = n_out;
    }
  };
  const int n_pe = pe_b < 63 ? pe_b : 63;
  for (int w = 0; w < 3 && B > 0; ++w) {
    if (only_w >= 0 && w != only_w) continue;
    layer_job(L.delta[0], T.g_z1[w], T.a_xb[w], (3 * B + 63) / 64, 0, 3 * B, w == 0 ? -1 : P.fl.delta1_fwd);
    layer_job(L.delta[1], T.g_z2[w], T.a_h1[w], 1, 0, 64);
    layer_job(L.delta[2], T.g_z[w], T.a_h2[w], 1, 0, 64);
    // bone tables of the frame: [g_xb | ws]^T [x y z 1 | g_qhr g_qhd]
    const int binv = w == 0 ? P.fl.binv_t : (w == 1 ? P.fl.binv_rest_partner : P.fl.binv_rest);
    const int se3 = w == 0 ? P.fl.se3_bwd : (w == 1 ? P.fl.se3_fwd_partner : P.fl.se3_fwd);
    WgradJob& j = job(T.g_xbw[w], 2, T.g_xg[w], 1, 1, view(0, 0, 3 * B, 4, 4, DST_FRAME, binv));
    j.n_views = 2;
    j.v[1] = view(96, 4, B, 8, 8, DST_FRAME, se3);
    if (d.dense) {  // soft deformation of this stage: backward_map for w = 0, forward_map (partner / own time code) for w = 1, 2
      const int m = w == 0 ? 1 : 0;
      const int row1 = w == 1 ? P.fl.dense1_partner : -1;  // bias row with the partner frame's time code
      layer_job(L.dense[3 * m + 0], T.g_d1[w], T.a_dpe[w], 1, 0, pe_dim(6), row1);
      layer_job(L.dense[3 * m + 1], T.g_d2[w], T.a_dh1[w], 4, 0, 256);
      layer_job(L.dense[3 * m + 2], T.g_d3[w], T.a_dh2[w], 4, 0, 256);
    }
  }
  if (only_w >= 0) return jobs;
  layer_job(L.vis[0], T.g_vis[0], T.a_pe, 1, 0, pe_v);
  layer_job(L.vis[1], T.g_vis[1], T.a_vis[0], 1, 0, 64);
  // heads: rows of the head chunk x their input operand
  {
    WgradJob& j = job(T.g_head, 1, T.a_vis[1], 1, 0, view(7, 0, 1, 64, 64, DST_CONST, P.cl.vis_w));
    j.colsum = 1; j.colsum_off = P.cl.scalars; j.colsum_n = 8;  // sdf.bias, rgb.2.bias, vis final bias
    job(T.g_head, 1, T.a_base[d.D], KC, 0, view(3, 0, 1, W, W, DST_CONST, P.cl.sdf_w));
    job(T.g_head, 1, T.a_rgb0, KC / 2, 0, view(4, 0, 3, HN, HN, DST_CONST, P.cl.rgb2_w));
  }
  for (int i = 0; i <= d.D; ++i) {
    if (i == 0) {
      layer_job(L.base[i], T.g_base[i], T.a_pe, 1, 0, n_pe);
    } else if (i == d.skip) {
      layer_job(L.base[i], T.g_base[i], T.a_pe, 1, 0, n_pe);
      layer_job(L.base[i], T.g_base[i], T.a_base[i - 1], KC, pe_b + 32, W);
    } else {
      layer_job(L.base[i], T.g_base[i], T.a_base[i - 1], KC, 0, W);
    }
  }
  layer_job(L.rgb0, T.g_rgb0, T.a_f2, KC, 0, W);
  if (d.L_dir == 0) layer_job(L.rgb0, T.g_rgb0, T.a_dir, 1, W, 3);
  layer_job(L.color[0], T.g_col[0], T.a_pe, 1, 0, pe_c < 63 ? pe_c : 63);
  if (pe_c > 63) layer_job(L.color[0], T.g_col[0], T.a_extra, 1, 63, pe_c - 63);
  layer_job(L.color[1], T.g_col[1], T.a_col[0], KC, 0, W);
  layer_job(L.color[2], T.g_col[2], T.a_col[1], KC, 0, W);
  if (d.has_feature) {
    layer_job(L.feat[0], T.g_feat[0], T.a_pe, 1, 0, pe_f);
    for (int i = 1; i < 4; ++i) layer_job(L.feat[i], T.g_feat[i], T.a_feat[i - 1], 2, 0, 128);
    layer_job(L.feat[4], T.g_feat[4], T.a_pe, 1, 0, pe_f);
    layer_job(L.feat[4], T.g_feat[4], T.a_feat[3], 2, pe_f, 128);
    layer_job(L.feat[5], T.g_feat[5], T.a_feat[4], 2, 0, 128);
  }
  return jobs;
}

// Weight-gradient jobs of the eikonal term (oracle/eikonal_backward.py weight_grads): G = reverse-chain tape (a_i), A = forward-chain
// tape (v_0, chain A, chain B); dW_i = a_i^T (vA_{i-1} + vB_{i-1}), the embedding columns of linear_1 / the skip layer from v_0,
// d sdf.weight = head^T (vA_D + vB_D).  No bias column sums: the chains carry no bias.
static std::vector<WgradJob> build_eik_jobs(const b200r_field_desc& d, const BuiltProgram& bp, const EikLayout& E, const int64_t* woff, int64_t sdf_off) {
  std::vector<WgradJob> jobs;
  const LayerIds L = layer_ids(d);
  const int W = d.W, KC = W / 64, pe_b = pe_dim(d.L_xyz), n_pe = pe_b < 63 ? pe_b : 63;
  auto job = [&](int g_chunk, int n_g, int a_chunk, int n_a, int row0, int rows, int cols, int ld, int64_t off) {
    WgradJob j;
    memset(&j, 0, sizeof(j));
    j.g_chunk = (int16_t)g_chunk; j.n_g = (int16_t)n_g; j.a_chunk = (int16_t)a_chunk; j.n_a = (int16_t)n_a;
    j.g_src = 1; j.a_src = 0;
    j.n_views = 1;
    j.v[0].row0 = row0; j.v[0].col0 = 0; j.v[0].rows = rows; j.v[0].cols = cols; j.v[0].ld = ld; j.v[0].per_frame = DST_WEIGHTS; j.v[0].dst_off = off;
    jobs.push_back(j);
  };
  for (int i = 0; i <= d.D; ++i) {
    const int layer = L.base[i], ld = bp.layer_in[layer];
    const int64_t off = woff[layer];
    if (i == 0) {
      job(E.a_base[i], KC, E.v0, 1, 0, W, n_pe, ld, off);
    } else if (i == d.skip) {
      job(E.a_base[i], KC, E.v0, 1, 0, W, n_pe, ld, off);
      job(E.a_base[i], KC, E.vA[i - 1], KC, 0, W, W, ld, off + pe_b + 32);
    } else {
      job(E.a_base[i], KC, E.vA[i - 1], KC, 0, W, W, ld, off);
      if (E.vB[i - 1] >= 0) job(E.a_base[i], KC, E.vB[i - 1], KC, 0, W, W, ld, off);
    }
  }
  job(E.a_head, 1, E.vA[d.D], KC, 3, 1, W, W, sdf_off);
  job(E.a_head, 1, E.vB[d.D], KC, 3, 1, W, W, sdf_off);
  return jobs;
}

// Split the (job, tile) line over `grid` CTAs so that every CTA gets the same estimated TIME.  Model (B200, measured shares):
// a half-tile stage is latency-bound below ~3 chunks (1.2 us per tile), bandwidth-bound above (0.36 us per 16-KB chunk at
// the SM's share of HBM); every work item ends with an accumulator flush of rows x cols fp32 atomics (~35 us per 64 K).
static double model_const(const char* name, double dflt) {  // tuning hook: B200R_WG_LAT / _BW / _FLUSH (microseconds)
  const char* e = getenv(name);
  return e ? atof(e) : dflt;
}
static double job_tile_us(const WgradJob& j) {
  static const double lat = model_const("B200R_WG_LAT", 2.5), per_chunk = model_const("B200R_WG_BW", 0.36);
  const double bw = per_chunk * (j.n_g + j.n_a);
  return bw > lat ? bw : lat;
}
static double job_flush_us(const WgradJob& j) {
  static const double fl = model_const("B200R_WG_FLUSH", 50.0);
  double area = 0;
  for (int v = 0; v < j.n_views; ++v) area += (double)j.v[v].rows * ((j.v[v].cols + 31) / 32 * 32);
  return 2.0 + fl * area / 65536.0;
}
// greedy assignment with a per-CTA time budget `per`; returns the largest load any CTA ends up with
static double assign_work(const std::vector<WgradJob>& jobs, int n_tiles, int tiles_per_frame, int grid, double per, std::vector<WgradWork>& work,
                          std::vector<int32_t>& first) {
  work.clear();
  first.assign(grid + 1, 0);
  int cta = 0;
  double used = 0, worst = 0;
  for (int ji = 0; ji < (int)jobs.size(); ++ji) {
    const WgradJob& j = jobs[ji];
    const double tt = job_tile_us(j) + (j.per_frame ? job_flush_us(j) / tiles_per_frame : 0.0);
    const double fl = j.per_frame ? 0.0 : job_flush_us(j);
    int t = 0;
    while (t < n_tiles) {
      const double room = per - used - fl;
      if (room < tt * 4 && cta + 1 < grid) {  // not worth a flush for a handful of tiles: next CTA
        ++cta;
        first[cta] = (int32_t)work.size();
        used = 0;
        continue;
      }
      long long take = (long long)(room / tt);
      if (take < 1) take = 1;
      if (cta + 1 == grid || take > n_tiles - t) take = n_tiles - t;
      work.push_back({ji, t, t + (int)take});
      used += take * tt + fl;
      worst = used > worst ? used : worst;
      t += (int)take;
    }
  }
  for (int c2 = cta + 1; c2 <= grid; ++c2) first[c2] = (int32_t)work.size();
  return worst;
}
static void build_work(const std::vector<WgradJob>& jobs, int n_tiles, int tiles_per_frame, int grid, std::vector<WgradWork>& work,
                       std::vector<int32_t>& first) {
  double total = 0;
  for (const auto& j : jobs) {
    const double frames = j.per_frame ? (double)n_tiles / tiles_per_frame : 1.0;
    total += job_tile_us(j) * n_tiles + job_flush_us(j) * (frames > 1 ? frames : 1.0);
  }
  // every cut adds a flush that `total` does not know about: smallest budget for which no CTA (the last one takes
  // whatever is left) exceeds it
  double lo = total / grid, hi = 2.0 * total / grid + 100.0;
  for (int it = 0; it < 24; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (assign_work(jobs, n_tiles, tiles_per_frame, grid, mid, work, first) <= mid * 1.0001) hi = mid;
    else lo = mid;
  }
  assign_work(jobs, n_tiles, tiles_per_frame, grid, hi, work, first);
}

}  // namespace b200r

extern "C" {

int b200r_tape_sizes(const b200r_field_desc* desc, int32_t M, int32_t N, int32_t D, size_t* a_bytes, size_t* g_bytes, size_t* mask_bytes) {
  if (!desc || M < 1 || N < 1 || D < 1) return B200R_E_INVALID;
  const b200r::TapeLayout T = b200r::tape_layout(*desc);
  const int tpf = (N * D + b200r::kTileRows - 1) / b200r::kTileRows, n_tiles = M * tpf;
  if (a_bytes) *a_bytes = b200r::tape_a_bytes(T, n_tiles);
  if (g_bytes) *g_bytes = b200r::tape_g_bytes(T, n_tiles + b200r::kMaxCtas);  // + scratch tiles of dead tile-pair halves
  if (mask_bytes) *mask_bytes = b200r::tape_mask_bytes(T, n_tiles);
  return B200R_OK;
}

int b200r_program_steps(const b200r_field_desc* desc, int32_t kind) {
  if (!desc) return B200R_E_INVALID;
  b200r_field_desc d = *desc;
  b200r_field_desc d16 = d;
  if (d16.operand_dtype == 2) d16.operand_dtype = 0;  // the backward-side programs run single 16-bit operands
  b200r::BuiltProgram bp;
  switch (kind) {
    case 0: bp = b200r::build_program(d, b200r::MODE_FIELD); break;
    case 1: bp = b200r::build_bwd_program(d16); break;
    case 2: bp = b200r::build_bwd_program(d16, true); break;
    case 3: bp = b200r::build_bwd_program(d16, false, 2); break;
    case 4: bp = b200r::build_bwd_program(d16, true, 0); break;
    case 5: bp = b200r::build_eik_chain_program(d, b200r::EIK_CHAIN_A); break;
    case 6: bp = b200r::build_eik_chain_program(d, b200r::EIK_CHAIN_B); break;
    case 7: bp = b200r::build_bwd_program(d16, false, 0); break;
    default: return B200R_E_INVALID;
  }
  return bp.ok ? bp.prog.n_steps : B200R_E_INVALID;
}

int b200r_get_block_layout(const b200r_field_desc* desc, b200r_block_layout* out) {
  if (!desc || !out) return B200R_E_INVALID;
  b200r::BuiltProgram bp = b200r::build_program(*desc);
  if (!bp.ok) return B200R_E_INVALID;
  const b200r::ConstLayout& C = bp.prog.cl;
  const b200r::FrameLayout& F = bp.prog.fl;
  memset(out, 0, sizeof(*out));
  out->const_floats = C.n_floats;
  out->frame_floats = F.n_floats;
  for (int i = 0; i < B200R_MAX_LAYERS; ++i) out->c_plain_bias[i] = C.plain_off[i];
  out->c_sdf_w = C.sdf_w; out->c_rgb2_w = C.rgb2_w; out->c_vis_w = C.vis_w; out->c_dir_w = desc->L_dir == 0 ? C.dir_w : -1;
  out->c_center = desc->n_bones > 0 ? C.center : -1;
  out->c_scalars = C.scalars;
  out->f_cam = F.cam; out->f_cam_partner = F.cam_partner;
  const bool sk = desc->n_bones > 0;
  out->f_binv_t = sk ? F.binv_t : -1; out->f_se3_bwd = sk ? F.se3_bwd : -1; out->f_binv_rest = sk ? F.binv_rest : -1;
  out->f_se3_fwd = sk ? F.se3_fwd : -1; out->f_binv_rest_partner = sk ? F.binv_rest_partner : -1;
  out->f_se3_fwd_partner = sk ? F.se3_fwd_partner : -1;
  out->n_cond = F.n_cond;
  for (int i = 0; i < F.n_cond && i < B200R_MAX_COND; ++i) {
    const b200r::CondRow& c = F.cond[i];
    out->cond[i].layer = c.layer; out->cond[i].n = c.n; out->cond[i].in_dim = c.in_dim; out->cond[i].frame_off = c.frame_off;
    out->cond[i].n_seg = c.n_seg;
    for (int s = 0; s < 2; ++s) { out->cond[i].col0[s] = c.col0[s]; out->cond[i].width[s] = c.width[s]; out->cond[i].code[s] = c.code[s]; }
  }
  return B200R_OK;
}

// shared body of b200r_field_bwd (rays) and b200r_warp_bwd (pts: one forward skinning warp of given points, cotangent g_points)
static int run_field_bwd(b200r_handle* h, const b200r_field_desc* desc, const void* packed_t, const b200r_field_params* par,
                         const b200r_frame_tables* fr, const b200r_ray_batch* rays_in, const b200r_point_batch* pts, const float* g_points,
                         float* g_points_out, const b200r_field_outputs* saved, const b200r_field_grads* grads, const b200r_tape* tape,
                         const b200r_param_grads* out, const b200r_frame_grads* fgr, void* workspace, size_t workspace_bytes, b200r_stream stream_) {
  if (!h) return B200R_E_INVALID;
  const char* who = pts ? "warp_bwd: " : "field_bwd: ";
  auto bad = [&](const char* msg) { return fail(h, B200R_E_INVALID, std::string(who) + msg); };
  if (!desc || !packed_t || !par || !fr || (!rays_in && !pts) || !saved || !grads || !tape || !out || !workspace) return bad("null argument");
  b200r_field_desc dsc = *desc;
  if (dsc.operand_dtype == 2) dsc.operand_dtype = 0;  // gradients run on single fp16 operands (scaled), whatever the forward used
  b200r::BuiltProgram bp = b200r::build_bwd_program(dsc, false, pts ? 2 : -1);
  if (!bp.ok) return bad(bp.err);
  b200r_ray_batch rays_pts;
  memset(&rays_pts, 0, sizeof(rays_pts));
  if (pts) { rays_pts.N = pts->P; rays_pts.D = 1; rays_pts.flow_thresh = -1.f; }
  const b200r_ray_batch* rays = pts ? &rays_pts : rays_in;
  const int M = fr->M, N = rays->N, D = rays->D;
  if (pts) {
    if (desc->n_bones <= 0) return bad("needs a skinned field");
    if (M < 1 || N < 1 || !pts->xyz || !g_points || !g_points_out) return bad("missing points or their cotangent");
    if (!fr->inst_skin || !fr->skin_t_embed || !fr->skin_t_embed_mean || !fr->t_art_qr || !fr->t_art_qd || !fr->rest_art_qr || !fr->rest_art_qd)
      return bad("missing skinning input");
  } else {
    if (M < 1 || N < 1 || D < 2) return bad("need M,N >= 1 and D >= 2");
    if (M >= 2 && (M & 1)) return bad("frames must come in adjacent pairs (M even)");
    if (!rays->hxy || !fr->Kinv || !fr->near_far || !fr->field2cam_q || !fr->field2cam_t) return bad("missing ray/camera input");
    if (!saved->xyz || !saved->rgb || !saved->sdf || (desc->has_feature && (!saved->feature || !saved->feat_norm))) return bad("missing saved forward outputs");
  }
  if (!out->flat || !out->const_block || !out->frame_block) return bad("missing gradient outputs");
  const b200r::TapeLayout T = b200r::tape_layout(*desc);
  const int ND = N * D, tpf = (ND + b200r::kTileRows - 1) / b200r::kTileRows, n_tiles = M * tpf;
  size_t na, ng, nm;
  b200r_tape_sizes(desc, M, N, D, &na, &ng, &nm);
  if (!tape->a || !tape->g || !tape->mask || tape->a_bytes < na || tape->g_bytes < ng || tape->mask_bytes < nm) return bad("tape buffers missing or too small");
  if ((reinterpret_cast<uintptr_t>(tape->g) & 1023)) return bad("tape buffers must be 1024-B aligned");
  if (workspace_bytes < b200r_workspace_bytes(desc, M)) return bad("workspace too small");
  b200r::DeviceGuard guard(h->device);
  if (!guard.ok) return fail(h, B200R_E_CUDA, "cudaSetDevice failed");
  cudaStream_t stream = (cudaStream_t)stream_;
  cudaError_t e;

  // per-frame blocks again (the workspace may have been reused since the forward)
  const int nl = (int)bp.layer_out.size();
  const b200r::LayerIds ids = b200r::layer_ids(*desc);
  b200r::PrologueParams pp;
  memset(&pp, 0, sizeof(pp));
  pp.cl = bp.prog.cl; pp.fl = bp.prog.fl; pp.desc = *desc; pp.par = *par; pp.fr = *fr;
  pp.workspace = (float*)workspace;
  pp.n_layers = nl; pp.rgb0_layer = ids.rgb0;
  for (int i = 0; i < nl; ++i) { pp.layer_out[i] = (int16_t)bp.layer_out[i]; pp.layer_in[i] = (int16_t)bp.layer_in[i]; }
  if (pts) {  // as in b200r_warp_fwd: no cameras; bias rows whose codes are absent are skipped
    pp.skip_cams = 1;
    for (int i = 0; i < nl; ++i)
      if (!par->bias[i]) pp.cl.plain_off[i] = -1;
  }
  if ((e = b200r::launch_prologue(pp, stream)) != cudaSuccess) return fail_cuda(h, e, "prologue kernel");

  // gradient scale
  b200r::ScaleParams sp;
  memset(&sp, 0, sizeof(sp));
  const size_t S = (size_t)M * ND;
  const float* gp[12] = {grads->rgb, grads->density, grads->vis, grads->feature, grads->xyz, grads->xyz_cam, grads->depth, grads->flow,
                         grads->cyc_dist, grads->delta_skin, grads->skin_entropy, grads->gauss_density};
  const int gw[12] = {3, 1, 1, 16, 3, 3, 1, 3, 1, 1, 1, 1};
  for (int i = 0; i < 12; ++i) { sp.ptr[i] = gp[i]; sp.n[i] = (long long)S * gw[i]; sp.weight[i] = i == 1 ? -1.f : 1.f; }
  if (pts) {
    memset(sp.ptr, 0, sizeof(sp.ptr));
    sp.ptr[0] = g_points; sp.n[0] = (long long)S * 3; sp.weight[0] = 1.f;
  }
  sp.logibeta = par->logibeta;
  sp.scale = h->d_scale;
  sp.amax_bits = reinterpret_cast<unsigned int*>(h->d_scale + 2);
  sp.bf16 = desc->operand_dtype == 1;
  if ((e = cudaMemsetAsync(h->d_scale + 2, 0, 4, stream)) != cudaSuccess) return fail_cuda(h, e, "memset");
  b200r::absmax_kernel<<<h->n_sm * 4, 256, 0, stream>>>(sp);
  b200r::scale_kernel<<<1, 1, 0, stream>>>(sp);
  if ((e = cudaGetLastError()) != cudaSuccess) return fail_cuda(h, e, "scale kernels");

  if ((e = cudaMemsetAsync(out->const_block, 0, (size_t)bp.prog.cl.n_floats * 4, stream)) != cudaSuccess) return fail_cuda(h, e, "memset");
  if ((e = cudaMemsetAsync(out->frame_block, 0, (size_t)M * bp.prog.fl.n_floats * 4, stream)) != cudaSuccess) return fail_cuda(h, e, "memset");

  b200r::BwdKernelParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.prog = bp.prog;
  kp.tape = T;
  kp.desc = dsc;
  kp.rays = *rays;
  kp.saved = *saved;
  kp.g = *grads;
  kp.packed_t = (const uint8_t*)packed_t;
  kp.workspace = (const float*)workspace;
  kp.tape_a = (const uint8_t*)tape->a;
  kp.tape_g = (uint8_t*)tape->g;
  kp.tape_mask = (const uint32_t*)tape->mask;
  kp.g_cblk = out->const_block;
  kp.g_fblk = out->frame_block;
  kp.scale = h->d_scale;
  if (desc->dense) {
    if (!saved->warp_pts) return bad("missing saved forward outputs (warp_pts)");
    kp.dense_w3[0] = par->weight[ids.dense[2]];
    kp.dense_w3[1] = par->weight[ids.dense[5]];
  }
  kp.M = M; kp.ND = ND; kp.tiles_per_frame = tpf; kp.n_tiles = n_tiles;
  if (pts) {
    kp.rays = rays_pts;
    kp.saved.xyz = const_cast<float*>(pts->xyz);  // the warp's input points take the canonical point's place
    kp.g_points = g_points;
    kp.g_points_out = g_points_out;
  }
  if ((e = b200r::launch_field_bwd(kp, h->n_sm, stream)) != cudaSuccess) return fail_cuda(h, e, "field_bwd kernel");

  // weight gradients
  std::vector<b200r::WgradJob> jobs = b200r::build_jobs(dsc, bp, T, out->weight_off, pts ? 2 : -1);
  std::vector<b200r::WgradWork> work;
  std::vector<int32_t> first;
  const int grid = h->n_sm;
  b200r::build_work(jobs, n_tiles, tpf, grid, work, first);
  struct { b200r_field_desc d; int n_tiles, tpf, grid, warp; } keyh = {dsc, n_tiles, tpf, grid, pts ? 1 : 0};
  const std::string key = b200r::table_key("wg", &keyh, sizeof(keyh), out->weight_off, sizeof(out->weight_off));
  void* d_jobs = b200r::cached_table(h, key + "j", jobs.data(), jobs.size() * sizeof(b200r::WgradJob), stream, &e);
  void* d_work = b200r::cached_table(h, key + "w", work.data(), work.size() * sizeof(b200r::WgradWork), stream, &e);
  void* d_first = b200r::cached_table(h, key + "f", first.data(), first.size() * sizeof(int32_t), stream, &e);
  if (!d_jobs || !d_work || !d_first) return fail_cuda(h, e, "job table upload");
  b200r::WgradParams wp;
  memset(&wp, 0, sizeof(wp));
  wp.tape_a = (const uint8_t*)tape->a;
  wp.tape_g = (const uint8_t*)tape->g;
  wp.n_a = T.n_a; wp.n_g = T.n_g;
  wp.jobs = (const b200r::WgradJob*)d_jobs;
  wp.work = (const b200r::WgradWork*)d_work;
  wp.cta_first = (const int32_t*)d_first;
  wp.grad = out->flat;
  wp.g_cblk = out->const_block;
  wp.g_fblk = out->frame_block;
  wp.frame_floats = bp.prog.fl.n_floats;
  wp.tiles_per_frame = tpf;
  wp.inv_scale = h->d_scale + 1;
  if ((e = b200r::launch_wgrad(wp, grid, dsc.operand_dtype, stream)) != cudaSuccess) return fail_cuda(h, e, "wgrad kernel");

  // backward of the per-frame prologue: blocks -> parameters (flat buffer) and per-frame inputs
  b200r::ChainParams cp;
  memset(&cp, 0, sizeof(cp));
  cp.cl = bp.prog.cl; cp.fl = bp.prog.fl; cp.desc = *desc; cp.par = *par; cp.fr = *fr;
  if (fgr) cp.gf = *fgr;
  cp.off = *out;
  cp.g_cblk = out->const_block; cp.g_fblk = out->frame_block; cp.grad = out->flat;
  cp.n_layers = nl;
  for (int i = 0; i < nl; ++i) cp.layer_out[i] = (int16_t)bp.layer_out[i];
  if (fgr) {  // code gradients are accumulated with atomics (mean / partner codes, shared rows)
    const size_t B4 = (size_t)desc->n_bones * 4;
    struct { float* ptr; size_t n; } z[] = {
        {fgr->inst_base, (size_t)M * 32}, {fgr->inst_color, (size_t)M * 32}, {fgr->inst_vis, (size_t)M * 32},
        {fgr->appr_code, (size_t)M * desc->appr_channels}, {fgr->inst_skin, (size_t)M * 32}, {fgr->skin_t_embed, (size_t)M * 128},
        {fgr->skin_t_embed_mean, 128}, {fgr->dense_t_embed, (size_t)M * 128}, {fgr->inst_dense_fwd, (size_t)M * 32},
        {fgr->inst_dense_bwd, (size_t)M * 32}, {fgr->t_art_qr, M * B4}, {fgr->t_art_qd, M * B4}, {fgr->rest_art_qr, M * B4},
        {fgr->rest_art_qd, M * B4}, {fgr->Kinv, (size_t)M * 9}, {fgr->field2cam_q, (size_t)M * 4}, {fgr->field2cam_t, (size_t)M * 3}};
    for (auto& it : z)
      if (it.ptr && it.n && (e = cudaMemsetAsync(it.ptr, 0, it.n * 4, stream)) != cudaSuccess) return fail_cuda(h, e, "memset");
  }
  if ((e = b200r::launch_chain(cp, stream)) != cudaSuccess) return fail_cuda(h, e, "chain kernel");
  return B200R_OK;
}

int b200r_field_bwd(b200r_handle* h, const b200r_field_desc* desc, const void* packed_t, const b200r_field_params* par,
                    const b200r_frame_tables* fr, const b200r_ray_batch* rays, const b200r_field_outputs* saved,
                    const b200r_field_grads* grads, const b200r_tape* tape, const b200r_param_grads* out,
                    const b200r_frame_grads* fgr, void* workspace, size_t workspace_bytes, b200r_stream stream_) {
  if (h && !rays) return fail(h, B200R_E_INVALID, "field_bwd: null argument");
  return run_field_bwd(h, desc, packed_t, par, fr, rays, nullptr, nullptr, nullptr, saved, grads, tape, out, fgr, workspace, workspace_bytes, stream_);
}

int b200r_warp_bwd(b200r_handle* h, const b200r_field_desc* desc, const void* packed_t, const b200r_field_params* par,
                   const b200r_frame_tables* fr, const b200r_point_batch* pts, const b200r_field_outputs* saved, const float* g_xyz,
                   const b200r_tape* tape, const b200r_param_grads* out, const b200r_frame_grads* fgr, float* g_points, void* workspace,
                   size_t workspace_bytes, b200r_stream stream_) {
  if (h && !pts) return fail(h, B200R_E_INVALID, "warp_bwd: null argument");
  b200r_field_grads none;
  memset(&none, 0, sizeof(none));
  return run_field_bwd(h, desc, packed_t, par, fr, nullptr, pts, g_xyz, g_points, saved, &none, tape, out, fgr, workspace, workspace_bytes, stream_);
}

int b200r_field_normals(b200r_handle* h, const b200r_field_desc* desc, const void* packed_t, const b200r_field_params* par,
                        const b200r_frame_tables* fr, const b200r_ray_batch* rays, const b200r_field_outputs* saved, const b200r_tape* tape,
                        float* g_cam, void* workspace, size_t workspace_bytes, b200r_stream stream_) {
  if (!h) return B200R_E_INVALID;
  auto bad = [&](const char* msg) { return fail(h, B200R_E_INVALID, std::string("field_normals: ") + msg); };
  if (!desc || !packed_t || !par || !fr || !rays || !saved || !tape || !g_cam || !workspace) return bad("null argument");
  b200r_field_desc dsc = *desc;
  if (dsc.operand_dtype == 2) dsc.operand_dtype = 0;
  b200r::BuiltProgram bp = b200r::build_bwd_program(dsc, /*density_only=*/true, /*warp_w=*/0);
  if (!bp.ok) return bad(bp.err);
  const int M = fr->M, N = rays->N, D = rays->D;
  if (M < 1 || N < 1 || D < 2) return bad("need M,N >= 1 and D >= 2");
  if (!rays->hxy || !fr->Kinv || !fr->near_far || !fr->field2cam_q || !fr->field2cam_t) return bad("missing ray/camera input");
  if (!saved->xyz || (desc->dense && !saved->warp_pts)) return bad("missing saved forward outputs (xyz; warp_pts for ComposedWarp fields)");
  const b200r::TapeLayout T = b200r::tape_layout(*desc);
  const int ND = N * D, tpf = (ND + b200r::kTileRows - 1) / b200r::kTileRows, n_tiles = M * tpf;
  size_t na, ng, nm;
  b200r_tape_sizes(desc, M, N, D, &na, &ng, &nm);
  if (!tape->a || !tape->g || !tape->mask || tape->a_bytes < na || tape->g_bytes < ng || tape->mask_bytes < nm) return bad("tape buffers missing or too small");
  if ((reinterpret_cast<uintptr_t>(tape->g) & 1023)) return bad("tape buffers must be 1024-B aligned");
  if (workspace_bytes < b200r_workspace_bytes(desc, M)) return bad("workspace too small");
  b200r::DeviceGuard guard(h->device);
  if (!guard.ok) return fail(h, B200R_E_CUDA, "cudaSetDevice failed");
  cudaStream_t stream = (cudaStream_t)stream_;
  cudaError_t e;
  const int nl = (int)bp.layer_out.size();
  const b200r::LayerIds ids = b200r::layer_ids(*desc);
  b200r::PrologueParams pp;
  memset(&pp, 0, sizeof(pp));
  pp.cl = bp.prog.cl; pp.fl = bp.prog.fl; pp.desc = *desc; pp.par = *par; pp.fr = *fr;
  pp.workspace = (float*)workspace;
  pp.n_layers = nl; pp.rgb0_layer = ids.rgb0;
  for (int i = 0; i < nl; ++i) { pp.layer_out[i] = (int16_t)bp.layer_out[i]; pp.layer_in[i] = (int16_t)bp.layer_in[i]; }
  if ((e = b200r::launch_prologue(pp, stream)) != cudaSuccess) return fail_cuda(h, e, "prologue kernel");
  b200r::BwdKernelParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.prog = bp.prog;
  kp.tape = T;
  kp.desc = dsc;
  kp.rays = *rays;
  kp.saved = *saved;
  kp.packed_t = (const uint8_t*)packed_t;
  kp.workspace = (const float*)workspace;
  kp.tape_a = (const uint8_t*)tape->a;
  kp.tape_g = (uint8_t*)tape->g;
  kp.tape_mask = (const uint32_t*)tape->mask;
  kp.g_points_out = g_cam;
  kp.normals = 1;
  kp.eik.scale_a = dsc.operand_dtype == 1 ? 1.0f : 256.0f;
  if (desc->dense) {
    kp.dense_w3[0] = par->weight[ids.dense[2]];
    kp.dense_w3[1] = par->weight[ids.dense[5]];
  }
  kp.M = M; kp.ND = ND; kp.tiles_per_frame = tpf; kp.n_tiles = n_tiles;
  if ((e = b200r::launch_field_bwd(kp, h->n_sm, stream)) != cudaSuccess) return fail_cuda(h, e, "normals kernel");
  return B200R_OK;
}

// ------------------------------------------------------------------ eikonal term
static constexpr float kEikScaleA = 256.0f;  // scale of the reverse chain's unit cotangent in its 16-bit operands

int b200r_eikonal_sizes(const b200r_field_desc* desc, int32_t n_rays, int32_t D, size_t* a_bytes, size_t* v_bytes) {
  if (!desc || n_rays < 1 || D < 1) return B200R_E_INVALID;
  const b200r::EikLayout E = b200r::eik_layout(*desc);
  const size_t tiles = ((size_t)n_rays * D + b200r::kTileRows - 1) / b200r::kTileRows + b200r::kMaxCtas;  // + scratch tiles of dead pair halves
  if (a_bytes) *a_bytes = tiles * E.n_a * b200r::kChunkBytes;
  if (v_bytes) *v_bytes = tiles * E.n_v * b200r::kChunkBytes;
  return B200R_OK;
}

static int eik_check(b200r_handle* h, const char* who, const b200r_field_desc* desc, const void* packed, const b200r_field_params* par,
                     const b200r_ray_batch* rays, int32_t M, const float* saved_xyz, const b200r_tape* tape, const b200r_eik_batch* eik, bool need_v) {
  auto bad = [&](const char* msg) { return fail(h, B200R_E_INVALID, std::string(who) + ": " + msg); };
  if (!desc || !packed || !par || !rays || !saved_xyz || !tape || !eik) return bad("null argument");
  if (M < 1 || rays->N < 1 || rays->D < 1 || eik->n_rays < 1 || !eik->rays) return bad("need M, N, D, n_rays >= 1");
  if (!tape->mask) return bad("the training tape (sign words) is missing");
  size_t na, nv;
  b200r_eikonal_sizes(desc, eik->n_rays, rays->D, &na, &nv);
  if (!eik->a || eik->a_bytes < na || (reinterpret_cast<uintptr_t>(eik->a) & 1023)) return bad("reverse-chain tape missing, too small or not 1024-B aligned");
  if (need_v && (!eik->v || eik->v_bytes < nv || (reinterpret_cast<uintptr_t>(eik->v) & 1023))) return bad("forward-chain tape missing, too small or not 1024-B aligned");
  return B200R_OK;
}

static void eik_common(b200r::BwdKernelParams& kp, const b200r_field_desc& dsc, const b200r_field_desc& tape_desc, const b200r_ray_batch* rays,
                       int32_t M, const float* saved_xyz, const b200r_tape* tape, const b200r_eik_batch* eik) {
  kp.tape = b200r::tape_layout(tape_desc);
  kp.desc = dsc;
  kp.rays = *rays;
  kp.saved.xyz = const_cast<float*>(saved_xyz);
  kp.tape_mask = (const uint32_t*)tape->mask;
  kp.M = M;
  kp.ND = rays->N * rays->D;
  kp.tiles_per_frame = (kp.ND + b200r::kTileRows - 1) / b200r::kTileRows;
  kp.eik.n_points = eik->n_rays * rays->D;
  kp.n_tiles = (kp.eik.n_points + b200r::kTileRows - 1) / b200r::kTileRows;
  kp.eik.rays_sel = eik->rays;
}

int b200r_eikonal_fwd(b200r_handle* h, const b200r_field_desc* desc, const void* packed_t, const b200r_field_params* par,
                      const b200r_ray_batch* rays, int32_t M, const float* saved_xyz, const b200r_tape* tape,
                      const b200r_eik_batch* eik, float* g_out, b200r_stream stream_) {
  if (!h) return B200R_E_INVALID;
  int rc = eik_check(h, "eikonal_fwd", desc, packed_t, par, rays, M, saved_xyz, tape, eik, false);
  if (rc != B200R_OK) return rc;
  if (!g_out || !par->sdf_w) return fail(h, B200R_E_INVALID, "eikonal_fwd: null argument");
  b200r_field_desc dsc = *desc;
  if (dsc.operand_dtype == 2) dsc.operand_dtype = 0;  // the W^T tiles are single fp16 (b200r_pack_weights_t)
  b200r::BuiltProgram bp = b200r::build_bwd_program(dsc, /*density_only=*/true);
  if (!bp.ok) return fail(h, B200R_E_INVALID, std::string("eikonal_fwd: ") + bp.err);
  const b200r::EikLayout E = b200r::eik_layout(*desc);
  b200r::DeviceGuard guard(h->device);
  if (!guard.ok) return fail(h, B200R_E_CUDA, "cudaSetDevice failed");
  b200r::BwdKernelParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.prog = bp.prog;
  eik_common(kp, dsc, *desc, rays, M, saved_xyz, tape, eik);
  kp.packed_t = (const uint8_t*)packed_t;
  kp.eik.mode = b200r::EIK_REVERSE;
  kp.eik.g_out = g_out;
  kp.eik.sdf_w = par->sdf_w;
  kp.eik.tape = (uint8_t*)eik->a;
  kp.eik.n_chunks = E.n_a;
  kp.eik.scale_a = dsc.operand_dtype == 1 ? 1.0f : kEikScaleA;
  kp.eik.head_chunk = (int16_t)E.a_head;
  kp.eik.v0_chunk = -1;
  for (int i = 0; i <= dsc.D; ++i) kp.eik.out_chunk[i] = (int16_t)E.a_base[i];
  cudaError_t e = b200r::launch_field_bwd(kp, h->n_sm, (cudaStream_t)stream_);
  if (e != cudaSuccess) return fail_cuda(h, e, "eikonal reverse chain");
  return B200R_OK;
}

int b200r_eikonal_bwd(b200r_handle* h, const b200r_field_desc* desc, const void* packed, const b200r_field_params* par,
                      const b200r_ray_batch* rays, int32_t M, const float* saved_xyz, const b200r_tape* tape,
                      const b200r_eik_batch* eik, const float* g_g, const b200r_param_grads* out, b200r_stream stream_) {
  if (!h) return B200R_E_INVALID;
  int rc = eik_check(h, "eikonal_bwd", desc, packed, par, rays, M, saved_xyz, tape, eik, true);
  if (rc != B200R_OK) return rc;
  if (!g_g || !out || !out->flat || out->sdf_w < 0) return fail(h, B200R_E_INVALID, "eikonal_bwd: null argument");
  b200r_field_desc dsc = *desc;            // the forward operand buffer's own layout (split mode: head + tail tiles) ...
  b200r_field_desc kdsc = *desc;
  if (kdsc.operand_dtype == 2) kdsc.operand_dtype = 0;  // ... read as single fp16 operands (the heads)
  const b200r::EikLayout E = b200r::eik_layout(*desc);
  const b200r::LayerIds ids = b200r::layer_ids(*desc);
  for (int i = 0; i <= desc->D; ++i)
    if (out->weight_off[ids.base[i]] < 0) return fail(h, B200R_E_INVALID, "eikonal_bwd: missing weight offset of a basefield layer");
  b200r::DeviceGuard guard(h->device);
  if (!guard.ok) return fail(h, B200R_E_CUDA, "cudaSetDevice failed");
  cudaStream_t stream = (cudaStream_t)stream_;
  cudaError_t e;
  const int P = eik->n_rays * rays->D;

  // scale of dL/dg: the largest entry lands at ~4 (v_0 multiplies it by up to 2^(L-1)); 1 / (scale * reverse-chain scale) for the flush
  float* d_sc = h->d_scale + 4;
  const float scale_a = kdsc.operand_dtype == 1 ? 1.0f : kEikScaleA;
  b200r::ScaleParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.ptr[0] = g_g; sp.n[0] = (long long)P * 3; sp.weight[0] = 256.0f;
  sp.scale = d_sc;
  sp.amax_bits = reinterpret_cast<unsigned int*>(d_sc + 2);
  sp.bf16 = kdsc.operand_dtype == 1;
  sp.inv_extra = 1.0f / scale_a;
  if ((e = cudaMemsetAsync(d_sc + 2, 0, 4, stream)) != cudaSuccess) return fail_cuda(h, e, "memset");
  b200r::absmax_kernel<<<h->n_sm, 256, 0, stream>>>(sp);
  b200r::scale_kernel<<<1, 1, 0, stream>>>(sp);
  if ((e = cudaGetLastError()) != cudaSuccess) return fail_cuda(h, e, "scale kernels");

  // forward chains A and B
  for (int which = b200r::EIK_CHAIN_A; which <= b200r::EIK_CHAIN_B; ++which) {
    b200r::BuiltProgram bp = b200r::build_eik_chain_program(dsc, which);
    if (!bp.ok) return fail(h, B200R_E_INVALID, std::string("eikonal_bwd: ") + bp.err);
    b200r::BwdKernelParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.prog = bp.prog;
    eik_common(kp, kdsc, *desc, rays, M, saved_xyz, tape, eik);
    kp.packed_t = (const uint8_t*)packed;
    kp.scale = d_sc;
    kp.eik.mode = which;
    kp.eik.gbar = g_g;
    kp.eik.tape = (uint8_t*)eik->v;
    kp.eik.n_chunks = E.n_v;
    kp.eik.head_chunk = -1;
    int nl = 0;
    if (which == b200r::EIK_CHAIN_A) {
      kp.eik.v0_chunk = (int16_t)E.v0;
      for (int i = 0; i <= desc->D; ++i, ++nl) { kp.eik.mask_slot[nl] = kp.tape.m_base[i]; kp.eik.out_chunk[nl] = (int16_t)E.vA[i]; }
    } else {
      kp.eik.v0_chunk = -1;
      for (int i = desc->skip; i <= desc->D; ++i, ++nl) { kp.eik.mask_slot[nl] = kp.tape.m_base[i]; kp.eik.out_chunk[nl] = (int16_t)E.vB[i]; }
    }
    kp.eik.n_layers = (int16_t)nl;
    if ((e = b200r::launch_field_bwd(kp, h->n_sm, stream)) != cudaSuccess) return fail_cuda(h, e, "eikonal forward chain");
  }

  // weight gradients: reverse-chain tape x forward-chain tape
  b200r::BuiltProgram bpl = b200r::build_program(kdsc);
  if (!bpl.ok) return fail(h, B200R_E_INVALID, std::string("eikonal_bwd: ") + bpl.err);
  std::vector<b200r::WgradJob> jobs = b200r::build_eik_jobs(kdsc, bpl, E, out->weight_off, out->sdf_w);
  std::vector<b200r::WgradWork> work;
  std::vector<int32_t> first;
  const int n_tiles = (P + b200r::kTileRows - 1) / b200r::kTileRows;
  const int grid = h->n_sm;
  b200r::build_work(jobs, n_tiles, n_tiles, grid, work, first);
  struct { b200r_field_desc d; int n_tiles, grid; int64_t sdf; } keyh = {kdsc, n_tiles, grid, out->sdf_w};
  const std::string key = b200r::table_key("ek", &keyh, sizeof(keyh), out->weight_off, sizeof(out->weight_off));
  void* d_jobs = b200r::cached_table(h, key + "j", jobs.data(), jobs.size() * sizeof(b200r::WgradJob), stream, &e);
  void* d_work = b200r::cached_table(h, key + "w", work.data(), work.size() * sizeof(b200r::WgradWork), stream, &e);
  void* d_first = b200r::cached_table(h, key + "f", first.data(), first.size() * sizeof(int32_t), stream, &e);
  if (!d_jobs || !d_work || !d_first) return fail_cuda(h, e, "job table upload");
  b200r::WgradParams wp;
  memset(&wp, 0, sizeof(wp));
  wp.tape_a = (const uint8_t*)eik->v;
  wp.tape_g = (const uint8_t*)eik->a;
  wp.n_a = E.n_v; wp.n_g = E.n_a;
  wp.jobs = (const b200r::WgradJob*)d_jobs;
  wp.work = (const b200r::WgradWork*)d_work;
  wp.cta_first = (const int32_t*)d_first;
  wp.grad = out->flat;
  wp.tiles_per_frame = n_tiles;
  wp.inv_scale = d_sc + 1;
  if ((e = b200r::launch_wgrad(wp, grid, kdsc.operand_dtype, stream)) != cudaSuccess) return fail_cuda(h, e, "eikonal wgrad kernel");
  return B200R_OK;
}

}  // extern "C"
