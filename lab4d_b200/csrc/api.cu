// C ABI of libb200render.so (declarations and reference citations: include/b200r.h).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "api_util.h"

static b200r::BuiltProgram build(const b200r_field_desc& d) { return b200r::build_program(d); }

extern "C" {

int b200r_layer_count(const b200r_field_desc* desc) {
  if (!desc) return B200R_E_INVALID;
  return b200r::layer_ids(*desc).count;
}

size_t b200r_packed_bytes(const b200r_field_desc* desc) {
  if (!desc) return 0;
  b200r::BuiltProgram bp = build(*desc);
  return bp.ok ? bp.packed_bytes : 0;
}

int b200r_create(int device, b200r_handle** out) {
  if (!out) return B200R_E_INVALID;
  *out = nullptr;
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return B200R_E_CUDA;
  if (prop.major != 10) return B200R_E_ARCH;  // tcgen05 / TMEM exist on sm_100 only: no fallback path
  b200r_handle* h = new b200r_handle();
  h->device = device;
  h->n_sm = prop.multiProcessorCount;
  h->d_scale = nullptr;
  cudaError_t e2 = cudaSetDevice(device);
  if (e2 == cudaSuccess) e2 = cudaMalloc((void**)&h->d_scale, 64);  // [0..2] field backward, [4..6] eikonal backward
  if (e2 != cudaSuccess) { delete h; return B200R_E_CUDA; }
  *out = h;
  return B200R_OK;
}

void b200r_destroy(b200r_handle* h) {
  if (!h) return;
  for (auto& kv : h->tables)
    if (kv.second.dev) cudaFree(kv.second.dev);
  if (h->d_scale) cudaFree(h->d_scale);
  delete h;
}

const char* b200r_last_error(const b200r_handle* h) { return h ? h->err.c_str() : "null handle"; }

static int pack_common(b200r_handle* h, const b200r_field_desc* desc, const b200r_field_params* params, float alpha, void* packed,
                       size_t packed_bytes, b200r_stream stream_, bool transposed) {
  if (!h) return B200R_E_INVALID;
  if (!desc || !params || !packed) return fail(h, B200R_E_INVALID, "pack_weights: null argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  b200r_field_desc dsc = *desc;
  if (transposed && dsc.operand_dtype == 2) dsc.operand_dtype = 0;  // the backward runs on single fp16 operands
  b200r::BuiltProgram bp = transposed ? b200r::build_bwd_program(dsc) : build(dsc);
  if (!bp.ok) return fail(h, B200R_E_INVALID, std::string("pack_weights: ") + bp.err);
  const int n_weights = (int)bp.layer_out.size();
  if (packed_bytes < bp.packed_bytes) return fail(h, B200R_E_INVALID, "pack_weights: packed buffer too small");
  if ((reinterpret_cast<uintptr_t>(packed) & 15) != 0) return fail(h, B200R_E_INVALID, "pack_weights: packed must be 16-B aligned");
  for (int i = 0; i < n_weights; ++i)
    if (!params->weight[i]) return fail(h, B200R_E_INVALID, "pack_weights: null weight pointer");
  b200r::DeviceGuard guard(h->device);
  if (!guard.ok) return fail(h, B200R_E_CUDA, "cudaSetDevice failed");
  // the slice table only depends on the architecture: cached per descriptor, uploaded once on the caller's stream
  cudaError_t e = cudaSuccess;
  void* d_slices = b200r::cached_table(h, b200r::table_key(transposed ? "slicesT" : "slices", &dsc, sizeof(dsc)), bp.slices.data(),
                                       bp.slices.size() * sizeof(b200r::PackSlice), stream, &e);
  if (!d_slices) return fail_cuda(h, e, "slice table upload");
  b200r::PackParams pp;
  memset(&pp, 0, sizeof(pp));
  pp.slices = (const b200r::PackSlice*)d_slices;
  for (int i = 0; i < n_weights; ++i) pp.weights[i] = params->weight[i];
  pp.n_slices = (int)bp.slices.size();
  pp.total_groups = (uint32_t)(bp.packed_bytes / 16);
  pp.packed = (uint8_t*)packed;
  pp.alpha = alpha;
  pp.L_base = desc->L_xyz;
  pp.L_color = desc->L_xyz + 2;
  if ((e = b200r::launch_pack(pp, dsc.operand_dtype, stream)) != cudaSuccess) return fail_cuda(h, e, "pack kernel");
  return B200R_OK;
}

int b200r_pack_weights(b200r_handle* h, const b200r_field_desc* desc, const b200r_field_params* params, float alpha,
                       void* packed, size_t packed_bytes, b200r_stream stream) {
  return pack_common(h, desc, params, alpha, packed, packed_bytes, stream, false);
}

size_t b200r_packed_t_bytes(const b200r_field_desc* desc) {
  if (!desc) return 0;
  b200r_field_desc dsc = *desc;
  if (dsc.operand_dtype == 2) dsc.operand_dtype = 0;
  b200r::BuiltProgram bp = b200r::build_bwd_program(dsc);
  return bp.ok ? bp.packed_bytes : 0;
}

int b200r_pack_weights_t(b200r_handle* h, const b200r_field_desc* desc, const b200r_field_params* params, float alpha,
                         void* packed_t, size_t packed_bytes, b200r_stream stream) {
  return pack_common(h, desc, params, alpha, packed_t, packed_bytes, stream, true);
}

size_t b200r_workspace_bytes(const b200r_field_desc* desc, int32_t M) {
  if (!desc || M < 1) return 0;
  b200r::BuiltProgram bp = build(*desc);
  if (!bp.ok) return 0;
  return b200r::workspace_bytes_total(bp.prog, M);
}

// shared body of b200r_field_fwd (rays != NULL) and b200r_points_fwd (pts != NULL)
static int run_field(b200r_handle* h, const b200r_field_desc* desc, const void* packed, const b200r_field_params* par,
                     const b200r_frame_tables* fr, const b200r_ray_batch* rays, const b200r_point_batch* pts, int mode,
                     const b200r_field_outputs* out, void* workspace, size_t workspace_bytes, b200r_stream stream_,
                     const b200r_tape* tape = nullptr) {
  const bool warp = mode == b200r::MODE_WARP_BWD || mode == b200r::MODE_WARP_FWD;
  const char* who = warp ? "warp_fwd" : (pts ? "points_fwd" : "field_fwd");
  auto bad = [&](const char* msg) { return fail(h, B200R_E_INVALID, std::string(who) + ": " + msg); };
  if (!desc || !packed || !par || !fr || (!rays && !pts) || !out || !workspace) return bad("null argument");
  b200r::BuiltProgram bp = b200r::build_program(*desc, mode);
  if (!bp.ok) return bad(bp.err);
  const int M = fr->M;
  const int N = pts ? pts->P : rays->N, D = pts ? 1 : rays->D;
  if (M < 1 || N < 1 || (!pts && D < 2)) return bad("need M,N >= 1 and D >= 2");
  if ((long long)M * N * D > 0x7fffffffLL) return bad("too many samples");
  if (!pts) {
    if (M >= 2 && (M & 1)) return bad("frames must come in adjacent pairs (M even)");
    if (!rays->hxy || !fr->Kinv || !fr->near_far || !fr->field2cam_q || !fr->field2cam_t) return bad("missing ray/camera input");
    if (!fr->inst_vis) return bad("missing instance codes");
    if (!par->vis_final_w || !par->vis_final_b) return bad("missing head weights");
  } else if (warp) {
    if (!pts->xyz) return bad("missing points");
    if (out->rgb || out->density || out->sdf || out->vis || out->xyz_cam || out->xyz_t || out->dir || out->depth || out->deltas ||
        out->feature || out->flow || out->cyc_dist || out->gauss_density)
      return bad("only xyz, skin_entropy and delta_skin (and warp_pts with a tape) are produced");
  } else {
    if (!pts->xyz) return bad("missing points");
    if (desc->L_dir == 0 && !pts->dir && out->rgb) return bad("rgb needs view directions for this field");
    if (out->vis || out->xyz_cam || out->xyz_t || out->dir || out->depth || out->deltas || out->feature || out->flow ||
        out->cyc_dist || out->delta_skin || out->skin_entropy || out->gauss_density)
      return bad("only rgb, density, sdf and xyz are produced");
  }
  if (!warp) {
    if (!fr->inst_base || !fr->inst_color) return bad("missing instance codes");
    if (desc->appr_channels > 0 && !fr->appr_code) return bad("missing appearance codes");
  }
  if (!par->sdf_w || !par->sdf_b || !par->rgb2_w || !par->rgb2_b || !par->logibeta || !par->logscale) return bad("missing head weights");
  const int nl = (int)bp.layer_out.size();
  const b200r::LayerIds ids = b200r::layer_ids(*desc);
  for (int i = 0; i < nl; ++i) {
    const bool chain_layer = (i >= ids.base[0] && i <= ids.color[2]);  // basefield, rgb.0, colorfield are contiguous
    const bool warp_layer = (desc->n_bones > 0 && i <= ids.delta[2]) || (desc->dense && i >= ids.dense[0]);
    const bool needed = !pts || (warp ? warp_layer : chain_layer);
    if (needed && (!par->weight[i] || !par->bias[i])) return bad("missing layer weight/bias");
  }
  if (desc->n_bones > 0 && (!pts || warp)) {
    if (!fr->inst_skin || !fr->skin_t_embed || !fr->skin_t_embed_mean || !fr->t_art_qr || !fr->t_art_qd || !fr->rest_art_qr ||
        !fr->rest_art_qd || !par->warp_logibeta || !par->log_gauss)
      return bad("missing skinning input");
  }
  if (desc->dense && (!pts || warp) && (!fr->dense_t_embed || !fr->inst_dense_fwd || !fr->inst_dense_bwd)) return bad("missing dense-warp codes");
  if (reinterpret_cast<uintptr_t>(packed) & 15) return bad("packed must be 16-B aligned");
  if (reinterpret_cast<uintptr_t>(workspace) & 15) return bad("workspace must be 16-B aligned");
  if (workspace_bytes < b200r_workspace_bytes(desc, M)) return bad("workspace too small (see b200r_workspace_bytes)");
  b200r::DeviceGuard guard(h->device);
  if (!guard.ok) return fail(h, B200R_E_CUDA, "cudaSetDevice failed");
  cudaError_t e = cudaSuccess;
  cudaStream_t stream = (cudaStream_t)stream_;

  static_assert(sizeof(b200r::PrologueParams) <= 4096 && sizeof(b200r::FieldKernelParams) <= 16384, "kernel parameter space");
  b200r::PrologueParams pp;
  memset(&pp, 0, sizeof(pp));
  pp.cl = bp.prog.cl;
  pp.fl = bp.prog.fl;
  pp.desc = *desc;
  pp.par = *par;
  pp.fr = *fr;
  pp.workspace = (float*)workspace;
  pp.skip_cams = pts != nullptr;
  pp.skip_bones = pts != nullptr && !warp;
  pp.n_layers = nl;
  pp.rgb0_layer = ids.rgb0;
  for (int i = 0; i < nl; ++i) { pp.layer_out[i] = (int16_t)bp.layer_out[i]; pp.layer_in[i] = (int16_t)bp.layer_in[i]; }
  if (pts) {  // plain bias rows of layers that are not evaluated may be absent
    for (int i = 0; i < nl; ++i)
      if (!par->bias[i]) pp.cl.plain_off[i] = -1;
  }
  if ((e = b200r::launch_prologue(pp, stream)) != cudaSuccess) return fail_cuda(h, e, "prologue kernel");

  b200r::FieldKernelParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.prog = bp.prog;
  kp.desc = *desc;
  if (rays) kp.rays = *rays;
  kp.rays.N = N;
  kp.rays.D = D;
  if (pts) { kp.points = pts->xyz; kp.point_dirs = warp ? nullptr : pts->dir; kp.rays.flow_thresh = -1.f; }
  kp.warp_mode = warp ? mode : 0;
  kp.out = *out;
  kp.packed = (const uint8_t*)packed;
  kp.workspace = (const float*)workspace;
  kp.M = M;
  kp.ND = N * D;
  kp.tiles_per_frame = (kp.ND + b200r::kTileRows - 1) / b200r::kTileRows;
  kp.n_tiles = M * kp.tiles_per_frame;
  kp.Lmax = desc->L_xyz + 2 > 10 ? 12 : 10;
  kp.scratch = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(workspace) + b200r::scratch_offset_bytes(bp.prog, M));
  if (tape) {  // training forward: record the operand chunks and ReLU sign words
    kp.tape = b200r::tape_layout(*desc);
    size_t need_a = b200r::tape_a_bytes(kp.tape, kp.n_tiles), need_m = b200r::tape_mask_bytes(kp.tape, kp.n_tiles);
    if (!tape->a || !tape->mask || tape->a_bytes < need_a || tape->mask_bytes < need_m) return bad("tape buffers missing or too small (b200r_tape_sizes)");
    if ((reinterpret_cast<uintptr_t>(tape->a) & 1023) || (reinterpret_cast<uintptr_t>(tape->mask) & 15)) return bad("tape buffers must be 1024-B / 16-B aligned");
    if (warp) {
      if (!out->xyz || (desc->dense && !out->warp_pts)) return bad("the training warp must keep xyz (and warp_pts for ComposedWarp fields)");
    } else if (!out->xyz || !out->rgb || !out->sdf || (desc->has_feature && (!out->feature || !out->feat_norm)) || (desc->dense && !out->warp_pts))
      return bad("the training forward must keep xyz, rgb, sdf (and feature, feat_norm; warp_pts for ComposedWarp fields)");
    kp.tape_a = (uint8_t*)tape->a;
    kp.tape_mask = (uint32_t*)tape->mask;
    e = b200r::launch_field_fwd_train(kp, h->n_sm, stream);
  } else {
    e = b200r::launch_field_fwd(kp, h->n_sm, stream);
  }
  if (e != cudaSuccess) return fail_cuda(h, e, "field_fwd kernel");
  return B200R_OK;
}

int b200r_field_fwd_train(b200r_handle* h, const b200r_field_desc* desc, const void* packed, const b200r_field_params* par,
                          const b200r_frame_tables* fr, const b200r_ray_batch* rays, const b200r_field_outputs* out,
                          const b200r_tape* tape, void* workspace, size_t workspace_bytes, b200r_stream stream_) {
  if (!h) return B200R_E_INVALID;
  if (!rays || !tape) return fail(h, B200R_E_INVALID, "field_fwd_train: null argument");
  return run_field(h, desc, packed, par, fr, rays, nullptr, b200r::MODE_FIELD, out, workspace, workspace_bytes, stream_, tape);
}

int b200r_field_fwd(b200r_handle* h, const b200r_field_desc* desc, const void* packed, const b200r_field_params* par,
                    const b200r_frame_tables* fr, const b200r_ray_batch* rays, const b200r_field_outputs* out,
                    void* workspace, size_t workspace_bytes, b200r_stream stream_) {
  if (!h) return B200R_E_INVALID;
  if (!rays) return fail(h, B200R_E_INVALID, "field_fwd: null argument");
  return run_field(h, desc, packed, par, fr, rays, nullptr, b200r::MODE_FIELD, out, workspace, workspace_bytes, stream_);
}

int b200r_points_fwd(b200r_handle* h, const b200r_field_desc* desc, const void* packed, const b200r_field_params* par,
                     const b200r_frame_tables* fr, const b200r_point_batch* pts, const b200r_field_outputs* out,
                     void* workspace, size_t workspace_bytes, b200r_stream stream_) {
  if (!h) return B200R_E_INVALID;
  if (!pts) return fail(h, B200R_E_INVALID, "points_fwd: null argument");
  return run_field(h, desc, packed, par, fr, nullptr, pts, b200r::MODE_POINTS, out, workspace, workspace_bytes, stream_);
}

int b200r_warp_fwd(b200r_handle* h, const b200r_field_desc* desc, const void* packed, const b200r_field_params* par,
                   const b200r_frame_tables* fr, const b200r_point_batch* pts, int32_t backward, const b200r_field_outputs* out,
                   void* workspace, size_t workspace_bytes, b200r_stream stream_) {
  if (!h) return B200R_E_INVALID;
  if (!pts) return fail(h, B200R_E_INVALID, "warp_fwd: null argument");
  return run_field(h, desc, packed, par, fr, nullptr, pts, backward ? b200r::MODE_WARP_BWD : b200r::MODE_WARP_FWD, out, workspace,
                   workspace_bytes, stream_);
}

int b200r_compose_bwd(b200r_handle* h, const b200r_compose_bwd_args* b, b200r_stream stream) {
  if (!h) return B200R_E_INVALID;
  if (!b) return fail(h, B200R_E_INVALID, "compose_bwd: null argument");
  if (b->R < 1 || b->Da < 1 || b->Db < 1 || !b->perm) return fail(h, B200R_E_INVALID, "compose_bwd: need R, Da, Db >= 1 and the forward's permutation");
  if (b->Da + b->Db > 8192) return fail(h, B200R_E_INVALID, "compose_bwd: more than 8192 samples per ray unsupported");
  if (b->n_channels < 0 || b->n_channels > B200R_MAX_CHANNELS) return fail(h, B200R_E_INVALID, "compose_bwd: bad channel count");
  for (int c = 0; c < b->n_channels; ++c)
    if (!b->g_dst[c] || b->nch[c] < 1) return fail(h, B200R_E_INVALID, "compose_bwd: null channel gradient or bad width");
  b200r::DeviceGuard guard(h->device);
  if (!guard.ok) return fail(h, B200R_E_CUDA, "cudaSetDevice failed");
  cudaError_t e = b200r::launch_compose_bwd(*b, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda(h, e, "compose backward kernel");
  return B200R_OK;
}

int b200r_warp_fwd_train(b200r_handle* h, const b200r_field_desc* desc, const void* packed, const b200r_field_params* par,
                         const b200r_frame_tables* fr, const b200r_point_batch* pts, const b200r_field_outputs* out, const b200r_tape* tape,
                         void* workspace, size_t workspace_bytes, b200r_stream stream_) {
  if (!h) return B200R_E_INVALID;
  if (!pts || !tape) return fail(h, B200R_E_INVALID, "warp_fwd_train: null argument");
  return run_field(h, desc, packed, par, fr, nullptr, pts, b200r::MODE_WARP_FWD, out, workspace, workspace_bytes, stream_, tape);
}

static int check_composite(b200r_handle* h, const b200r_composite_args* a) {
  if (!a) return fail(h, B200R_E_INVALID, "composite: null argument");
  if (a->R < 1 || a->D < 1) return fail(h, B200R_E_INVALID, "composite: need R >= 1, D >= 1");
  if (a->D > 2048) return fail(h, B200R_E_INVALID, "composite: D > 2048 unsupported");
  if (!a->density || !a->deltas) return fail(h, B200R_E_INVALID, "composite: missing density/deltas");
  if (a->n_channels < 0 || a->n_channels > B200R_MAX_CHANNELS) return fail(h, B200R_E_INVALID, "composite: bad channel count");
  for (int c = 0; c < a->n_channels; ++c) {
    if (!a->src[c]) return fail(h, B200R_E_INVALID, "composite: null channel source");
    if (a->nch[c] < 1) return fail(h, B200R_E_INVALID, "composite: bad channel width");
    if (a->mode[c] < 0 || a->mode[c] > B200R_CH_VIS) return fail(h, B200R_E_INVALID, "composite: bad channel mode");
    if (a->mode[c] == B200R_CH_FLOW && a->nch[c] != 3) return fail(h, B200R_E_INVALID, "composite: flow needs 3 channels");
    if ((a->mode[c] == B200R_CH_WEIGHTSUM || a->mode[c] == B200R_CH_VIS) && a->nch[c] != 1)
      return fail(h, B200R_E_INVALID, "composite: density-type channel needs 1 channel");
  }
  return B200R_OK;
}

int b200r_composite_fwd(b200r_handle* h, const b200r_composite_args* a, b200r_stream stream) {
  if (!h) return B200R_E_INVALID;
  int rc = check_composite(h, a);
  if (rc) return rc;
  for (int c = 0; c < a->n_channels; ++c)
    if (!a->dst[c]) return fail(h, B200R_E_INVALID, "composite_fwd: null channel destination");
  b200r::DeviceGuard guard(h->device);
  if (!guard.ok) return fail(h, B200R_E_CUDA, "cudaSetDevice failed");
  cudaError_t e = cudaSuccess;
  if ((e = b200r::launch_composite_fwd(*a, (cudaStream_t)stream)) != cudaSuccess) return fail_cuda(h, e, "composite_fwd kernel");
  return B200R_OK;
}

int b200r_composite_bwd(b200r_handle* h, const b200r_composite_bwd_args* b, b200r_stream stream) {
  if (!h) return B200R_E_INVALID;
  if (!b) return fail(h, B200R_E_INVALID, "composite_bwd: null argument");
  int rc = check_composite(h, &b->fwd);
  if (rc) return rc;
  if (!b->g_density) return fail(h, B200R_E_INVALID, "composite_bwd: g_density required");
  b200r::DeviceGuard guard(h->device);
  if (!guard.ok) return fail(h, B200R_E_CUDA, "cudaSetDevice failed");
  cudaError_t e = cudaSuccess;
  if ((e = b200r::launch_composite_bwd(*b, (cudaStream_t)stream)) != cudaSuccess) return fail_cuda(h, e, "composite_bwd kernel");
  return B200R_OK;
}

int b200r_importance_fwd(b200r_handle* h, const b200r_importance_args* a, b200r_stream stream) {
  if (!h) return B200R_E_INVALID;
  if (!a) return fail(h, B200R_E_INVALID, "importance: null argument");
  if (a->R < 1 || a->Dc < 4 || a->Dc > 4096) return fail(h, B200R_E_INVALID, "importance: need R >= 1 and 4 <= Dc <= 4096");
  if (!a->depth_c || !a->weights || !a->depth_out) return fail(h, B200R_E_INVALID, "importance: null buffer");
  b200r::DeviceGuard guard(h->device);
  if (!guard.ok) return fail(h, B200R_E_CUDA, "cudaSetDevice failed");
  cudaError_t e = cudaSuccess;
  if ((e = b200r::launch_importance_fwd(*a, (cudaStream_t)stream)) != cudaSuccess) return fail_cuda(h, e, "importance kernel");
  return B200R_OK;
}

int b200r_compose_fwd(b200r_handle* h, const b200r_compose_args* a, b200r_stream stream) {
  if (!h) return B200R_E_INVALID;
  if (!a) return fail(h, B200R_E_INVALID, "compose: null argument");
  if (a->R < 1 || a->Da < 1 || a->Db < 1) return fail(h, B200R_E_INVALID, "compose: need R, Da, Db >= 1");
  if (a->Da + a->Db > 8192) return fail(h, B200R_E_INVALID, "compose: more than 8192 samples per ray unsupported");
  if (!a->depth_a || !a->depth_b) return fail(h, B200R_E_INVALID, "compose: missing depths");
  if (a->n_channels < 0 || a->n_channels > B200R_MAX_CHANNELS) return fail(h, B200R_E_INVALID, "compose: bad channel count");
  for (int c = 0; c < a->n_channels; ++c) {
    if (!a->dst[c]) return fail(h, B200R_E_INVALID, "compose: null channel destination");
    if (!a->src_a[c] && !a->src_b[c]) return fail(h, B200R_E_INVALID, "compose: channel without any source");
    if (a->nch[c] < 1) return fail(h, B200R_E_INVALID, "compose: bad channel width");
  }
  b200r::DeviceGuard guard(h->device);
  if (!guard.ok) return fail(h, B200R_E_CUDA, "cudaSetDevice failed");
  cudaError_t e = cudaSuccess;
  if ((e = b200r::launch_compose_fwd(*a, (cudaStream_t)stream)) != cudaSuccess) return fail_cuda(h, e, "compose kernel");
  return B200R_OK;
}

static int match_check(b200r_handle* h, const b200r_match_args* a, const char* who) {
  auto bad = [&](const char* m) { return fail(h, B200R_E_INVALID, std::string(who) + ": " + m); };
  if (!a) return bad("null argument");
  if (a->R < 1 || a->K < 1 || a->K > B200R_MATCH_MAX_K) return bad("need R >= 1 and 1 <= K <= 2048");
  if (!a->feat_px || !a->feat_can || !a->xyz_can || !a->idx || !a->logsigma || !a->xyz_matched) return bad("null tensor");
  return B200R_OK;
}

int b200r_match_fwd(b200r_handle* h, const b200r_match_args* a, b200r_stream stream) {
  if (!h) return B200R_E_INVALID;
  int rc = match_check(h, a, "match_fwd");
  if (rc != B200R_OK) return rc;
  b200r::DeviceGuard guard(h->device);
  if (!guard.ok) return fail(h, B200R_E_CUDA, "cudaSetDevice failed");
  cudaError_t e = b200r::launch_match_fwd(*a, h->n_sm, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda(h, e, "match kernel");
  return B200R_OK;
}

size_t b200r_match_scratch_floats(int32_t R, int32_t K) { return R < 1 || K < 1 ? 0 : b200r::match_partial_floats(R, K); }

int b200r_match_bwd(b200r_handle* h, const b200r_match_bwd_args* b, b200r_stream stream) {
  if (!h) return B200R_E_INVALID;
  if (!b) return fail(h, B200R_E_INVALID, "match_bwd: null argument");
  int rc = match_check(h, &b->fwd, "match_bwd");
  if (rc != B200R_OK) return rc;
  if (!b->g_out || !b->scratch || !b->fwd.lse) return fail(h, B200R_E_INVALID, "match_bwd: missing g_out, scratch or the forward's lse");
  b200r::DeviceGuard guard(h->device);
  if (!guard.ok) return fail(h, B200R_E_CUDA, "cudaSetDevice failed");
  cudaError_t e = b200r::launch_match_bwd(*b, b->scratch, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda(h, e, "match backward kernels");
  return B200R_OK;
}

static int loss_check(b200r_handle* h, const b200r_loss_args* a, const char* who) {
  auto bad = [&](const char* m) { return fail(h, B200R_E_INVALID, std::string(who) + ": " + m); };
  if (!a) return bad("null argument");
  if (a->M < 1 || a->N < 1 || a->field_type < 0 || a->field_type > 2 || !(a->train_res > 0.f)) return bad("bad M, N, field_type or train_res");
  if (!a->r_mask || !a->r_rgb || !a->r_depth || !a->r_flow || !a->b_mask || !a->b_vis2d || !a->b_is_detected || !a->b_rgb || !a->b_depth ||
      !a->b_flow || !a->b_flow_uct || !a->loss || !a->stats)
    return bad("null tensor");
  if (a->field_type == 2 && !a->r_mask_fg) return bad("comp needs rendered mask_fg");
  if (a->field_type != 1 && (!a->a_feature || !a->a_xy_reproj || !a->b_feature || !a->b_hxy)) return bad("fg / comp need the feature and reprojection tensors");
  return B200R_OK;
}

int b200r_loss_fwd(b200r_handle* h, const b200r_loss_args* a, b200r_stream stream) {
  if (!h) return B200R_E_INVALID;
  int rc = loss_check(h, a, "loss_fwd");
  if (rc != B200R_OK) return rc;
  b200r::DeviceGuard guard(h->device);
  if (!guard.ok) return fail(h, B200R_E_CUDA, "cudaSetDevice failed");
  cudaError_t e = b200r::launch_loss_fwd(*a, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda(h, e, "loss kernel");
  return B200R_OK;
}

int b200r_loss_bwd(b200r_handle* h, const b200r_loss_bwd_args* b, b200r_stream stream) {
  if (!h) return B200R_E_INVALID;
  if (!b || !b->g_loss) return fail(h, B200R_E_INVALID, "loss_bwd: null argument");
  int rc = loss_check(h, &b->fwd, "loss_bwd");
  if (rc != B200R_OK) return rc;
  b200r::DeviceGuard guard(h->device);
  if (!guard.ok) return fail(h, B200R_E_CUDA, "cudaSetDevice failed");
  cudaError_t e = b200r::launch_loss_bwd(*b, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda(h, e, "loss backward kernel");
  return B200R_OK;
}

static int quat_check(b200r_handle* h, const char* who, int64_t B, int32_t D1, int32_t D2) {
  if (B < 0 || !(D1 == 3 || D1 == 4) || !(D2 == 3 || D2 == 4)) return fail(h, B200R_E_INVALID, std::string(who) + ": need B >= 0 and operand widths 3 or 4");
  return B200R_OK;
}
#define B200R_QUAT_ENTRY(CALL, WHO)                                                      \
  b200r::DeviceGuard guard(h->device);                                                   \
  if (!guard.ok) return fail(h, B200R_E_CUDA, "cudaSetDevice failed");                   \
  if (B == 0) return B200R_OK;                                                           \
  cudaError_t e = CALL;                                                                  \
  if (e != cudaSuccess) return fail_cuda(h, e, WHO);                                     \
  return B200R_OK;

int b200r_quat_mul_fwd(b200r_handle* h, const float* a, const float* b, float* out, int64_t B, int32_t D1, int32_t D2, b200r_stream stream) {
  if (!h) return B200R_E_INVALID;
  if (quat_check(h, "quat_mul_fwd", B, D1, D2)) return B200R_E_INVALID;
  if (B > 0 && (!a || !b || !out)) return fail(h, B200R_E_INVALID, "quat_mul_fwd: null tensor");
  B200R_QUAT_ENTRY(b200r::launch_quat_mul_fwd(a, b, out, B, D1, D2, (cudaStream_t)stream), "quat_mul kernel")
}
int b200r_quat_mul_bwd(b200r_handle* h, const float* grad, const float* a, const float* b, float* g_a, float* g_b, int64_t B, int32_t D1, int32_t D2,
                       b200r_stream stream) {
  if (!h) return B200R_E_INVALID;
  if (quat_check(h, "quat_mul_bwd", B, D1, D2)) return B200R_E_INVALID;
  if (B > 0 && (!grad || !a || !b || !g_a || !g_b)) return fail(h, B200R_E_INVALID, "quat_mul_bwd: null tensor");
  B200R_QUAT_ENTRY(b200r::launch_quat_mul_bwd(grad, a, b, g_a, g_b, B, D1, D2, (cudaStream_t)stream), "quat_mul backward kernel")
}
int b200r_quat_mul_bwd_bwd(b200r_handle* h, const float* u1, const float* u2, const float* grad, const float* a, const float* b, float* g_grad, float* g_a,
                           float* g_b, int64_t B, int32_t D1, int32_t D2, b200r_stream stream) {
  if (!h) return B200R_E_INVALID;
  if (quat_check(h, "quat_mul_bwd_bwd", B, D1, D2)) return B200R_E_INVALID;
  if (B > 0 && (!u1 || !u2 || !grad || !a || !b || !g_grad || !g_a || !g_b)) return fail(h, B200R_E_INVALID, "quat_mul_bwd_bwd: null tensor");
  B200R_QUAT_ENTRY(b200r::launch_quat_mul_bwd_bwd(u1, u2, grad, a, b, g_grad, g_a, g_b, B, D1, D2, (cudaStream_t)stream), "quat_mul second backward kernel")
}
int b200r_quat_conj(b200r_handle* h, const float* q, float* out, int64_t B, b200r_stream stream) {
  if (!h) return B200R_E_INVALID;
  if (B < 0 || (B > 0 && (!q || !out))) return fail(h, B200R_E_INVALID, "quat_conj: bad argument");
  B200R_QUAT_ENTRY(b200r::launch_quat_conj(q, out, B, (cudaStream_t)stream), "quat_conj kernel")
}

}  // extern "C"
