// C ABI of libb200render.so (declarations and reference citations: include/b200r.h).
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "kernels.h"

struct b200r_handle {
  int device;
  int n_sm;
  std::string err;
};

static int fail(b200r_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}
static int fail_cuda(b200r_handle* h, cudaError_t e, const char* where) {
  return fail(h, B200R_E_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}

extern "C" {

int b200r_layer_count(const b200r_field_desc* desc) {
  if (!desc) return B200R_E_INVALID;
  return b200r::layer_ids(*desc).count;
}

size_t b200r_packed_bytes(const b200r_field_desc* desc) {
  if (!desc) return 0;
  b200r::BuiltProgram bp = b200r::build_program(*desc);
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
  *out = h;
  return B200R_OK;
}

void b200r_destroy(b200r_handle* h) {
  if (!h) return;
  delete h;
}

const char* b200r_last_error(const b200r_handle* h) { return h ? h->err.c_str() : "null handle"; }

int b200r_pack_weights(b200r_handle* h, const b200r_field_desc* desc, const float* const* weights, int n_weights,
                       float alpha, void* packed, size_t packed_bytes, b200r_stream stream_) {
  if (!h) return B200R_E_INVALID;
  if (!desc || !weights || !packed) return fail(h, B200R_E_INVALID, "pack_weights: null argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  b200r::BuiltProgram bp = b200r::build_program(*desc);
  if (!bp.ok) return fail(h, B200R_E_INVALID, std::string("pack_weights: ") + bp.err);
  if (n_weights != (int)bp.layer_out.size()) return fail(h, B200R_E_INVALID, "pack_weights: wrong number of layers");
  if (packed_bytes < bp.packed_bytes) return fail(h, B200R_E_INVALID, "pack_weights: packed buffer too small");
  if ((reinterpret_cast<uintptr_t>(packed) & 15) != 0) return fail(h, B200R_E_INVALID, "pack_weights: packed must be 16-B aligned");
  for (int i = 0; i < n_weights; ++i)
    if (!weights[i]) return fail(h, B200R_E_INVALID, "pack_weights: null weight pointer");
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess) return fail_cuda(h, e, "cudaSetDevice");
  if ((int)bp.slices.size() > b200r::kMaxPackSlices) return fail(h, B200R_E_INVALID, "pack_weights: too many slices");
  b200r::PackParams pp;
  memset(&pp, 0, sizeof(pp));
  for (size_t i = 0; i < bp.slices.size(); ++i) pp.slices[i] = bp.slices[i];
  for (int i = 0; i < n_weights; ++i) pp.weights[i] = weights[i];
  pp.n_slices = (int)bp.slices.size();
  pp.total_groups = (uint32_t)(bp.packed_bytes / 16);
  pp.packed = (uint8_t*)packed;
  pp.alpha = alpha;
  pp.L_base = desc->L_xyz;
  pp.L_color = desc->L_xyz + 2;
  if ((e = b200r::launch_pack(pp, desc->operand_dtype, stream)) != cudaSuccess) return fail_cuda(h, e, "pack kernel");
  return B200R_OK;
}

int b200r_field_fwd(b200r_handle* h, const b200r_field_desc* desc, const void* packed, const b200r_field_args* a,
                    b200r_stream stream_) {
  if (!h) return B200R_E_INVALID;
  if (!desc || !packed || !a) return fail(h, B200R_E_INVALID, "field_fwd: null argument");
  b200r::BuiltProgram bp = b200r::build_program(*desc);
  if (!bp.ok) return fail(h, B200R_E_INVALID, std::string("field_fwd: ") + bp.err);
  if (a->M < 1 || a->N < 1 || a->D < 2) return fail(h, B200R_E_INVALID, "field_fwd: need M,N >= 1 and D >= 2");
  if ((long long)a->M * a->N * a->D > 0x7fffffffLL) return fail(h, B200R_E_INVALID, "field_fwd: too many samples");
  if (a->M >= 2 && (a->M & 1)) return fail(h, B200R_E_INVALID, "field_fwd: frames must come in adjacent pairs (M even)");
  if (!a->hxy || !a->Kinv || !a->near_far || !a->field2cam || !a->logibeta || !a->logscale)
    return fail(h, B200R_E_INVALID, "field_fwd: missing ray/camera input");
  if (!a->sdf_w || !a->sdf_b || !a->rgb2_w || !a->rgb2_b || !a->vis_final_w || !a->vis_final_b)
    return fail(h, B200R_E_INVALID, "field_fwd: missing head weights");
  if (desc->L_dir == 0 && !a->rgb0_dir_w) return fail(h, B200R_E_INVALID, "field_fwd: rgb0_dir_w required when L_dir == 0");
  const int nl = (int)bp.layer_out.size();
  for (int i = 0; i < nl; ++i) {
    if (!a->bias[i]) return fail(h, B200R_E_INVALID, "field_fwd: missing bias row");
    if ((reinterpret_cast<uintptr_t>(a->bias[i]) & 15) || (a->bias_stride[i] & 3))
      return fail(h, B200R_E_INVALID, "field_fwd: bias rows must be 16-B aligned");
  }
  if (desc->n_bones > 0) {
    if (!a->bone_inv_t || !a->bone_inv_rest || !a->se3_bwd || !a->se3_fwd || !a->inv_gauss || !a->bone_center ||
        !a->warp_logibeta || !a->delta1_bias_fwd)
      return fail(h, B200R_E_INVALID, "field_fwd: missing skinning table");
  }
  if (reinterpret_cast<uintptr_t>(packed) & 15) return fail(h, B200R_E_INVALID, "field_fwd: packed must be 16-B aligned");
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess) return fail_cuda(h, e, "cudaSetDevice");
  e = b200r::launch_field_fwd_desc(*desc, bp.prog, *a, packed, h->n_sm, (cudaStream_t)stream_);
  if (e != cudaSuccess) return fail_cuda(h, e, "field_fwd kernel");
  return B200R_OK;
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
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess) return fail_cuda(h, e, "cudaSetDevice");
  if ((e = b200r::launch_composite_fwd(*a, (cudaStream_t)stream)) != cudaSuccess) return fail_cuda(h, e, "composite_fwd kernel");
  return B200R_OK;
}

int b200r_composite_bwd(b200r_handle* h, const b200r_composite_bwd_args* b, b200r_stream stream) {
  if (!h) return B200R_E_INVALID;
  if (!b) return fail(h, B200R_E_INVALID, "composite_bwd: null argument");
  int rc = check_composite(h, &b->fwd);
  if (rc) return rc;
  if (!b->g_density) return fail(h, B200R_E_INVALID, "composite_bwd: g_density required");
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess) return fail_cuda(h, e, "cudaSetDevice");
  if ((e = b200r::launch_composite_bwd(*b, (cudaStream_t)stream)) != cudaSuccess) return fail_cuda(h, e, "composite_bwd kernel");
  return B200R_OK;
}

}  // extern "C"
