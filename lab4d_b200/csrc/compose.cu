// Depth-merge of two fields' per-sample arrays along every ray (MultiFields.compose_fields,
// lab4d/nnutils/multifields.py:339-398: concatenate the fields' samples, argsort by depth, gather every key).
//
// Both inputs are already sorted by depth along the ray (uniform sample placement), so the sort is a merge: the
// output position of sample i of field A is i + #{b : depth_b < depth_a[i]}, of sample j of field B it is
// j + #{a : depth_a <= depth_b[j]} (A first on ties, i.e. the stable order of the concatenation [A; B]).
// One 128-thread block per ray: binary searches in shared memory, then every channel array is written as one
// flat coalesced run of (Da + Db) * nch floats.  HBM/L2-bound: 4 B x (1 + sum of widths) per sample in and out.
#include <cuda_runtime.h>

#include "kernels.h"

namespace b200r {

constexpr int kComposeThreads = 128;

__global__ void __launch_bounds__(kComposeThreads) compose_fwd_kernel(const b200r_compose_args a) {
  extern __shared__ float sm[];
  const int Da = a.Da, Db = a.Db, Dt = Da + Db;
  float* da = sm;                                   // [Da]
  float* db = sm + Da;                              // [Db]
  int* perm = reinterpret_cast<int*>(sm + Dt);      // [Dt] output position -> index into [A; B]
  const int r = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < Da; i += kComposeThreads) da[i] = a.depth_a[(size_t)r * Da + i];
  for (int j = tid; j < Db; j += kComposeThreads) db[j] = a.depth_b[(size_t)r * Db + j];
  __syncthreads();
  for (int t = tid; t < Dt; t += kComposeThreads) {
    int lo = 0, hi, pos;
    if (t < Da) {  // lower bound of depth_a[t] in B
      const float key = da[t];
      hi = Db;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (db[mid] < key) lo = mid + 1; else hi = mid; }
      pos = t + lo;
    } else {       // upper bound of depth_b[j] in A
      const float key = db[t - Da];
      hi = Da;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (da[mid] <= key) lo = mid + 1; else hi = mid; }
      pos = (t - Da) + lo;
    }
    perm[pos] = t;
  }
  __syncthreads();
  if (a.perm)
    for (int t = tid; t < Dt; t += kComposeThreads) a.perm[(size_t)r * Dt + t] = perm[t];
  for (int c = 0; c < a.n_channels; ++c) {
    const int nch = a.nch[c];
    const float* sa = a.src_a[c];
    const float* sb = a.src_b[c];
    float* dst = a.dst[c] + (size_t)r * Dt * nch;
    const int n_el = Dt * nch;
    for (int e = tid; e < n_el; e += kComposeThreads) {
      const int pos = e / nch, ch = e - pos * nch;
      const int t = perm[pos];
      float v = 0.f;  // a key one field lacks reads as zeros (multifields.py:372-380)
      if (t < Da) { if (sa) v = sa[((size_t)r * Da + t) * nch + ch]; }
      else if (sb) v = sb[((size_t)r * Db + (t - Da)) * nch + ch];
      dst[e] = v;
    }
  }
}

// Backward: the merge is a permutation of the concatenated samples, so every key's gradient goes back through it - merged
// sample `pos` of ray r came from concatenated sample perm[pos].  Reads run flat over the merged gradient (coalesced), writes land
// in the two fields' arrays (rows of nch floats); a field that lacks the key (NULL) drops its share.  One launch for up to 16 keys
// instead of one index build + one torch.gather per key and field.
__global__ void __launch_bounds__(kComposeThreads) compose_bwd_kernel(const b200r_compose_bwd_args b) {
  extern __shared__ int sperm[];
  const int Da = b.Da, Dt = b.Da + b.Db;
  const int r = blockIdx.x, tid = threadIdx.x;
  for (int t = tid; t < Dt; t += kComposeThreads) sperm[t] = b.perm[(size_t)r * Dt + t];
  __syncthreads();
  for (int c = 0; c < b.n_channels; ++c) {
    const int nch = b.nch[c];
    const float* g = b.g_dst[c] + (size_t)r * Dt * nch;
    float* ga = b.g_a[c];
    float* gb = b.g_b[c];
    const int n_el = Dt * nch;
    for (int e = tid; e < n_el; e += kComposeThreads) {
      const int pos = e / nch, ch = e - pos * nch;
      const int t = sperm[pos];
      const float v = g[e];
      if (t < Da) { if (ga) ga[((size_t)r * Da + t) * nch + ch] = v; }
      else if (gb) gb[((size_t)r * b.Db + (t - Da)) * nch + ch] = v;
    }
  }
}

cudaError_t launch_compose_bwd(const b200r_compose_bwd_args& b, cudaStream_t stream) {
  const size_t smem = (size_t)(b.Da + b.Db) * sizeof(int);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(compose_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  compose_bwd_kernel<<<b.R, kComposeThreads, smem, stream>>>(b);
  return cudaGetLastError();
}

cudaError_t launch_compose_fwd(const b200r_compose_args& a, cudaStream_t stream) {
  const size_t smem = (size_t)(a.Da + a.Db) * 2 * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(compose_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  compose_fwd_kernel<<<a.R, kComposeThreads, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace b200r
