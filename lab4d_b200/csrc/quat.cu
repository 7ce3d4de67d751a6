// Stand-alone quaternion operators: the dqtorch extension of the reference (lab4d/third_party/quaternion/src/quaternion.cu:29-217,
// bound by src/bindings.cpp:7-16 and wrapped by third_party/quaternion/quaternion.py) - Hamilton product with its backward and
// backward-of-backward, conjugate.  3-vector operands are pure quaternions (w = 0), as in the reference kernels (:46-57).
// Everything is a Hamilton product:  out = a b;   g_a = cut(G b*), g_b = cut(a* G);   and for cotangents (u1, u2) of (g_a, g_b):
// g_G = u1 b + a u2,  g_a' = cut(G u2*),  g_b' = cut(u1* G)   (checked against second-order autograd in tests/test_quat_cpu.py).
// Elementwise, HBM-bound: one thread per quaternion, every operand read once, 128-bit accesses for 4-wide operands.
// (Inside the field kernels the same algebra is fused - csrc/field_fwd_kernel.cuh, field_bwd.cu, prologue.cu, chain.cu; these
// entries serve the reference's remaining torch code: per-frame pose modules, forward_project, eval-mode normals.)
#include <cuda_runtime.h>

#include "kernels.h"

namespace b200r {

struct Q { float w, x, y, z; };
__device__ __forceinline__ Q qld(const float* p, long long i, int D) {
  if (D == 4) { const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i); return {v.x, v.y, v.z, v.w}; }
  return {0.f, __ldg(p + 3 * i), __ldg(p + 3 * i + 1), __ldg(p + 3 * i + 2)};
}
__device__ __forceinline__ void qst(float* p, long long i, int D, const Q& q) {
  if (D == 4) { reinterpret_cast<float4*>(p)[i] = make_float4(q.w, q.x, q.y, q.z); return; }
  p[3 * i] = q.x; p[3 * i + 1] = q.y; p[3 * i + 2] = q.z;
}
__device__ __forceinline__ Q qm(const Q& a, const Q& b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
}
__device__ __forceinline__ Q qc(const Q& a) { return {a.w, -a.x, -a.y, -a.z}; }

__global__ void quat_mul_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long B, int D1, int D2) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= B) return;
  qst(out, i, 4, qm(qld(a, i, D1), qld(b, i, D2)));
}
__global__ void quat_mul_bwd_kernel(const float* __restrict__ g, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ ga,
                                    float* __restrict__ gb, long long B, int D1, int D2) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= B) return;
  const Q G = qld(g, i, 4), A = qld(a, i, D1), Bq = qld(b, i, D2);
  qst(ga, i, D1, qm(G, qc(Bq)));
  qst(gb, i, D2, qm(qc(A), G));
}
__global__ void quat_mul_bwd_bwd_kernel(const float* __restrict__ u1, const float* __restrict__ u2, const float* __restrict__ g, const float* __restrict__ a,
                                        const float* __restrict__ b, float* __restrict__ gg, float* __restrict__ gga, float* __restrict__ ggb,
                                        long long B, int D1, int D2) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= B) return;
  const Q U1 = qld(u1, i, D1), U2 = qld(u2, i, D2), G = qld(g, i, 4), A = qld(a, i, D1), Bq = qld(b, i, D2);
  const Q p = qm(U1, Bq), q = qm(A, U2);
  qst(gg, i, 4, Q{p.w + q.w, p.x + q.x, p.y + q.y, p.z + q.z});
  qst(gga, i, D1, qm(G, qc(U2)));
  qst(ggb, i, D2, qm(qc(U1), G));
}
__global__ void quat_conj_kernel(const float* __restrict__ q, float* __restrict__ out, long long B) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= B) return;
  qst(out, i, 4, qc(qld(q, i, 4)));
}

static inline unsigned qblocks(long long B) { return (unsigned)((B + 255) / 256); }

cudaError_t launch_quat_mul_fwd(const float* a, const float* b, float* out, long long B, int D1, int D2, cudaStream_t s) {
  quat_mul_fwd_kernel<<<qblocks(B), 256, 0, s>>>(a, b, out, B, D1, D2);
  return cudaGetLastError();
}
cudaError_t launch_quat_mul_bwd(const float* g, const float* a, const float* b, float* ga, float* gb, long long B, int D1, int D2, cudaStream_t s) {
  quat_mul_bwd_kernel<<<qblocks(B), 256, 0, s>>>(g, a, b, ga, gb, B, D1, D2);
  return cudaGetLastError();
}
cudaError_t launch_quat_mul_bwd_bwd(const float* u1, const float* u2, const float* g, const float* a, const float* b, float* gg, float* gga, float* ggb,
                                    long long B, int D1, int D2, cudaStream_t s) {
  quat_mul_bwd_bwd_kernel<<<qblocks(B), 256, 0, s>>>(u1, u2, g, a, b, gg, gga, ggb, B, D1, D2);
  return cudaGetLastError();
}
cudaError_t launch_quat_conj(const float* q, float* out, long long B, cudaStream_t s) {
  quat_conj_kernel<<<qblocks(B), 256, 0, s>>>(q, out, B);
  return cudaGetLastError();
}

}  // namespace b200r
