// Weight packing: nn.Linear fp32 (out,in) row-major -> 16-bit [n_pad x 64] K-major SWIZZLE_128B chunks,
// laid out exactly as tcgen05.mma reads them from shared memory, so the field kernel can fetch a chunk
// with one 1-D TMA bulk copy.  Optionally folds the PosEmbedding annealing window
// (lab4d/nnutils/embedding.py:112-125) into the columns that multiply Fourier features.
#include <cuda_runtime.h>

#include "kernels.h"
#include "ptx.cuh"

namespace b200r {

// SPLIT: every [n_pad x 64] tile is followed by the tile of the fp16 rounding errors w - fp16(w) (operand_dtype 2):
// the field kernel multiplies both, so the weight enters the product with ~22 mantissa bits.
template <class Op, bool SPLIT>
__global__ void pack_kernel(const __grid_constant__ PackParams p) {
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= p.total_groups) return;
  const uint32_t byte = gid * 16u;
  int si = 0;
  for (int i = 1; i < p.n_slices; ++i)
    if (p.slices[i].dst_off <= byte) si = i;
  const PackSlice S = p.slices[si];
  uint32_t local = byte - S.dst_off;
  bool tail = false;
  if (SPLIT && local >= (uint32_t)S.n_pad * 128u) { tail = true; local -= (uint32_t)S.n_pad * 128u; }
  const uint32_t row = local / 128u;
  const uint32_t slot = (local % 128u) >> 4;   // physical 16-B slot in the row
  const uint32_t g = slot ^ (row & 7u);        // logical group (columns 8g..8g+7)
  const float* Wsrc = p.weights[S.layer];
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float v[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int col = (int)g * 8 + j * 2 + h;
      float x = 0.f;
      if ((int)row < S.n && col < S.ncols) {
        // forward tiles: rows = out-features; transposed (dgrad) tiles: rows = in-features, columns = out-features
        x = S.transpose ? Wsrc[(size_t)(S.col0 + col) * S.in_dim + S.row0 + (int)row] : Wsrc[(size_t)(S.row0 + (int)row) * S.in_dim + S.col0 + col];
        if (S.pe_window != 0 && p.alpha >= 0.f) {
          const int e = S.pe_col0 + (S.transpose ? (int)row : col);  // index inside the positional embedding
          if (e >= 3) {
            const int L = S.pe_window == 1 ? p.L_base : p.L_color;
            const int kf = (e - 3) / 6;
            float t = fminf(fmaxf(p.alpha * (float)L - (float)kf, 0.f), 1.f);
            x *= 0.5f * (1.f + cosf(3.14159265358979323846f * t + 3.14159265358979323846f));
          }
        }
      }
      if (SPLIT && tail) x -= __half2float(__float2half_rn(x));
      v[h] = x;
    }
    w[j] = Op::pack2(v[0], v[1]);
  }
  *reinterpret_cast<uint4*>(p.packed + byte) = make_uint4(w[0], w[1], w[2], w[3]);
}

cudaError_t launch_pack(const PackParams& p, int operand_dtype, cudaStream_t stream) {
  const int threads = 256;
  const int blocks = (int)((p.total_groups + threads - 1) / threads);
  if (operand_dtype == 1) pack_kernel<OpBF16, false><<<blocks, threads, 0, stream>>>(p);
  else if (operand_dtype == 2) pack_kernel<OpF16, true><<<blocks, threads, 0, stream>>>(p);
  else pack_kernel<OpF16, false><<<blocks, threads, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace b200r
