// Probe: tcgen05.mma with BOTH operands MN-major (the wgrad form dW = G^T A, reduction over tile rows).
// Operands are [128 rows x 64 features] fp16 chunks in the K-major SWIZZLE_128B image the field kernels write
// (row r at r*128 B, 16-B group g at ((g ^ (r & 7)) << 4)); read as MN-major they are 64 MN x 8 K atoms of 1024 B,
// SBO = 1024 (next 8 rows), LBO = 16384 (next 64 features = next chunk).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I lab4d_b200/csrc tools/mnmajor_probe.cu -o tools/mnmajor_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "ptx.cuh"
using namespace b200r;

__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// MC chunks of G (M = 64*MC features... M fixed 128 -> MC = 2), NC chunks of A (N = 64*NC)
template <int NC>
__global__ void __launch_bounds__(128, 1) probe(const uint8_t* G, const uint8_t* A, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sG = smem;
  uint8_t* sA = smem + 2 * 16384;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + NC * 16384);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(slot, 512);
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  const uint32_t tb = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], (2 + NC) * 16384);
    tma_bulk_g2s(sG, G, 2 * 16384, &bars[0]);
    tma_bulk_g2s(sA, A, NC * 16384, &bars[0]);
  }
  mbar_wait(&bars[0], 0);
  tc_fence_after_sync();
  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t N = 64 * NC;
      const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((128u >> 4) << 24);
      for (int k = 0; k < 8; ++k)  // 16 rows per MMA = 2 KB of every chunk
        umma_f16_ss(tb, desc_mn_sw128(smem_u32(sG) + 2048 * k, 16384), desc_mn_sw128(smem_u32(sA) + 2048 * k, 16384), idesc, k ? 1u : 0u);
      umma_commit(&bars[1]);
    }
    __syncwarp();
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after_sync();
  const uint32_t t_lane = tb + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < 64 * NC; c0 += 32) {
    float v[32];
    tmem_ld32(t_lane + c0, v);
    for (int j = 0; j < 32; ++j) out[(size_t)(warp * 32 + lane) * (64 * NC) + c0 + j] = v[j];
  }
  tc_fence_before_sync(); __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

static void fill_chunks(uint8_t* img, float* ref, int nchunks, unsigned seed) {  // ref [128 rows][64*nchunks]
  srand(seed);
  for (int c = 0; c < nchunks; ++c)
    for (int r = 0; r < 128; ++r)
      for (int f = 0; f < 64; ++f) {
        float v = (float)(rand() % 2001 - 1000) / 1000.f;
        __half h = __float2half(v);
        ref[(size_t)r * 64 * nchunks + c * 64 + f] = __half2float(h);
        size_t off = (size_t)c * 16384 + r * 128 + ((((f >> 3) ^ (r & 7)) << 4)) + 2 * (f & 7);
        *reinterpret_cast<__half*>(img + off) = h;
      }
}

template <int NC>
static int run() {
  const int N = 64 * NC;
  uint8_t *hG = (uint8_t*)malloc(2 * 16384), *hA = (uint8_t*)malloc(NC * 16384);
  float *rG = (float*)malloc(128 * 128 * 4), *rA = (float*)malloc(128 * N * 4), *hout = (float*)malloc(128 * N * 4);
  fill_chunks(hG, rG, 2, 1);
  fill_chunks(hA, rA, NC, 2);
  uint8_t *dG, *dA; float* dout;
  cudaMalloc(&dG, 2 * 16384); cudaMalloc(&dA, NC * 16384); cudaMalloc(&dout, 128 * N * 4);
  cudaMemcpy(dG, hG, 2 * 16384, cudaMemcpyHostToDevice); cudaMemcpy(dA, hA, NC * 16384, cudaMemcpyHostToDevice);
  const int smem = 1024 + (2 + NC) * 16384 + 256;
  cudaFuncSetAttribute(probe<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<NC><<<1, 128, smem>>>(dG, dA, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("NC=%d: %s\n", NC, cudaGetErrorString(e)); return 1; }
  cudaMemcpy(hout, dout, 128 * N * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int r = 0; r < 128; ++r) s += (double)rG[r * 128 + m] * rA[(size_t)r * N + n];
      maxerr = fmax(maxerr, fabs(s - hout[(size_t)m * N + n]));
      maxref = fmax(maxref, fabs(s));
    }
  printf("MN-major wgrad probe N=%d: max |err| %.3e (max |ref| %.3e) -> %s\n", N, maxerr, maxref, maxerr < 1e-3 * maxref ? "OK" : "MISMATCH");
  return maxerr < 1e-3 * maxref ? 0 : 2;
}

int main() {
  int rc = run<1>();
  rc |= run<2>();
  rc |= run<4>();
  return rc;
}
