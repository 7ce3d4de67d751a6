// Micro-probe: sustained tcgen05.mma issue/execute rate per SM for the operand modes the field kernel uses.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I lab4d_b200/csrc tools/mma_probe.cu -o tools/mma_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include "ptx.cuh"
using namespace b200r;

// mode: 0 = SS, 1 = TS ; N ; fill: 1 = a second warp streams 16 KB bulk copies into a ring meanwhile
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// variant: 0 = predicated asm (lane==0), 1 = elect + plain asm, 2 = elect + plain asm + TMEM base assumed 0
template <int MODE, int NN, int VARIANT>
__global__ void __launch_bounds__(96, 1) probe2(int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* A = smem; uint8_t* Bm = smem + 16384;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 49152 + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(slot, 512);
  fence_proxy_async_smem(); tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  const uint32_t tb_real = *slot;
  if (warp == 0) {
    const uint32_t tb = VARIANT == 2 ? 0u : tb_real;
    const uint32_t issue = lane == 0;
    const uint64_t ad = umma_desc_k_sw128(smem_u32(A)), bd = umma_desc_k_sw128(smem_u32(Bm));
    constexpr uint32_t idesc = umma_idesc_f16(0, NN);
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int r = 0; r < reps; ++r) {
      if (VARIANT == 0) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          if (MODE == 0) umma_f16_ss_pred(tb, ad + 2 * (k & 3), bd + 2 * (k & 3), idesc, k ? 1u : 0u, issue);
          else umma_f16_ts_pred(tb, tb + 256 + 8 * k, bd + 2 * (k & 3), idesc, k ? 1u : 0u, issue);
        }
        if ((r & 3) == 3) { umma_commit_pred(&bars[0], issue); mbar_wait(&bars[0], ph); ph ^= 1; }
      } else {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            if (MODE == 0) umma_f16_ss(tb, ad + 2 * (k & 3), bd + 2 * (k & 3), idesc, k ? 1u : 0u);
            else umma_f16_ts(tb, tb + 256 + 8 * k, bd + 2 * (k & 3), idesc, k ? 1u : 0u);
          }
          if ((r & 3) == 3) umma_commit(&bars[0]);
        }
        __syncwarp();
        if ((r & 3) == 3) { mbar_wait(&bars[0], ph); ph ^= 1; }
      }
    }
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb_real, 512);
}

__global__ void __launch_bounds__(96, 1) probe(int mode, int N, int reps, int fill, const uint8_t* gsrc, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* A = smem;                 // 16 KB  [128 x 64] halves
  uint8_t* Bm = smem + 16384;        // 32 KB  [256 x 64] halves
  uint8_t* ring = smem + 49152;      // 4 x 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(slot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = *slot;
  volatile __shared__ int stop;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  if (warp == 1 && fill) {
    if (lane == 0) {
      uint32_t st = 0, ph = 0; int n = 0;
      while (!stop) {
        mbar_arrive_expect_tx(&bars[4 + st], 16384);
        tma_bulk_g2s(ring + st * 16384, gsrc + (size_t)(n & 63) * 16384, 16384, &bars[4 + st]);
        mbar_wait(&bars[4 + st], ph);
        if (++st == 4) { st = 0; ph ^= 1; }
        ++n;
      }
    }
  } else if (warp == 0) {
    const uint32_t issue = lane == 0;
    const uint64_t ad = umma_desc_k_sw128(smem_u32(A)), bd = umma_desc_k_sw128(smem_u32(Bm));
    const uint32_t idesc = umma_idesc_f16(0, N);
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int r = 0; r < reps; ++r) {
      // 16 MMAs (one K=256 layer-half) then commit; wait every 4th group to bound the queue
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        if (mode == 0) umma_f16_ss_pred(tb, ad + 2 * (k & 3), bd + 2 * (k & 3), idesc, k ? 1u : 0u, issue);
        else umma_f16_ts_pred(tb, tb + 256 + 8 * k, bd + 2 * (k & 3), idesc, k ? 1u : 0u, issue);
      }
      if ((r & 3) == 3) {
        umma_commit_pred(&bars[0], issue);
        mbar_wait(&bars[0], ph);
        ph ^= 1;
      }
    }
    long long t1 = clock64();
    if (lane == 0) { out[blockIdx.x] = t1 - t0; stop = 1; }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  uint8_t* g; cudaMalloc(&g, 64 * 16384); cudaMemset(g, 0, 64 * 16384);
  long long* out; cudaMallocManaged(&out, 148 * sizeof(long long));
  const int smem = 1024 + 49152 + 65536 + 256;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 2000;
  struct { int mode, N, fill; const char* name; } cfg[] = {
      {0, 256, 0, "SS N=256"}, {0, 128, 0, "SS N=128"}, {0, 64, 0, "SS N=64"}, {1, 256, 0, "TS N=256"}, {1, 128, 0, "TS N=128"},
      {0, 256, 1, "SS N=256 + TMA fill"}, {0, 128, 1, "SS N=128 + TMA fill"}, {1, 128, 1, "TS N=128 + TMA fill"}, {1, 256, 1, "TS N=256 + TMA fill"}};
  for (auto& c : cfg) {
    for (int grid : {1, 148}) {
      probe<<<grid, 96, smem>>>(c.mode, c.N, reps, c.fill, g, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
      double cyc = (double)out[0] / (reps * 16.0);
      printf("%-24s grid %3d: %7.1f cycles / MMA (ideal %d)\n", c.name, grid, cyc, c.N / 2);
    }
  }
  printf("--- issue-code variants (0 = predicated asm, 1 = elect + plain asm, 2 = elect + TMEM base 0)\n");
#define RUN(MODE, NN, V)                                                                          \
  {                                                                                               \
    cudaFuncSetAttribute(probe2<MODE, NN, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
    probe2<MODE, NN, V><<<148, 96, smem>>>(reps, out);                                            \
    cudaError_t e = cudaDeviceSynchronize();                                                      \
    if (e != cudaSuccess) { printf("probe2: %s\n", cudaGetErrorString(e)); return 1; }            \
    printf("%s N=%3d variant %d: %7.1f cycles / MMA\n", MODE ? "TS" : "SS", NN, V, (double)out[0] / (reps * 16.0)); \
  }
  RUN(0, 256, 0) RUN(0, 256, 1) RUN(0, 256, 2)
  RUN(0, 128, 0) RUN(0, 128, 1) RUN(0, 128, 2)
  RUN(0, 64, 0) RUN(0, 64, 1) RUN(0, 64, 2)
  RUN(1, 128, 0) RUN(1, 128, 1) RUN(1, 128, 2)
  RUN(1, 256, 2)
  return 0;
}
