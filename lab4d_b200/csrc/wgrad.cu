// Weight-gradient kernel of the field backward: dW = sum over tiles of G^T A, G = masked gradient of a layer's
// pre-activation, A = the layer's input operand, both read from the training tape (program.h TapeLayout) as
// [128 rows x 64] 16-bit chunks.  The reduction runs over the tile's ROWS, so both operands are MN-major for tcgen05.mma:
// a chunk's K-major SWIZZLE_128B image is, read the other way, a stack of 64 (MN) x 8 (K) swizzle atoms of 1024 B
// (SBO = 1024: next 8 rows; LBO = chunk stride: next 64 features) - no transposition anywhere.
//
// Work is a host-built list of (job, tile range) items, a few per CTA (split-K over tiles).  A job multiplies up to 4 G chunks
// (M = 2 x 128 accumulator halves, 256 TMEM columns each) with up to 4 A chunks (N <= 256); its result is added
// (fp32 atomics, scaled by 1/grad_scale) into one or two strided views: a weight matrix in the reference's (out, in)
// layout, or a block of the per-frame gradient table (bone tables, flushed at every frame boundary).  Idle warps sum the
// G columns of the stages they pass through: bias gradients (per frame for the bias rows that carry a per-frame code).
//
// Warp roles (192 threads): warp 0 = TMA producer (64-row half tiles, 3-stage ring of 64 KB), warp 1 = MMA issuer,
// warps 2-5 = column sums + accumulator flush.  HBM-bound by design: ~2 MMAs worth of time per 64 KB stage.
#include <cuda_runtime.h>
#include <string.h>

#include "kernels.h"
#include "ptx.cuh"

namespace b200r {
namespace wg {

constexpr int kStages = 3;
constexpr int kHalfRows = 64;
constexpr int kHalfChunk = kHalfRows * 128;       // 8 KB: rows [64 h, 64 h + 64) of a chunk
constexpr int kOperandBytes = 4 * kHalfChunk;     // 32 KB per operand per stage
constexpr int kStageBytes = 2 * kOperandBytes;    // 64 KB
constexpr int kThreads = 192;

__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <class Op>
__global__ void __launch_bounds__(kThreads, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full_bar = bars;                 // [kStages]
  uint64_t* empty_bar = bars + kStages;      // [kStages]  MMA commit + 4 consumer warps
  uint64_t* acc_full = bars + 2 * kStages;   // MMA -> consumers: a segment's accumulators are complete
  uint64_t* acc_empty = acc_full + 1;        // consumers -> MMA: accumulators drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 5); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int w0 = p.cta_first[blockIdx.x], w1 = p.cta_first[blockIdx.x + 1];

  if (warp == 0) {
    // ================================================================= producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int wi = w0; wi < w1; ++wi) {
        const WgradWork W = p.work[wi];
        const WgradJob& J = p.jobs[W.job];
        const uint8_t* gsrc = J.g_src ? p.tape_g : p.tape_a;
        const uint8_t* asrc = J.a_src ? p.tape_g : p.tape_a;
        const size_t gstride = (size_t)(J.g_src ? p.n_g : p.n_a) * kChunkBytes, astride = (size_t)(J.a_src ? p.n_g : p.n_a) * kChunkBytes;
        const int n_g = J.n_g, n_a = J.n_a;
        const int gn = J.g_src ? p.n_g : p.n_a, an = J.a_src ? p.n_g : p.n_a;  // chunks per tile of each operand's tape
        const uint32_t gbytes = (uint32_t)n_g * kHalfChunk, abytes = (uint32_t)n_a * kHalfChunk;
        for (int t = W.tile0; t < W.tile1; ++t) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], gbytes + abytes);
            uint8_t* dst = smem + stage * kStageBytes;
            // tape layout [tile][64-row half][chunk]: the job's chunks of one half are one contiguous run
            tma_bulk_g2s(dst, gsrc + (size_t)t * gstride + ((size_t)hh * gn + J.g_chunk) * kHalfChunk, gbytes, &full_bar[stage]);
            tma_bulk_g2s(dst + kOperandBytes, asrc + (size_t)t * astride + ((size_t)hh * an + J.a_chunk) * kHalfChunk, abytes, &full_bar[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    uint32_t stage = 0, phase = 0, acc_phase = 0;
    bool drained = true;  // the accumulators are free
    for (int wi = w0; wi < w1; ++wi) {
      const WgradWork W = p.work[wi];
      const WgradJob& J = p.jobs[W.job];
      const uint32_t N = 64u * (uint32_t)J.n_a, n_mh = ((uint32_t)J.n_g + 1u) >> 1;
      const uint32_t idesc = (1u << 4) | (Op::kFmt << 7) | (Op::kFmt << 10) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((128u >> 4) << 24);
      uint32_t acc = 0;
      for (int t = W.tile0; t < W.tile1; ++t) {
        const bool seg_end = t + 1 == W.tile1 || (J.per_frame && (t + 1) % p.tiles_per_frame == 0);
        for (int hh = 0; hh < 2; ++hh) {
          if (!drained && acc == 0) {  // first MMA of a segment overwrites the accumulators: wait until they were read
            mbar_wait(acc_empty, acc_phase);
            acc_phase ^= 1;
            tc_fence_after_sync();
            drained = true;
          }
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          if (elect_one()) {
            const uint32_t sg = smem_u32(smem + stage * kStageBytes), sa = sg + kOperandBytes;
            for (uint32_t k = 0; k < 4; ++k)
              for (uint32_t mh = 0; mh < n_mh; ++mh)
                umma_f16_ss(tmem_base + 256u * mh, desc_mn_sw128(sg + mh * 2u * kHalfChunk + 2048u * k, kHalfChunk), desc_mn_sw128(sa + 2048u * k, kHalfChunk), idesc,
                            acc | k);
            umma_commit(&empty_bar[stage]);
            if (seg_end && hh == 1) umma_commit(acc_full);
          }
          __syncwarp();
          acc = 1;
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (seg_end) { acc = 0; drained = false; }
      }
    }
  } else {
    // ================================================================= consumers: column sums and accumulator flush
    const int q = warp & 3;                   // TMEM lane quadrant this warp may read
    const int ct = (warp - 2) * 32 + lane;    // 0..127
    const int cs_chunk = ct >> 5, cs_pair = ct & 31;
    uint32_t stage = 0, phase = 0, accf_phase = 0;
    const float inv_scale = p.inv_scale ? __ldg(p.inv_scale) : 1.0f;
    for (int wi = w0; wi < w1; ++wi) {
      const WgradWork W = p.work[wi];
      const WgradJob& J = p.jobs[W.job];
      float cs0 = 0.f, cs1 = 0.f;
      const bool do_cs = J.colsum != 0 && cs_chunk < J.n_g;
      for (int t = W.tile0; t < W.tile1; ++t) {
        const bool seg_end = t + 1 == W.tile1 || (J.per_frame && (t + 1) % p.tiles_per_frame == 0);
        const bool cs_end = t + 1 == W.tile1 || (J.colsum == 2 && (t + 1) % p.tiles_per_frame == 0);
        const int f = t / p.tiles_per_frame;
        for (int hh = 0; hh < 2; ++hh) {
          mbar_wait(&full_bar[stage], phase);
          if (do_cs) {
            const uint32_t base = smem_u32(smem + stage * kStageBytes) + cs_chunk * kHalfChunk + (cs_pair & 3) * 4;
#pragma unroll 8
            for (int r = 0; r < kHalfRows; ++r) {
              uint32_t v;
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(base + r * 128 + ((((uint32_t)cs_pair >> 2) ^ ((uint32_t)r & 7u)) << 4)));
              const float2 fv = Op::unpack2(v);
              cs0 += fv.x;
              cs1 += fv.y;
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (do_cs && cs_end) {
          float* dst = (J.colsum == 2 ? p.g_fblk + (size_t)f * p.frame_floats : p.g_cblk) + J.colsum_off;
          const int c0 = 64 * cs_chunk + 2 * cs_pair;
          if (c0 < J.colsum_n) atomicAdd(dst + c0, cs0 * inv_scale);
          if (c0 + 1 < J.colsum_n) atomicAdd(dst + c0 + 1, cs1 * inv_scale);
          cs0 = cs1 = 0.f;
        }
        if (seg_end) {
          mbar_wait(acc_full, accf_phase);
          accf_phase ^= 1;
          tc_fence_after_sync();
          const uint32_t n_mh = ((uint32_t)J.n_g + 1u) >> 1;
          for (int vi = 0; vi < J.n_views; ++vi) {
            const WgradView& V = J.v[vi];
            // destination: 0 flat weight buffer, 1 this frame's block gradient, 2 constant-block gradient
            float* dst = (V.per_frame == 1 ? p.g_fblk + (size_t)f * p.frame_floats : (V.per_frame == 2 ? p.g_cblk : p.grad)) + V.dst_off;
            for (uint32_t mh = 0; mh < n_mh; ++mh) {
              const int R = (int)mh * 128 + q * 32 + lane - V.row0;  // row of the view held by this thread
              if ((int)mh * 128 + 127 < V.row0 || (int)mh * 128 >= V.row0 + V.rows) continue;
              const bool row_ok = R >= 0 && R < V.rows;
              for (int cb = V.col0 & ~31; cb < V.col0 + V.cols; cb += 32) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + 256u * mh + (uint32_t)cb, v);
                if (row_ok) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) {
                    const int c = cb + j - V.col0;
                    if (c >= 0 && c < V.cols) atomicAdd(dst + (size_t)R * V.ld + c, v[j] * inv_scale);
                  }
                }
              }
            }
          }
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty);
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace wg

cudaError_t launch_wgrad(const WgradParams& p, int grid, int operand_dtype, cudaStream_t stream) {
  const int smem = 1024 + wg::kStages * wg::kStageBytes + 256;
  auto kern = operand_dtype == 1 ? wg::wgrad_kernel<OpBF16> : wg::wgrad_kernel<OpF16>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, wg::kThreads, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace b200r
