// Fused per-ray-sample forward of one Lab4D field (training-mode query_field) for sm_100a.
//
// Persistent kernel, one CTA per SM, CTAs paired in clusters of 2 that share every weight chunk
// through TMA multicast.  Each CTA walks 128-sample tiles (all samples of a tile belong to one frame).
// Warp roles (320 threads):
//   warps 0-7 : compute / epilogue.  Two threads per sample: warp w owns TMEM lanes 32*(w%4).. and
//               the column half (w/4) of every accumulator; bones and Fourier frequencies are split
//               the same way.  They place the sample on its ray, move it camera -> field space, run
//               dual-quaternion blend skinning, write 16-bit operand rows into swizzled shared memory
//               and run every layer's epilogue straight out of TMEM.
//   warp 8    : TMA producer - streams its half of each pre-packed weight chunk (cp.async.bulk,
//               multicast to both CTAs of the cluster) through a 3-stage ring of 32 KB.
//   warp 9    : tcgen05.mma issuer (one elected lane) + TMEM owner; walks the MmaStep list of program.h.
// The 256-wide chains (basefield, colorfield) are software-pipelined: every layer is issued as two
// N-halves into two TMEM accumulators, warps 0-3 / 4-7 run the epilogue of half 0 / 1 and write the
// 16-bit activations back to TMEM, where the next layer's MMAs read them as the A operand (TS form),
// so the epilogue of one half overlaps the MMAs of the other and of the next layer.
// Per-frame tables (cameras, bias rows with the per-frame codes folded in, bone transforms) and the
// constant block (plain biases, head weights) are staged in shared memory; hidden activations never
// leave the SM; HBM sees O(100 B) per sample of outputs.
//
// Restates (not ports) lab4d/nnutils/{nerf,deformable,feature,warping,skinning,embedding,visibility}.py
// and lab4d/utils/{render_utils,geom_utils,quat_transform}.py - see include/b200r.h for file:line.
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <type_traits>

#include "kernels.h"
#include "ptx.cuh"

#ifndef B200R_CLUSTER
#define B200R_CLUSTER 2
#endif

namespace b200r {

constexpr int kCluster = B200R_CLUSTER;
constexpr int kNumStages = 3;
constexpr int kComputeWarps = 8;
constexpr int kComputeThreads = kComputeWarps * 32;
constexpr int kThreads = kComputeThreads + 64;
constexpr int kSmemArena = kArenaChunks * kAChunkBytes;  //  96 KB
constexpr int kSmemRing = kNumStages * kWStageBytes;     //  96 KB

// ---------------------------------------------------------------------------------- small math
struct Q4 { float w, x, y, z; };
__device__ __forceinline__ Q4 qmul(const Q4& a, const Q4& b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
}
__device__ __forceinline__ Q4 qconj(const Q4& a) { return {a.w, -a.x, -a.y, -a.z}; }
// quaternion_apply: (q (0,p) q*)_xyz
__device__ __forceinline__ float3 qrot(const Q4& q, const float3& p) {
  Q4 t = qmul(q, Q4{0.f, p.x, p.y, p.z});
  Q4 r = qmul(t, qconj(q));
  return make_float3(r.x, r.y, r.z);
}
__device__ __forceinline__ Q4 ldq(const float* p) {
  float4 v = *reinterpret_cast<const float4*>(p);
  return {v.x, v.y, v.z, v.w};
}
template <class Op>
__device__ __forceinline__ void store_group(uint8_t* chunk, uint32_t row, uint32_t g, const float* v) {
  *reinterpret_cast<uint4*>(chunk + sw128_off(row, g)) =
      make_uint4(Op::pack2(v[0], v[1]), Op::pack2(v[2], v[3]), Op::pack2(v[4], v[5]), Op::pack2(v[6], v[7]));
}

// explicit shared-window accessors (32-bit addresses from smem_u32)
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts128f(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts16(uint32_t a, uint16_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(v) : "memory"); }
template <class Op>
__device__ __forceinline__ void sts_group(uint32_t a, const float* v) {
  sts128(a, make_uint4(Op::pack2(v[0], v[1]), Op::pack2(v[2], v[3]), Op::pack2(v[4], v[5]), Op::pack2(v[6], v[7])));
}

template <int V>
using IC = std::integral_constant<int, V>;

// ---------------------------------------------------------------------------------- the kernel
template <class Op, int B, int LMAX, bool DENSE>
__global__ void __launch_bounds__(kThreads, 1) field_fwd_kernel(const __grid_constant__ FieldKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* arena = smem;
  uint8_t* ring = smem + kSmemArena;
  float* cblk = reinterpret_cast<float*>(ring + kSmemRing);
  float* fblk = cblk + p.prog.cl.n_floats;
  uint64_t* bars = reinterpret_cast<uint64_t*>(fblk + p.prog.fl.n_floats);
  uint64_t* full_bar = bars;                  // [kNumStages]
  uint64_t* empty_bar = bars + kNumStages;    // [kNumStages]
  uint64_t* c2m = bars + 2 * kNumStages;      // [4] compute warps -> MMA thread, indexed by BAR_*
  uint64_t* m2c = bars + 2 * kNumStages + 4;  // [4] MMA thread (tcgen05.commit) -> compute warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kNumStages + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kNumStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], kCluster); }
    // one arrival per compute warp (lane 0 after __syncwarp), not per thread: 32x less mbarrier traffic
    mbar_init(&c2m[BAR_ALL], kComputeWarps);
    mbar_init(&c2m[BAR_H0], kComputeWarps);
    mbar_init(&c2m[BAR_H1], kComputeWarps);
    for (int i = 1; i < 4; ++i) mbar_init(&m2c[i], 1);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();  // peer barriers are initialised before any multicast can land
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const Program& P = p.prog;
  const int iters = (p.n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;  // identical in both CTAs of a cluster
  const uint32_t cta_rank = kCluster > 1 ? cluster_ctarank() : 0;
  const uint16_t cmask = (uint16_t)((1u << kCluster) - 1);

  if (warp == 8) {
    // =============================================================== TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int it = 0; it < iters; ++it) {
        for (int st = 0; st < P.n_steps; ++st) {
          const MmaStep& S = P.steps[st];
          const uint32_t bytes = (uint32_t)S.n * 128u * S.n_sub;
          const uint32_t part = bytes / kCluster;
          mbar_wait(&empty_bar[stage], phase ^ 1);  // both CTAs' MMAs are done with this slot
          mbar_arrive_expect_tx(&full_bar[stage], bytes);
          const uint8_t* src = p.packed + S.w_off + cta_rank * part;
          uint8_t* dst = ring + stage * kWStageBytes + cta_rank * part;
          if (kCluster > 1) tma_bulk_g2s_mcast(dst, src, part, &full_bar[stage], cmask);
          else tma_bulk_g2s(dst, src, part, &full_bar[stage]);
          if (++stage == kNumStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 9) {
    // =============================================================== MMA issuer
    // The whole warp walks the step list converged (all lanes poll the barriers); one elected lane issues.
    // Descriptors are base + small integer offsets (same swizzle/stride fields), so a step costs a handful of
    // integer adds besides the barrier polls; the common shapes (4 or 4+4 k-steps) are fully unrolled.
    {
      uint32_t stage = 0, phase = 0;
      uint32_t bar_phase = 0;  // bit i = parity of c2m[i]
      const uint64_t adesc0 = umma_desc_k_sw128(smem_u32(arena)), bdesc0 = umma_desc_k_sw128(smem_u32(ring));
      for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int st = 0; st < P.n_steps; ++st) {
          const MmaStep& S = P.steps[st];
          const uint32_t idesc = umma_idesc_f16(Op::kFmt, S.n);
          const uint32_t wt = S.wait;
          if (wt) {
            mbar_wait(&c2m[wt], (bar_phase >> wt) & 1u);
            bar_phase ^= 1u << wt;
          }
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint64_t bd = bdesc0 + (uint64_t)(stage * (kWStageBytes >> 4));
          const uint64_t bd2 = bd + (uint64_t)((uint32_t)S.n << 3);  // second tile: n * 128 B further
          const uint32_t d = tmem_base + S.d_col;
          const uint32_t acc0 = S.accumulate, ks = S.ksteps, ks2 = S.ksteps2;
          if (elect_one()) {
            if (S.a_kind == 0) {
              const uint64_t ad = adesc0 + (uint64_t)((uint32_t)S.a_chunk * (kAChunkBytes >> 4));
              const uint64_t ad2 = adesc0 + (uint64_t)((uint32_t)S.a_chunk2 * (kAChunkBytes >> 4));
              if (ks == 4) {
                umma_f16_ss(d, ad, bd, idesc, acc0);
                umma_f16_ss(d, ad + 2, bd + 2, idesc, 1u);
                umma_f16_ss(d, ad + 4, bd + 4, idesc, 1u);
                umma_f16_ss(d, ad + 6, bd + 6, idesc, 1u);
              } else {
                for (uint32_t k = 0; k < ks; ++k) umma_f16_ss(d, ad + 2 * k, bd + 2 * k, idesc, k ? 1u : acc0);
              }
              if (ks2 == 4) {
                umma_f16_ss(d, ad2, bd2, idesc, 1u);
                umma_f16_ss(d, ad2 + 2, bd2 + 2, idesc, 1u);
                umma_f16_ss(d, ad2 + 4, bd2 + 4, idesc, 1u);
                umma_f16_ss(d, ad2 + 6, bd2 + 6, idesc, 1u);
              } else {
                for (uint32_t k = 0; k < ks2; ++k) umma_f16_ss(d, ad2 + 2 * k, bd2 + 2 * k, idesc, 1u);
              }
            } else {
              const uint32_t a = tmem_base + S.a_tmem_col;  // 16 halves per k-step = 8 TMEM columns
              umma_f16_ts(d, a, bd, idesc, acc0);
              umma_f16_ts(d, a + 8, bd + 2, idesc, 1u);
              umma_f16_ts(d, a + 16, bd + 4, idesc, 1u);
              umma_f16_ts(d, a + 24, bd + 6, idesc, 1u);
              if (ks2) {
                umma_f16_ts(d, a + 32, bd2, idesc, 1u);
                umma_f16_ts(d, a + 40, bd2 + 2, idesc, 1u);
                umma_f16_ts(d, a + 48, bd2 + 4, idesc, 1u);
                umma_f16_ts(d, a + 56, bd2 + 6, idesc, 1u);
              }
            }
            // frees the ring slot (in both CTAs) once these MMAs have read it
            if (kCluster > 1) umma_commit_mcast(&empty_bar[stage], cmask);
            else umma_commit(&empty_bar[stage]);
            if (S.commit) umma_commit(&m2c[S.commit]);
          }
          __syncwarp();
          if (++stage == kNumStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // =============================================================== compute / epilogue warps
    const int q = warp & 3, hsel = warp >> 2;
    const uint32_t row = (uint32_t)(q * 32 + lane);  // tile row == TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t all_phase = 0, half_phase = 0;  // parity of m2c[BAR_ALL]; bit n of half_phase = parity of m2c[BAR_H0 + n]
    const int W = p.desc.W, HN = W / 2;
    // canonical layer ids (same enumeration as program.h layer_ids)
    const int lid_delta = 0, lid_vis = B > 0 ? 3 : 0, lid_base = lid_vis + 2, lid_rgb0 = lid_base + p.desc.D + 1,
              lid_color = lid_rgb0 + 1, lid_feat = lid_color + 3, lid_dense = lid_feat + (p.desc.has_feature ? 6 : 0);
    const ConstLayout& CL = P.cl;
    const FrameLayout& FL = P.fl;
    // 32-bit shared-window addresses (explicit ld/st.shared keeps the hot loops off the generic path)
    const uint32_t arena_s = smem_u32(arena), cblk_s = smem_u32(cblk), fblk_s = smem_u32(fblk);
    const uint32_t rowx = row * 128u + ((row & 7u) << 4);  // row base with the swizzle phase folded in:
                                                           // group g of this row lives at chunk + (rowx ^ (g << 4))
    const uint32_t sc_s = cblk_s + 4u * CL.scalars;

    // stage the constant block once (made visible by the first named barrier of the tile loop)
    {
      const float4* src = reinterpret_cast<const float4*>(p.workspace);
      float4* dst = reinterpret_cast<float4*>(cblk);
      for (int i = threadIdx.x; i < CL.n_floats / 4; i += kComputeThreads) dst[i] = __ldg(src + i);
    }

    // Exchange between the two threads of a row: 12 floats per thread inside the unused part of the
    // CH_EXTRA rows (the MMA only reads their first 32 B).  Every exchange round uses its own floats
    // and rounds that reuse an address are separated by a run_gemm() (a barrier of all compute threads).
    const uint32_t extra_s = arena_s + CH_EXTRA * kAChunkBytes;
    const uint32_t my_x0 = extra_s + (rowx ^ ((2u + 3u * hsel) << 4)), my_x1 = extra_s + (rowx ^ ((3u + 3u * hsel) << 4)),
                   my_x2 = extra_s + (rowx ^ ((4u + 3u * hsel) << 4));
    const uint32_t pr_x0 = extra_s + (rowx ^ ((2u + 3u * (hsel ^ 1)) << 4)), pr_x1 = extra_s + (rowx ^ ((3u + 3u * (hsel ^ 1)) << 4)),
                   pr_x2 = extra_s + (rowx ^ ((4u + 3u * (hsel ^ 1)) << 4));
    auto pair_sync = [&]() { named_bar_sync(1 + q, 64); };

    // hand the operands to the MMA warp, then wait for the layer's accumulator
    // every lane orders its own writes (generic -> async proxy, tcgen05), the warp converges, lane 0 signals
    auto warp_arrive = [&](uint64_t* bar) {
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    auto arrive_all = [&]() {
      fence_proxy_async_smem();
      tc_fence_before_sync();
      warp_arrive(&c2m[BAR_ALL]);
    };
    auto wait_all = [&]() {
      mbar_wait(&m2c[BAR_ALL], all_phase);
      all_phase ^= 1;
      tc_fence_after_sync();
    };
    auto run_gemm = [&]() { arrive_all(); wait_all(); };
    auto bias_s = [&](int layer) -> uint32_t { return (P.bias[layer].frame ? fblk_s : cblk_s) + 4u * P.bias[layer].off; };
    // relu(acc + bias) -> 16-bit operand rows; this thread covers its half of the columns.
    auto epi_relu_store = [&](uint32_t bias, int n_pad, int dst_chunk) {
      const int ncols = n_pad >> 1, cb = hsel * ncols, nblk = ncols >> 5;
      const uint32_t t0 = t_lane + kTmemD0 + cb;
      auto process = [&](uint32_t (&r)[32], int c0) {
        const uint32_t chunk = arena_s + (uint32_t)(dst_chunk + (c0 >> 6)) * kAChunkBytes;
        const uint32_t gbase = (uint32_t)(c0 & 63) >> 3;
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const float4 b0 = lds128(bias + 4u * (c0 + g8 * 8));
          const float4 b1 = lds128(bias + 4u * (c0 + g8 * 8 + 4));
          uint4 o;
          o.x = Op::pack2_relu(__uint_as_float(r[g8 * 8 + 0]) + b0.x, __uint_as_float(r[g8 * 8 + 1]) + b0.y);
          o.y = Op::pack2_relu(__uint_as_float(r[g8 * 8 + 2]) + b0.z, __uint_as_float(r[g8 * 8 + 3]) + b0.w);
          o.z = Op::pack2_relu(__uint_as_float(r[g8 * 8 + 4]) + b1.x, __uint_as_float(r[g8 * 8 + 5]) + b1.y);
          o.w = Op::pack2_relu(__uint_as_float(r[g8 * 8 + 6]) + b1.z, __uint_as_float(r[g8 * 8 + 7]) + b1.w);
          sts128(chunk + (rowx ^ ((gbase + g8) << 4)), o);
        }
      };
#pragma unroll 1
      for (int blk = 0; blk < nblk; ++blk) {
        uint32_t ra[32];
        tmem_ld32_issue(t0 + 32 * blk, ra);
        tmem_ld_wait32(ra);
        process(ra, cb + 32 * blk);
      }
    };
    // Pipelined chain: epilogue of N-half `nh` (accumulator D<nh>, HN columns) of one layer.  All 8 warps take
    // part: this thread covers HN/2 of the half's columns for its row.  relu(acc + bias) -> 16-bit activations ->
    // TMEM buffer `wbuf`, then signal the MMA thread that D<nh> is free and this part of the operand is written.
    auto wait_half = [&](int nh) {
      mbar_wait(&m2c[BAR_H0 + nh], (half_phase >> nh) & 1u);
      half_phase ^= 1u << nh;
      tc_fence_after_sync();
    };
    auto epi_half_to_tmem = [&](int layer, int wbuf, int nh) {
      wait_half(nh);
      const int c_lo = hsel * (HN >> 1);                // first column (inside the half) of this thread
      const int feat0 = nh * HN + c_lo;                 // same, as a feature index of the W-wide layer
      const uint32_t bias = bias_s(layer) + 4u * (uint32_t)feat0;
      const uint32_t tsrc = t_lane + (nh ? kTmemD1 : kTmemD0) + (uint32_t)c_lo;
      const uint32_t tdst = t_lane + (wbuf ? kTmemA1 : kTmemA0) + (uint32_t)(feat0 >> 1);
      auto pack_store = [&](uint32_t (&ra)[32], int blk) {
        uint32_t o[16];
#pragma unroll
        for (int g4 = 0; g4 < 8; ++g4) {
          const float4 b = lds128(bias + 4u * (32 * blk + 4 * g4));
          o[2 * g4] = Op::pack2_relu(__uint_as_float(ra[4 * g4 + 0]) + b.x, __uint_as_float(ra[4 * g4 + 1]) + b.y);
          o[2 * g4 + 1] = Op::pack2_relu(__uint_as_float(ra[4 * g4 + 2]) + b.z, __uint_as_float(ra[4 * g4 + 3]) + b.w);
        }
        tmem_st16(tdst + 16 * blk, o);
      };
      if (HN == 128) {  // 64 columns per thread: both TMEM loads in flight before the first wait
        uint32_t ra[32], rb[32];
        tmem_ld32_issue(tsrc, ra);
        tmem_ld32_issue(tsrc + 32, rb);
        tmem_ld_wait32(ra);
        pack_store(ra, 0);
        tmem_ld_wait32(rb);
        pack_store(rb, 1);
      } else {
        uint32_t ra[32];
        tmem_ld32_issue(tsrc, ra);
        tmem_ld_wait32(ra);
        pack_store(ra, 0);
      }
      tmem_st_wait();
      tc_fence_before_sync();
      warp_arrive(&c2m[BAR_H0 + nh]);
    };

    // 16-bit element `c` (0..63) of this row in operand chunk `chunk_s`
    auto put16 = [&](uint32_t chunk_s, int c, float val) { sts16(chunk_s + (rowx ^ ((uint32_t)(c >> 3) << 4)) + 2u * (c & 7), Op::cvt(val)); };
    // DenseWarp.forward (nnutils/warping.py:143-170): x + 0.1 * CondMLP([PE6(x), t, inst]).  The time / instance codes
    // are folded into the linear_1 bias row `bias1`; lid0 = canonical id of the map's linear_1.  Both threads of a
    // row compute the same result.
    auto dense_warp = [&](const float3& x, uint32_t bias1, int lid0) -> float3 {
      const uint32_t pe_s = arena_s + CH_PE * kAChunkBytes;
      if (hsel == 0) { put16(pe_s, 0, x.x); put16(pe_s, 1, x.y); put16(pe_s, 2, x.z); }
      else {
        put16(pe_s, 39, 0.f);  // 39 embedding columns; the third k-step reads up to column 47
        sts128(pe_s + (rowx ^ (5u << 4)), make_uint4(0u, 0u, 0u, 0u));
      }
      float fr = hsel == 0 ? 1.0f : 8.0f;
#pragma unroll 1
      for (int kf = 3 * hsel; kf < 3 * hsel + 3; ++kf) {
        float sv[3], cv[3];
        sincosf(fr * x.x, &sv[0], &cv[0]);
        sincosf(fr * x.y, &sv[1], &cv[1]);
        sincosf(fr * x.z, &sv[2], &cv[2]);
        const int e0 = 3 + 6 * kf;
#pragma unroll
        for (int c = 0; c < 3; ++c) { put16(pe_s, e0 + c, sv[c]); put16(pe_s, e0 + 3 + c, cv[c]); }
        fr *= 2.0f;
      }
      run_gemm();
      epi_relu_store(bias1, 256, CH_H0);
      run_gemm();
      epi_relu_store(bias_s(lid0 + 1), 256, CH_H0);
      run_gemm();
      float m[16];
      tmem_ld16(t_lane + kTmemD0, m);
      const uint32_t b3 = bias_s(lid0 + 2);
      return make_float3(x.x + 0.1f * (m[0] + lds32(b3)), x.y + 0.1f * (m[1] + lds32(b3 + 4)), x.z + 0.1f * (m[2] + lds32(b3 + 8)));
    };

    for (int it = 0; it < iters; ++it) {
      const int tile_raw = it * (int)gridDim.x + (int)blockIdx.x;
      const bool dead_tile = tile_raw >= p.n_tiles;
      const int tile = dead_tile ? p.n_tiles - 1 : tile_raw;
      const int f = tile / p.tiles_per_frame;
      const int r_raw = (tile - f * p.tiles_per_frame) * kTileRows + (int)row;
      const bool live = !dead_tile && r_raw < p.ND;
      const int r_in = r_raw < p.ND ? r_raw : p.ND - 1;
      const int n = r_in / p.rays.D;
      const int k = r_in - n * p.rays.D;
      const size_t s = (size_t)f * p.ND + r_in;

      // ------------------------------------------------ stage this frame's block in shared memory
      named_bar_sync(5, kComputeThreads);  // everyone is done with the previous block
      {
        const float4* src = reinterpret_cast<const float4*>(p.workspace + CL.n_floats + (size_t)f * FL.n_floats);
        float4* dst = reinterpret_cast<float4*>(fblk);
        for (int i = threadIdx.x; i < FL.n_floats / 4; i += kComputeThreads) dst[i] = __ldg(src + i);
      }
      named_bar_sync(5, kComputeThreads);

      // ------------------------------------------------ sample placement (sample_cam_rays)
      const float* hx = p.rays.hxy + ((size_t)f * p.rays.N + n) * 3;
      const float h0 = __ldg(hx), h1 = __ldg(hx + 1), h2 = __ldg(hx + 2);
      const float* cam = fblk + FL.cam;
      float3 d = make_float3(h0 * cam[0] + h1 * cam[1] + h2 * cam[2], h0 * cam[3] + h1 * cam[4] + h2 * cam[5],
                             h0 * cam[6] + h1 * cam[7] + h2 * cam[8]);
      const float dn = sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
      const float nearv = cam[9], farv = cam[10];
      const int Dn = p.rays.D;
      const float step = 1.0f / (float)(Dn - 1);
      auto lin = [&](int i) { return i < Dn / 2 ? step * (float)i : 1.0f - step * (float)(Dn - 1 - i); };
      auto depth_at = [&](int i) { float z = lin(i); return nearv * (1.0f - z) + farv * z; };
      const float depth = depth_at(k);
      const float delta = (k + 1 < Dn ? depth_at(k + 1) - depth : depth - depth_at(k - 1)) * dn;
      const float3 xyz_cam = make_float3(d.x * depth, d.y * depth, d.z * depth);
      const float3 dir_cam = make_float3(d.x / dn, d.y / dn, d.z / dn);

      // ------------------------------------------------ camera -> field (cam_to_field)
      const Q4 qc = {cam[11], cam[12], cam[13], cam[14]};
      const Q4 qi = qconj(qc);
      const float3 ti = qrot(qi, make_float3(-cam[15], -cam[16], -cam[17]));
      float3 xyz_t = qrot(qi, xyz_cam);
      xyz_t.x += ti.x; xyz_t.y += ti.y; xyz_t.z += ti.z;
      const float3 dir_f = qrot(qi, dir_cam);

      // ------------------------------------------------ skinning warps (SkinningWarp.forward), three per sample:
      //   w = 0 backward warp (time-t -> canonical), w = 1 forward warp with the pair partner's
      //   articulation (flow), w = 2 forward warp with the frame's own articulation (cycle).
      // bone coordinates -> delta MLP on the tensor pipe -> softmax -> dual-quaternion blend.
      // Half 0 owns bones [0,BS), half 1 bones [BS,B); operand groups are split at a 16-B boundary.
      constexpr int BS = B == 25 ? 13 : (B == 18 ? 8 : 0);
      constexpr int XTRA = (8 - (3 * BS) % 8) % 8;  // values of bone BS that complete half 0's last group
      constexpr int I0 = 3 * BS + XTRA;             // first operand column written by half 1
      constexpr int BH = BS > B - BS ? BS : B - BS;
      auto skin_warp = [&](auto half_tag, const float3& x, uint32_t binv, uint32_t se3, uint32_t bias1, float& entropy,
                           float& delta_skin) -> float3 {
        constexpr int HALF = decltype(half_tag)::value;
        constexpr int b_lo = HALF == 0 ? 0 : BS;
        constexpr int b_hi = HALF == 0 ? BS : B;
        constexpr int NB = b_hi - b_lo;
        constexpr int NV = HALF == 0 ? 3 * BS + XTRA : 3 * (B - BS);
        float dist2[BH > 0 ? BH : 1];
        {
          float v[NV > 0 ? NV : 1];
#pragma unroll
          for (int j = 0; j < (NV + 2) / 3; ++j) {
            const uint32_t ba = binv + 48u * (b_lo + j);
            const float4 r0 = lds128(ba), r1 = lds128(ba + 16), r2 = lds128(ba + 32);
            const float xb0 = r0.x * x.x + r0.y * x.y + r0.z * x.z + r0.w;
            const float xb1 = r1.x * x.x + r1.y * x.y + r1.z * x.z + r1.w;
            const float xb2 = r2.x * x.x + r2.y * x.y + r2.z * x.z + r2.w;
            if (j < NB) dist2[j] = xb0 * xb0 + xb1 * xb1 + xb2 * xb2;
            if (3 * j < NV) v[3 * j] = xb0;
            if (3 * j + 1 < NV) v[3 * j + 1] = xb1;
            if (3 * j + 2 < NV) v[3 * j + 2] = xb2;
          }
          if (HALF == 0) {
#pragma unroll
            for (int g = 0; g < NV / 8; ++g) sts_group<Op>(arena_s + CH_H0 * kAChunkBytes + (rowx ^ (g << 4)), v + 8 * g);
          } else {
            constexpr int END = (3 * B + 15) / 16 * 16;  // zero-padded to whole UMMA_K steps
#pragma unroll
            for (int idx = I0; idx < END; idx += 8) {
              float w8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) w8[j] = (idx + j < 3 * B) ? v[(idx + j < 3 * B) ? idx + j - 3 * BS : 0] : 0.f;
              sts_group<Op>(arena_s + (idx < 64 ? CH_H0 : CH_H1) * kAChunkBytes + (rowx ^ (((idx & 63) >> 3) << 4)), w8);
            }
          }
        }
        // delta_field.linear_1 / linear_2 (ReLU) and linear_final
        run_gemm();
        epi_relu_store(bias1, 64, CH_H2);
        run_gemm();
        epi_relu_store(bias_s(lid_delta + 1), 64, CH_H2);
        run_gemm();
        float dl[32];
        tmem_ld32(t_lane + kTmemD0, dl);
        const uint32_t b3 = bias_s(lid_delta + 2);
        float mx = -INFINITY, dsum = 0.f;
        int amax = 0;
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const float dv = 0.1f * fmaxf(dl[b_lo + j] + lds32(b3 + 4u * (b_lo + j)), 0.f);
          dsum += dv * dv;
          const float lg = -(dist2[j] + dv);
          dist2[j] = lg;
          if (lg > mx) { mx = lg; amax = b_lo + j; }
        }
        // round 1: global max / anchor bone (first maximum wins, like argmax)
        sts32(my_x0 + 8, mx);
        sts32(my_x0 + 12, __int_as_float(amax));
        pair_sync();
        {
          const float omx = lds32(pr_x0 + 8);
          const int oam = __float_as_int(lds32(pr_x0 + 12));
          const bool take = HALF == 0 ? (omx > mx) : (omx >= mx);
          if (take) { mx = omx; amax = oam; }
        }
        const float4 qa = lds128(se3 + 32u * amax);
        float se = 0.f;
        float4 qr = make_float4(0.f, 0.f, 0.f, 0.f), qd = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const uint32_t sa = se3 + 32u * (b_lo + j);
          const float e = __expf(dist2[j] - mx);
          se += e;
          const float4 r = lds128(sa), dq = lds128(sa + 16);
          const float dot = qa.x * r.x + qa.y * r.y + qa.z * r.z + qa.w * r.w;
          const float wgt = dot > 0.f ? e : -e;  // the softmax denominator cancels in the normalisation below
          qr.x += wgt * r.x; qr.y += wgt * r.y; qr.z += wgt * r.z; qr.w += wgt * r.w;
          qd.x += wgt * dq.x; qd.y += wgt * dq.y; qd.z += wgt * dq.z; qd.w += wgt * dq.w;
        }
        // round 2: partial sums
        sts32(my_x0, se);
        sts32(my_x0 + 4, dsum);
        sts128f(my_x1, qr);
        sts128f(my_x2, qd);
        pair_sync();
        {
          const float ose = lds32(pr_x0), ods = lds32(pr_x0 + 4);
          const float4 o1 = lds128(pr_x1), o2 = lds128(pr_x2);
          // add in bone order (half 0 first) so both threads of the row get bit-identical results
          if (HALF == 0) {
            se = se + ose; dsum = dsum + ods;
            qr = make_float4(qr.x + o1.x, qr.y + o1.y, qr.z + o1.z, qr.w + o1.w);
            qd = make_float4(qd.x + o2.x, qd.y + o2.y, qd.z + o2.z, qd.w + o2.w);
          } else {
            se = ose + se; dsum = ods + dsum;
            qr = make_float4(o1.x + qr.x, o1.y + qr.y, o1.z + qr.z, o1.w + qr.w);
            qd = make_float4(o2.x + qd.x, o2.y + qd.y, o2.z + qd.z, o2.w + qd.w);
          }
        }
        entropy = __logf(se);  // logsumexp - max  (cross_entropy_skin_loss)
        delta_skin = dsum / (float)(B > 0 ? B : 1);
        // stored order is (w,x,y,z) in (.x,.y,.z,.w)
        const float inv = rsqrtf(qr.x * qr.x + qr.y * qr.y + qr.z * qr.z + qr.w * qr.w);
        const Q4 Qr = {qr.x * inv, qr.y * inv, qr.z * inv, qr.w * inv};
        const Q4 Qd = {qd.x * inv, qd.y * inv, qd.z * inv, qd.w * inv};
        const Q4 tq = qmul(Qd, qconj(Qr));
        float3 o = qrot(Qr, x);
        o.x += 2.f * tq.x; o.y += 2.f * tq.y; o.z += 2.f * tq.z;
        return o;
      };

      float3 xyz = xyz_t, x_next = xyz_t;
      float ent_b = 0.f, dsk_b = 0.f, ent_out = 0.f, dsk_out = 0.f, cyc = 0.f;
      if constexpr (B > 0) {
        // ComposedWarp (warping.py:445-483) interleaves the DenseWarp soft deformation: backward = skin then dense,
        // forward = dense then skin.  One loop over stages keeps a single inlined copy of either body.
        constexpr int NST = DENSE ? 6 : 3;
        float3 cur = xyz_t;
#pragma unroll 1
        for (int stg = 0; stg < NST; ++stg) {
          const int w = DENSE ? (stg >> 1) : stg;
          if (DENSE && (stg == 1 || stg == 2 || stg == 4)) {
            const uint32_t bias1 = stg == 1 ? bias_s(lid_dense + 3) : (stg == 2 ? fblk_s + 4u * FL.dense1_partner : bias_s(lid_dense));
            cur = dense_warp(stg == 1 ? cur : xyz, bias1, stg == 1 ? lid_dense + 3 : lid_dense);
            if (stg == 1) xyz = cur;
            continue;
          }
          const float3 src = w == 0 ? xyz_t : (DENSE ? cur : xyz);
          const uint32_t binv = fblk_s + 4u * (w == 0 ? FL.binv_t : (w == 1 ? FL.binv_rest_partner : FL.binv_rest));
          const uint32_t se3 = fblk_s + 4u * (w == 0 ? FL.se3_bwd : (w == 1 ? FL.se3_fwd_partner : FL.se3_fwd));
          const uint32_t bias1 = w == 0 ? bias_s(lid_delta) : fblk_s + 4u * FL.delta1_fwd;  // forward warps: mean time code
          float e, dk;
          const float3 o = hsel == 0 ? skin_warp(IC<0>{}, src, binv, se3, bias1, e, dk) : skin_warp(IC<1>{}, src, binv, se3, bias1, e, dk);
          if (w == 0) { cur = o; xyz = o; ent_b = e; dsk_b = dk; }
          else if (w == 1) { x_next = o; }
          else {
            const float dx = o.x - xyz_t.x, dy = o.y - xyz_t.y, dz = o.z - xyz_t.z;
            cyc = sqrtf(dx * dx + dy * dy + dz * dz);
            ent_out = 0.5f * (e + ent_b);
            dsk_out = 0.5f * (dk + dsk_b);
          }
        }
      } else {
        x_next = xyz;
      }

      // ------------------------------------------------ positional embedding (PosEmbedding.forward)
      // Column e of the embedding: e < 3 -> x_e, else frequency (e-3)/6, sin for (e-3)%6 < 3.  Columns 0..62 live in
      // CH_PE (column 63 = 0), columns 63.. in CH_EXTRA.  Half 0 writes frequencies 0..LMAX/2-1, half 1 the rest.
      {
        const uint32_t pe_s = arena_s + CH_PE * kAChunkBytes;
        auto put = [&](int e, float val) { put16(e < 63 ? pe_s : extra_s, e < 63 ? e : e - 63, val); };  // embedding column e
        if (hsel == 0) { put(0, xyz.x); put(1, xyz.y); put(2, xyz.z); }
        else {
          sts16(pe_s + (rowx ^ (7u << 4)) + 14u, (uint16_t)0);  // zero pad column 63 of CH_PE
          if (LMAX > 10) {  // CH_EXTRA holds 12 values; its k-step reads 16 columns
            sts32(extra_s + (rowx ^ (1u << 4)) + 8u, 0.f);
            sts32(extra_s + (rowx ^ (1u << 4)) + 12u, 0.f);
          }
        }
        const int k0 = hsel == 0 ? 0 : LMAX / 2, k1 = hsel == 0 ? LMAX / 2 : LMAX;
        float fr = hsel == 0 ? 1.0f : (float)(1 << (LMAX / 2));
#pragma unroll 1
        for (int kf = k0; kf < k1; ++kf) {
          float sv[3], cv[3];
          sincosf(fr * xyz.x, &sv[0], &cv[0]);
          sincosf(fr * xyz.y, &sv[1], &cv[1]);
          sincosf(fr * xyz.z, &sv[2], &cv[2]);
          const int e0 = 3 + 6 * kf;
#pragma unroll
          for (int c = 0; c < 3; ++c) { put(e0 + c, sv[c]); put(e0 + 3 + c, cv[c]); }
          fr *= 2.0f;
        }
      }

      // ------------------------------------------------ visibility MLP (VisField.forward)
      run_gemm();
      epi_relu_store(bias_s(lid_vis), 64, CH_H0);
      run_gemm();
      float vis_out;
      {
        const uint32_t b2 = bias_s(lid_vis + 1) + 128u * hsel, vw = cblk_s + 4u * CL.vis_w + 128u * hsel;
        float v[32];
        tmem_ld32(t_lane + kTmemD0 + 32 * hsel, v);
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const float4 ba = lds128(b2 + 4u * j), bb = lds128(b2 + 4u * j + 16), wa = lds128(vw + 4u * j), wb = lds128(vw + 4u * j + 16);
          a0 += fmaxf(v[j] + ba.x, 0.f) * wa.x; a1 += fmaxf(v[j + 1] + ba.y, 0.f) * wa.y;
          a0 += fmaxf(v[j + 2] + ba.z, 0.f) * wa.z; a1 += fmaxf(v[j + 3] + ba.w, 0.f) * wa.w;
          a0 += fmaxf(v[j + 4] + bb.x, 0.f) * wb.x; a1 += fmaxf(v[j + 5] + bb.y, 0.f) * wb.y;
          a0 += fmaxf(v[j + 6] + bb.z, 0.f) * wb.z; a1 += fmaxf(v[j + 7] + bb.w, 0.f) * wb.w;
        }
        const float accv = a0 + a1;
        sts32(my_x0, accv);
        pair_sync();
        const float other = lds32(pr_x0);
        vis_out = (hsel == 0 ? accv + other : other + accv) + lds32(sc_s + 4u * SC_VIS_B);
      }

      // ------------------------------------------------ feature field (FeatureNeRF.compute_feat)
      float feat[16];
      if (p.desc.has_feature) {
#pragma unroll 1
        for (int i = 0; i < 5; ++i) {
          run_gemm();
          epi_relu_store(bias_s(lid_feat + i), 128, CH_H0);
        }
        run_gemm();
        if (hsel == 0) {  // warp-uniform: 16 outputs, one thread per row
          float v16[16];
          tmem_ld16(t_lane + kTmemD0, v16);
          const uint32_t bf = bias_s(lid_feat + 5);
          float nn = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) { feat[j] = v16[j] + lds32(bf + 4u * j); nn += feat[j] * feat[j]; }
          const float inv = rsqrtf(nn);
#pragma unroll
          for (int j = 0; j < 16; ++j) feat[j] *= inv;
        }
      }

      // ------------------------------------------------ density + colour chains (NeRF.forward), pipelined:
      // the MMA thread issues every W-wide layer as two N-halves; this thread finishes its half's epilogue
      // (writing 16-bit activations to TMEM) while the other half's / the next layer's MMAs run.
      arrive_all();  // embedding operands written, accumulators free
      int buf = 0;   // TMEM activation buffer the current layer READS; its epilogue writes buf ^ 1
#pragma unroll 1
      for (int j = 0; j < p.desc.D; ++j) {
        epi_half_to_tmem(lid_base + j, buf ^ 1, 0);
        epi_half_to_tmem(lid_base + j, buf ^ 1, 1);
        buf ^= 1;
      }
      float sdf;
      {
        // basefield.linear_final: features go to shared memory (rgb.0 reads them at the very end), sdf head in fp32
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 1
        for (int nh = 0; nh < 2; ++nh) {
          wait_half(nh);
          const int c_lo = hsel * (HN >> 1), feat0 = nh * HN + c_lo;
          const uint32_t bb = bias_s(lid_base + p.desc.D) + 4u * (uint32_t)feat0, sw = cblk_s + 4u * CL.sdf_w + 4u * (uint32_t)feat0;
          const uint32_t tsrc = t_lane + (nh ? kTmemD1 : kTmemD0) + (uint32_t)c_lo;
#pragma unroll 1
          for (int c0 = 0; c0 < (HN >> 1); c0 += 32) {
            float v[32];
            tmem_ld32(tsrc + c0, v);
            const int col = feat0 + c0;  // column of the full W-wide feature
            const uint32_t chunk = arena_s + (uint32_t)(CH_H0 + (col >> 6)) * kAChunkBytes;
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              const float4 b0 = lds128(bb + 4u * (c0 + g8 * 8)), b1 = lds128(bb + 4u * (c0 + g8 * 8 + 4));
              const float4 w0 = lds128(sw + 4u * (c0 + g8 * 8)), w1 = lds128(sw + 4u * (c0 + g8 * 8 + 4));
              float y[8];
              y[0] = fmaxf(v[g8 * 8 + 0] + b0.x, 0.f); y[1] = fmaxf(v[g8 * 8 + 1] + b0.y, 0.f);
              y[2] = fmaxf(v[g8 * 8 + 2] + b0.z, 0.f); y[3] = fmaxf(v[g8 * 8 + 3] + b0.w, 0.f);
              y[4] = fmaxf(v[g8 * 8 + 4] + b1.x, 0.f); y[5] = fmaxf(v[g8 * 8 + 5] + b1.y, 0.f);
              y[6] = fmaxf(v[g8 * 8 + 6] + b1.z, 0.f); y[7] = fmaxf(v[g8 * 8 + 7] + b1.w, 0.f);
              a0 += y[0] * w0.x; a1 += y[1] * w0.y; a2 += y[2] * w0.z; a3 += y[3] * w0.w;
              a0 += y[4] * w1.x; a1 += y[5] * w1.y; a2 += y[6] * w1.z; a3 += y[7] * w1.w;
              sts_group<Op>(chunk + (rowx ^ ((((uint32_t)(col & 63) >> 3) + g8) << 4)), y);
            }
          }
          fence_proxy_async_smem();
          tc_fence_before_sync();
          warp_arrive(&c2m[BAR_H0 + nh]);  // D<nh> is free, this part of the features is in shared memory
        }
        buf ^= 1;
        const float accs = (a0 + a1) + (a2 + a3);
        sts32(my_x1, accs);
        pair_sync();
        const float other = lds32(pr_x1);
        sdf = (hsel == 0 ? accs + other : other + accs) + lds32(sc_s + 4u * SC_SDF_B);
      }
      const float ibeta = lds32(sc_s + 4u * SC_IBETA);
      const float sgn = sdf > 0.f ? 1.f : (sdf < 0.f ? -1.f : 0.f);
      const float density = (0.5f + 0.5f * sgn * expm1f(-fabsf(sdf) * ibeta)) * ibeta;

      // colorfield: three more pipelined layers (the first reads the embedding again)
#pragma unroll 1
      for (int j = 0; j < 3; ++j) {
        epi_half_to_tmem(lid_color + j, buf ^ 1, 0);
        epi_half_to_tmem(lid_color + j, buf ^ 1, 1);
        buf ^= 1;
      }
      // rgb.0 on (base features from shared memory) + (colour features from TMEM), then rgb.2 + sigmoid
      wait_all();
      float rgb[3];
      {
        const int ncols = HN >> 1, cb = hsel * ncols;
        const uint32_t b0 = bias_s(lid_rgb0), w2 = cblk_s + 4u * CL.rgb2_w, wd = cblk_s + 4u * CL.dir_w;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll 1
        for (int c0 = cb; c0 < cb + ncols; c0 += 32) {
          float v[32];
          tmem_ld32(t_lane + kTmemD0 + c0, v);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bv = lds128(b0 + 4u * (c0 + j));
            const float4 wr = lds128(w2 + 4u * (c0 + j)), wg = lds128(w2 + 4u * (HN + c0 + j)), wb = lds128(w2 + 4u * (2 * HN + c0 + j));
            float pre[4] = {v[j] + bv.x, v[j + 1] + bv.y, v[j + 2] + bv.z, v[j + 3] + bv.w};
            if (p.desc.L_dir == 0) {
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const uint32_t da = wd + 12u * (c0 + j + u);
                pre[u] += lds32(da) * dir_f.x + lds32(da + 4) * dir_f.y + lds32(da + 8) * dir_f.z;
              }
            }
            const float h0_ = fmaxf(pre[0], 0.f), h1_ = fmaxf(pre[1], 0.f), h2_ = fmaxf(pre[2], 0.f), h3_ = fmaxf(pre[3], 0.f);
            a0 += h0_ * wr.x + h1_ * wr.y + h2_ * wr.z + h3_ * wr.w;
            a1 += h0_ * wg.x + h1_ * wg.y + h2_ * wg.z + h3_ * wg.w;
            a2 += h0_ * wb.x + h1_ * wb.y + h2_ * wb.z + h3_ * wb.w;
          }
        }
        sts128f(my_x2, make_float4(a0, a1, a2, 0.f));
        pair_sync();
        const float4 o = lds128(pr_x2);
        if (hsel == 0) { a0 = a0 + o.x; a1 = a1 + o.y; a2 = a2 + o.z; }
        else { a0 = o.x + a0; a1 = o.y + a1; a2 = o.z + a2; }
        a0 += lds32(sc_s + 4u * SC_RGB2_B0); a1 += lds32(sc_s + 4u * SC_RGB2_B1); a2 += lds32(sc_s + 4u * SC_RGB2_B2);
        rgb[0] = 1.f / (1.f + __expf(-a0)); rgb[1] = 1.f / (1.f + __expf(-a1)); rgb[2] = 1.f / (1.f + __expf(-a2));
      }
      if (hsel == 1 || !live) continue;  // half 0 writes the sample's outputs

      float flow[3];
      {
        // field_to_cam with the partner frame's camera, pinhole projection, flow (nerf.py:948-997)
        const float* cn = fblk + FL.cam_partner;
        const Q4 qn = {cn[11], cn[12], cn[13], cn[14]};
        float3 xc = qrot(qn, x_next);
        xc.x += cn[15]; xc.y += cn[16]; xc.z += cn[17];
        const float k0 = cn[0], k1 = cn[4], k2 = cn[2], k3 = cn[5];
        const float fx = 1.0f / k0, fy = 1.0f / k1, cx = -k2 / k0, cy = -k3 / k1;
        const float hxn = (fx * xc.x + cx * xc.z) / (xc.z + 1e-6f);
        const float hyn = (fy * xc.y + cy * xc.z) / (xc.z + 1e-6f);
        flow[0] = hxn - h0;
        flow[1] = hyn - h1;
        bool valid = xc.z > 1e-6f;
        if (p.rays.flow_thresh >= 0.f) valid = valid && (sqrtf(flow[0] * flow[0] + flow[1] * flow[1]) < p.rays.flow_thresh);
        flow[2] = valid ? 1.f : 0.f;
      }

      // ------------------------------------------------ Gaussian bone density (compute_gauss_density)
      // max_b exp(-d2_b / 2) = exp(-min_b d2_b / 2)
      float gdens = 0.f;
      if constexpr (B > 0) {
        float best = INFINITY;
        const uint32_t ctr = cblk_s + 4u * CL.center;
#pragma unroll 5
        for (int b = 0; b < B; ++b) {
          const float4 c = lds128(ctr + 16u * b);
          const float dx = xyz.x - c.x, dy = xyz.y - c.y, dz = xyz.z - c.z;
          best = fminf(best, dx * dx + dy * dy + dz * dz);
        }
        gdens = expf(-0.5f * (best / (0.01f * 0.01f))) * lds32(sc_s + 4u * SC_WARP_IBETA);
      }

      // ------------------------------------------------ per-sample outputs
      {
        auto st3 = [&](float* dst, float a, float b, float c) { if (dst) { dst[s * 3] = a; dst[s * 3 + 1] = b; dst[s * 3 + 2] = c; } };
        auto st1 = [&](float* dst, float a) { if (dst) dst[s] = a; };
        st3(p.out.rgb, rgb[0], rgb[1], rgb[2]);
        st1(p.out.density, density);
        st1(p.out.sdf, sdf);
        st1(p.out.vis, vis_out);
        st3(p.out.xyz, xyz.x, xyz.y, xyz.z);
        st3(p.out.xyz_cam, xyz_cam.x, xyz_cam.y, xyz_cam.z);
        st3(p.out.xyz_t, xyz_t.x, xyz_t.y, xyz_t.z);
        st3(p.out.dir, dir_f.x, dir_f.y, dir_f.z);
        st1(p.out.depth, depth * lds32(sc_s + 4u * SC_INV_SCALE));
        st1(p.out.deltas, delta);
        st3(p.out.flow, flow[0], flow[1], flow[2]);
        st1(p.out.cyc_dist, cyc);
        st1(p.out.delta_skin, dsk_out);
        st1(p.out.skin_entropy, ent_out);
        st1(p.out.gauss_density, gdens);
        if (p.out.feature && p.desc.has_feature) {
          float4* fo = reinterpret_cast<float4*>(p.out.feature + s * 16);
          fo[0] = make_float4(feat[0], feat[1], feat[2], feat[3]);
          fo[1] = make_float4(feat[4], feat[5], feat[6], feat[7]);
          fo[2] = make_float4(feat[8], feat[9], feat[10], feat[11]);
          fo[3] = make_float4(feat[12], feat[13], feat[14], feat[15]);
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();  // no CTA exits while its peer may still signal its barriers
  if (warp == 9) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <class Op, int B, int LMAX, bool DENSE>
static cudaError_t launch_one(const FieldKernelParams& p, int n_sm, cudaStream_t stream) {
  auto kern = field_fwd_kernel<Op, B, LMAX, DENSE>;
  const int smem = 1024 + kSmemArena + kSmemRing + (p.prog.cl.n_floats + p.prog.fl.n_floats) * 4 + 128;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  int grid = p.n_tiles < n_sm ? p.n_tiles : n_sm;
  grid = (grid + kCluster - 1) / kCluster * kCluster;
  if (grid > n_sm) grid -= kCluster;
  if (grid < kCluster) grid = kCluster;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

cudaError_t launch_field_fwd(const FieldKernelParams& p, int n_sm, cudaStream_t stream) {
  const bool bf = p.desc.operand_dtype == 1;
#define B200R_CASE(BN, LM, DN)                                                    \
  if (p.desc.n_bones == BN && p.Lmax == LM && (p.desc.dense != 0) == DN)          \
    return bf ? launch_one<OpBF16, BN, LM, DN>(p, n_sm, stream) : launch_one<OpF16, BN, LM, DN>(p, n_sm, stream);
  B200R_CASE(0, 10, false)
  B200R_CASE(0, 12, false)
  B200R_CASE(18, 12, false)
  B200R_CASE(25, 12, false)
  B200R_CASE(18, 12, true)
  B200R_CASE(25, 12, true)
#undef B200R_CASE
  return cudaErrorInvalidValue;
}

}  // namespace b200r
