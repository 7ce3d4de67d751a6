// Backward of the fused field kernel (training-mode query_field) for sm_100a: the data-gradient pass.
//
// Same machine as the forward (field_fwd_kernel.cuh): persistent CTAs in clusters of 2, two 128-sample tiles in flight
// per CTA, one thread per sample, a TMA producer warp streaming packed weight tiles (here W^T tiles, csrc/program.h
// build_bwd_program) through the 3-slot ring with 2-CTA multicast, one tcgen05.mma issuer warp per tile group, gradient
// rows living in TMEM as 16-bit pairs (TS-form A operand), fp32 accumulators drained by the same threads.
// Per sample the thread
//   * recomputes the cheap fp32 geometry (sample placement, camera -> field, softmax / dual-quaternion blend of every
//     skinning warp from the delta-MLP outputs kept on the tape),
//   * turns the cotangents of the per-sample outputs (from b200r_composite_bwd) into the gradient of every layer's
//     pre-activation, layer by layer in reverse: G_{l-1} = (G_l W_l) * relu'(z_{l-1}) with the ReLU signs read from the
//     tape, writes each G to the gradient tape (operands of the weight-gradient kernel, csrc/wgrad.cu),
//   * back-propagates through the Fourier embeddings, the blend skinning, the flow projection and the camera in fp32,
//   * reduces the per-frame camera gradients with a warp butterfly + atomics into the gradient of the frame block.
// Gradients are multiplied by the power-of-two `grad_scale` on entry so that they survive the 16-bit operands; the
// weight-gradient kernel and the host divide it out.
//
// Hand-derived from (not autograd of) lab4d/nnutils/{nerf,deformable,feature,warping,skinning,embedding,visibility}.py and
// lab4d/utils/{render_utils,geom_utils,quat_transform}.py; checked stage by stage against oracle/field_backward.py.
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include "kernels.h"
#include "ptx.cuh"

#ifndef B200R_CLUSTER
#define B200R_CLUSTER 2
#endif

namespace b200r {
namespace bwd {

constexpr int kCluster = B200R_CLUSTER;
constexpr int kNumStages = 3;
constexpr int kGroups = 2;
constexpr int kGroupThreads = 128;
constexpr int kComputeThreads = kGroups * kGroupThreads;
constexpr int kThreads = kComputeThreads + 128;
constexpr int kRegsCompute = 208, kRegsAux = 88;
constexpr int kSmemRing = kNumStages * kWStageBytes;
constexpr int kTmemAcc = 0, kTmemAct = 256, kTmemGroup = 128;

struct Q4 { float w, x, y, z; };
__device__ __forceinline__ Q4 qmul(const Q4& a, const Q4& b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
}
__device__ __forceinline__ Q4 qconj(const Q4& a) { return {a.w, -a.x, -a.y, -a.z}; }
__device__ __forceinline__ Q4 qadd(const Q4& a, const Q4& b) { return {a.w + b.w, a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ float3 qrot(const Q4& q, const float3& p) {
  Q4 t = qmul(q, Q4{0.f, p.x, p.y, p.z});
  Q4 r = qmul(t, qconj(q));
  return make_float3(r.x, r.y, r.z);
}
// r = vec(q (0,p) q*): cotangent g of r -> (g_q, g_p); q need not be unit (oracle/field_backward.py _qrot_bwd)
__device__ __forceinline__ void qrot_bwd(const Q4& q, const float3& p, const float3& g, Q4& g_q, float3& g_p) {
  const Q4 G = {0.f, g.x, g.y, g.z}, Pq = {0.f, p.x, p.y, p.z};
  const Q4 u = qmul(q, Pq);
  const Q4 g_u = qmul(G, q);
  g_q = qadd(qmul(qconj(G), u), qmul(g_u, qconj(Pq)));
  const Q4 t = qmul(qconj(q), g_u);
  g_p = make_float3(t.x, t.y, t.z);
}
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

// backward of the Fourier embedding (nnutils/embedding.py:69-125): accumulator columns [0, 3 + 6 nfreq) of this thread's
// TMEM lane hold dL/d e; adds dL/dx = g_e[0:3] + sum_k 2^k (g_sin_k cos(2^k x) - g_cos_k sin(2^k x)) to gx, same
// double-angle walk as the forward.  One out-of-line copy: it is called from six places of the kernel.
__device__ __noinline__ void pe_backward_fn(uint32_t tD, const float3& x, int nfreq, float3& gx) {
  float ge[96];  // dynamically indexed below: lives in local memory (L1), not in registers
  const int ncol = 3 + 6 * nfreq;
#pragma unroll 1
  for (int c0 = 0; c0 < ncol; c0 += 32) {
    float v[32];
    tmem_ld32(tD + c0, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) ge[c0 + j] = v[j];
  }
  float ax = ge[0], ay = ge[1], az = ge[2];
  float fr = 1.0f, s0 = 0.f, s1 = 0.f, s2 = 0.f, c0 = 1.f, c1 = 1.f, c2 = 1.f;
#pragma unroll 1
  for (int kf = 0; kf < nfreq; ++kf) {
    if ((kf & 3) == 0) {
      sincosf(fr * x.x, &s0, &c0);
      sincosf(fr * x.y, &s1, &c1);
      sincosf(fr * x.z, &s2, &c2);
    } else {
      const float t0 = 2.f * s0 * c0, t1 = 2.f * s1 * c1, t2 = 2.f * s2 * c2;
      c0 = 1.f - 2.f * s0 * s0; c1 = 1.f - 2.f * s1 * s1; c2 = 1.f - 2.f * s2 * s2;
      s0 = t0; s1 = t1; s2 = t2;
    }
    const float* e = ge + 3 + 6 * kf;
    ax += fr * (e[0] * c0 - e[3] * s0);
    ay += fr * (e[1] * c1 - e[4] * s1);
    az += fr * (e[2] * c2 - e[5] * s2);
    fr *= 2.0f;
  }
  gx.x += ax; gx.y += ay; gx.z += az;
}

// directional derivative of the Fourier embedding along gb (eikonal forward chains): e[0:3] = gb, e[3 + 6k + d] = 2^k cos(2^k x_d) gb_d,
// e[6 + 6k + d] = -2^k sin(2^k x_d) gb_d, zeros up to 64; same double-angle walk as the forward.  e lives in local memory.
__device__ __noinline__ void pe_tangent_fn(const float3& x, int nfreq, const float3& gb, float* e) {
#pragma unroll 1
  for (int i = 0; i < 64; ++i) e[i] = 0.f;
  e[0] = gb.x; e[1] = gb.y; e[2] = gb.z;
  float fr = 1.0f, s0 = 0.f, s1 = 0.f, s2 = 0.f, c0 = 1.f, c1 = 1.f, c2 = 1.f;
#pragma unroll 1
  for (int kf = 0; kf < nfreq; ++kf) {
    if ((kf & 3) == 0) {
      sincosf(fr * x.x, &s0, &c0);
      sincosf(fr * x.y, &s1, &c1);
      sincosf(fr * x.z, &s2, &c2);
    } else {
      const float t0 = 2.f * s0 * c0, t1 = 2.f * s1 * c1, t2 = 2.f * s2 * c2;
      c0 = 1.f - 2.f * s0 * s0; c1 = 1.f - 2.f * s1 * s1; c2 = 1.f - 2.f * s2 * s2;
      s0 = t0; s1 = t1; s2 = t2;
    }
    float* o = e + 3 + 6 * kf;
    o[0] = fr * c0 * gb.x; o[1] = fr * c1 * gb.y; o[2] = fr * c2 * gb.z;
    o[3] = -fr * s0 * gb.x; o[4] = -fr * s1 * gb.y; o[5] = -fr * s2 * gb.z;
    fr *= 2.0f;
  }
}

// EIK: the eikonal instantiation (kernels.h EikParams) - the same producer / issuer / epilogue machinery running the masked
// linear chains of the eikonal term on a list of rays; a separate instantiation, so the field backward's code is untouched.
// WARPONLY: backward of one forward skinning warp (+ soft deformation) of GIVEN points (b200r_warp_bwd: the backward of
// FeatureNeRF.forward_project's warp, nnutils/feature.py:207-226) - the w = 2 iteration of the warp loop below with the points'
// own cotangent; everything else of the field backward is compiled out.
// NORMALS: d sdf / d xyz_cam of every sample (NeRF.compute_normal, nnutils/nerf.py:455-493: the gradient of the sdf through the
// basefield AND the backward warp w.r.t. the camera-space point) - the density chain started from a unit sdf cotangent, the
// w = 0 iteration of the warp loop, the camera rotation; no parameter gradients (b200r_field_normals).
template <class Op, int B, int WIDTH, bool DENSE, bool EIK = false, bool WARPONLY = false, bool NORMALS = false>
__global__ void __launch_bounds__(kThreads, 1) field_bwd_kernel(const __grid_constant__ BwdKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  float* cblk = reinterpret_cast<float*>(ring + kSmemRing);
  float* fblk = cblk + p.prog.cl.n_floats;
  uint64_t* bars = reinterpret_cast<uint64_t*>(fblk + kGroups * p.prog.fl.n_floats);
  uint64_t* full_bar = bars;                    // [group][kNumStages]
  uint64_t* empty_bar = bars + 2 * kNumStages;  // [kNumStages]
  uint64_t* c2m = bars + 3 * kNumStages;        // [group][4]
  uint64_t* m2c = bars + 3 * kNumStages + 8;    // [group][4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kNumStages + 16);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kNumStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&full_bar[kNumStages + i], 1); mbar_init(&empty_bar[i], kCluster); }
    for (int i = 0; i < 8; ++i) { mbar_init(&c2m[i], 4); mbar_init(&m2c[i], 1); }
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();
  tc_fence_after_sync();
  if (*tmem_slot != 0) __trap();
  const Program& P = p.prog;
  const int pair_stride = kGroups * (int)gridDim.x;
  const int iters = (p.n_tiles + pair_stride - 1) / pair_stride;
  const uint32_t cta_rank = kCluster > 1 ? cluster_ctarank() : 0;
  const uint16_t cmask = (uint16_t)((1u << kCluster) - 1);

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsAux));
    if (warp == 8) {
      // =============================================================== TMA producer (as in the forward kernel)
      if (lane == 0) {
        uint32_t stage = 0, phase = 0;
        for (int it = 0; it < iters; ++it) {
          int st = 0;
          while (st < P.n_steps) {
            int end = st;
            while (P.steps[end].commit == 0) ++end;
            ++end;
            for (int g = 0; g < kGroups; ++g) {
              for (int s = st; s < end; ++s) {
                const MmaStep& S = P.steps[s];
                const uint32_t bytes = (uint32_t)S.n * 128u * S.n_sub;
                const uint32_t part = bytes / kCluster;
                uint64_t* fb = &full_bar[g * kNumStages + stage];
                mbar_wait(&empty_bar[stage], phase ^ 1);
                mbar_arrive_expect_tx(fb, bytes);
                const uint8_t* src = p.packed_t + S.w_off + cta_rank * part;
                uint8_t* dst = ring + stage * kWStageBytes + cta_rank * part;
                if (kCluster > 1) tma_bulk_g2s_mcast(dst, src, part, fb, cmask);
                else tma_bulk_g2s(dst, src, part, fb);
                if (++stage == kNumStages) { stage = 0; phase ^= 1; }
              }
            }
            st = end;
          }
        }
      }
    } else if (warp == 9 || warp == 10) {
      // =============================================================== MMA issuers (one per tile group): every operand is TS
      const int g = warp - 9;
      uint32_t stage = 0, full_par = 0, bar_phase = 0;
      uint64_t* full_g = full_bar + g * kNumStages;
      const uint32_t desc_hi = (uint32_t)(umma_desc_k_sw128(0) >> 32);
      const uint32_t bd_lo0 = (uint32_t)umma_desc_k_sw128(smem_u32(ring));
      const uint32_t d = kTmemAcc + kTmemGroup * g, act0 = kTmemAct + kTmemGroup * g;
      uint64_t* c2m_g = c2m + 4 * g;
      uint64_t* m2c_g = m2c + 4 * g;
      auto mk = [&](uint32_t lo) { return ((uint64_t)desc_hi << 32) | lo; };
      auto advance = [&]() { if (++stage == kNumStages) stage = 0; };
      for (int it = 0; it < iters; ++it) {
        int st = 0;
#pragma unroll 1
        while (st < P.n_steps) {
          int end = st;
          while (P.steps[end].commit == 0) ++end;
          ++end;
          const uint32_t cnt = (uint32_t)(end - st);
          if (g == 1) for (uint32_t j = 0; j < cnt; ++j) advance();  // group 0's slots of this block
          const uint32_t wt = P.steps[st].wait, cm = P.steps[end - 1].commit;
          if (wt) {
            mbar_wait(&c2m_g[wt], (bar_phase >> wt) & 1u);
            bar_phase ^= 1u << wt;
          }
          uint32_t a = act0, acc = 0;
#pragma unroll 1
          for (int s = st; s < end; ++s) {
            const MmaStep& S = P.steps[s];
            const uint32_t n = S.n, ks = S.ksteps, ks2 = S.n_sub > 1 ? S.ksteps2 : 0u;
            const uint32_t idesc = umma_idesc_f16(Op::kFmt, 0) | ((n >> 3) << 17);
            mbar_wait(&full_g[stage], (full_par >> stage) & 1u);
            full_par ^= 1u << stage;
            tc_fence_after_sync();
            const uint32_t bd = bd_lo0 + stage * (kWStageBytes >> 4), bd2 = bd + (n << 3);
            if (elect_one()) {
              for (uint32_t k = 0; k < ks; ++k) umma_f16_ts(d, a + 8 * k, mk(bd + 2 * k), idesc, acc | k);
              for (uint32_t k = 0; k < ks2; ++k) umma_f16_ts(d, a + 32 + 8 * k, mk(bd2 + 2 * k), idesc, 1u);
              if (kCluster > 1) umma_commit_mcast(&empty_bar[stage], cmask);
              else umma_commit(&empty_bar[stage]);
              if (s + 1 == end && cm) umma_commit(&m2c_g[cm]);
            }
            __syncwarp();
            advance();
            acc = 1;
            a += S.n_sub > 1 ? 64u : 32u;
          }
          if (g == 0) for (uint32_t j = 0; j < cnt; ++j) advance();
          st = end;
        }
      }
    }
  } else {
    // =============================================================== compute warps: one thread per sample
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsCompute));
    const int g = warp >> 2, q = warp & 3;
    const int gtid = threadIdx.x & (kGroupThreads - 1);
    const uint32_t row = (uint32_t)(q * 32 + lane);
    const uint32_t t_lane = ((uint32_t)(q * 32) << 16);
    const uint32_t tD = t_lane + kTmemAcc + kTmemGroup * g;
    const uint32_t tA = t_lane + kTmemAct + kTmemGroup * g;
    uint64_t* c2m_g = c2m + 4 * g;
    uint64_t* m2c_g = m2c + 4 * g;
    uint32_t all_phase = 0, half_phase = 0;
    constexpr int HN = WIDTH / 2, NBLK = HN / 32;
    const int Dn = p.desc.D;
    const int lid_delta = 0, lid_vis = B > 0 ? 3 : 0, lid_base = lid_vis + 2, lid_rgb0 = lid_base + Dn + 1;
    (void)lid_delta; (void)lid_rgb0;
    const ConstLayout& CL = P.cl;
    const FrameLayout& FL = P.fl;
    const TapeLayout& TL = p.tape;
    float* fblk_g = fblk + g * FL.n_floats;
    const uint32_t cblk_s = smem_u32(cblk), fblk_s = smem_u32(fblk_g);
    const uint32_t rowx = row * 128u + ((row & 7u) << 4);
    const uint32_t sc_s = cblk_s + 4u * CL.scalars;
    if (p.workspace) {  // the eikonal modes read no block
      const float4* src = reinterpret_cast<const float4*>(p.workspace);
      float4* dst = reinterpret_cast<float4*>(cblk);
      for (int i = threadIdx.x; i < CL.n_floats / 4; i += kComputeThreads) dst[i] = __ldg(src + i);
    }
    named_bar_sync(3, kComputeThreads);
    const float S = p.scale ? __ldg(p.scale) : 1.0f;

    auto warp_arrive = [&](uint64_t* bar) {
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    auto arrive_all = [&]() {
      tc_fence_before_sync();
      warp_arrive(&c2m_g[BAR_ALL]);
    };
    auto wait_all = [&]() {
      mbar_wait(&m2c_g[BAR_ALL], all_phase);
      all_phase ^= 1;
      tc_fence_after_sync();
    };
    auto gemm = [&]() { arrive_all(); wait_all(); };
    auto wait_half = [&](int nh) {
      mbar_wait(&m2c_g[BAR_H0 + nh], (half_phase >> nh) & 1u);
      half_phase ^= 1u << nh;
      tc_fence_after_sync();
    };

    // per-tile pointers
    uint8_t* gt_tile = nullptr;          // this tile's gradient chunks
    int n_gt = TL.n_g;                   // chunks per tile of the tape being written (the eikonal modes have their own)
    const uint8_t* at_tile = nullptr;    // this tile's forward chunks
    const uint32_t* mask_row = nullptr;  // this row's ReLU sign words
    auto gtape_st32 = [&](int chunk, int col0, const uint32_t (&o)[16]) {
      chunk_st32(gt_tile + tape_row_off(n_gt, chunk + (col0 >> 6), row), row, (uint32_t)(col0 & 63) >> 3, o);
    };
    auto gtape_zero_row = [&](int chunk) {
      uint8_t* base = gt_tile + tape_row_off(n_gt, chunk, row);
#pragma unroll
      for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(base + 16 * j) = make_uint4(0u, 0u, 0u, 0u);
    };
    // 32 gradient values * relu' (sign word: pair i -> bits 15-i / 31-i, 1 = inactive) -> 16 packed registers
    auto mask_pack32 = [&](const float (&v)[32], uint32_t w, uint32_t (&o)[16]) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a = (w & (1u << (15 - i))) ? 0.f : v[2 * i], b = (w & (1u << (31 - i))) ? 0.f : v[2 * i + 1];
        o[i] = Op::pack2_sat(a, b);
      }
    };
    // finished GEMM of n (<= 128) columns: G = acc * relu'(slot) -> activations [0, n) and the gradient tape
    auto seq_dgrad = [&](int n, int mask_slot, int save_chunk) {
      const uint4 mw = __ldg(reinterpret_cast<const uint4*>(mask_row + (size_t)mask_slot * (kTileRows * kMaskWords)));
      const uint32_t mws[4] = {mw.x, mw.y, mw.z, mw.w};
#pragma unroll 1
      for (int blk = 0; blk < (n >> 5); ++blk) {
        float v[32];
        uint32_t o[16];
        tmem_ld32(tD + 32 * blk, v);
        mask_pack32(v, mws[blk], o);
        tmem_st16(tA + 16 * blk, o);
        gtape_st32(save_chunk, 32 * blk, o);
      }
      tmem_st_wait();
    };
    // One WIDTH-wide dgrad issued as two N-halves (program.h pipe): G = acc * relu'(mask_slot) -> activations (in place,
    // half 0 held in registers until the layer's MMAs have read their input) and the gradient tape.
    //   MODE 1 (rgb.0's input = base + colour features): the colour branch takes relu'(colorfield.linear_final) -> activations;
    //   the density branch adds the sdf head's gradient, takes relu'(basefield.linear_final) and goes to the tape only.
    float g_sdf = 0.f;
    uint4 pm_a = make_uint4(0u, 0u, 0u, 0u), pm_b = pm_a;  // sign words requested ahead for the next wide layer
    auto prefetch_mask = [&](int slot) {
      if (slot < 0) return;
      const uint4* mp = reinterpret_cast<const uint4*>(mask_row + (size_t)slot * (kTileRows * kMaskWords));
      pm_a = __ldg(mp);
      if (NBLK > 2) pm_b = __ldg(mp + 1);
    };
    auto wide_dgrad = [&](auto mode_tag, int mask_slot, int save_chunk, int next_slot, bool last = false) {
      constexpr int MODE = decltype(mode_tag)::value;
      uint32_t hold[NBLK][16];
      (void)mask_slot;  // its words were requested by the previous layer (prefetch_mask)
      const uint32_t mws[8] = {pm_a.x, pm_a.y, pm_a.z, pm_a.w, pm_b.x, pm_b.y, pm_b.z, pm_b.w};
      prefetch_mask(next_slot);
      uint32_t mw2[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
      if (MODE == 1) {
        const uint4* mp2 = reinterpret_cast<const uint4*>(mask_row + (size_t)TL.m_base[Dn] * (kTileRows * kMaskWords));
        const uint4 a2 = __ldg(mp2), b2 = NBLK > 2 ? __ldg(mp2 + 1) : make_uint4(0u, 0u, 0u, 0u);
        mw2[0] = a2.x; mw2[1] = a2.y; mw2[2] = a2.z; mw2[3] = a2.w; mw2[4] = b2.x; mw2[5] = b2.y; mw2[6] = b2.z; mw2[7] = b2.w;
      }
      auto math = [&](float (&v)[32], int wi, int col0, uint32_t (&o)[16]) {
        mask_pack32(v, mws[wi], o);
        gtape_st32(save_chunk, col0, o);
        if (MODE == 1) {
          const uint32_t wa = cblk_s + 4u * (CL.sdf_w + col0);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 w4 = lds128(wa + 4u * j);
            v[j] += g_sdf * w4.x; v[j + 1] += g_sdf * w4.y; v[j + 2] += g_sdf * w4.z; v[j + 3] += g_sdf * w4.w;
          }
          uint32_t o2[16];
          mask_pack32(v, mw2[wi], o2);
          gtape_st32(TL.g_base[Dn], col0, o2);
        }
      };
      wait_half(0);
#pragma unroll
      for (int bp = 0; bp < NBLK; bp += 2) {  // two 32-column blocks per TMEM round trip
        uint32_t rp[2][32];
        tmem_ld32_issue(tD + 32 * bp, rp[0]);
        tmem_ld32_issue(tD + 32 * (bp + 1), rp[1]);
        tmem_ld_wait32(rp[0]);
        tmem_ld_wait32(rp[1]);
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rp[h2][j]);
          math(v, bp + h2, 32 * (bp + h2), hold[bp + h2]);
        }
      }
      tc_fence_before_sync();
      warp_arrive(&c2m_g[BAR_H0]);
      wait_half(1);
#pragma unroll
      for (int blk = 0; blk < NBLK; ++blk) tmem_st16(tA + 16 * blk, hold[blk]);
#pragma unroll
      for (int bp = 0; bp < NBLK; bp += 2) {
        uint32_t rp[2][32];
        tmem_ld32_issue(tD + 32 * bp, rp[0]);
        tmem_ld32_issue(tD + 32 * (bp + 1), rp[1]);
        tmem_ld_wait32(rp[0]);
        tmem_ld_wait32(rp[1]);
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int blk = bp + h2;
          float v[32];
          uint32_t o[16];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rp[h2][j]);
          math(v, NBLK + blk, HN + 32 * blk, o);
          tmem_st16(tA + (HN >> 1) + 16 * blk, o);
        }
      }
      tmem_st_wait();
      tc_fence_before_sync();
      if (!last) warp_arrive(&c2m_g[BAR_H1]);  // `last`: no block of this tile is left to consume the arrival
    };
    // gradient rows of a tape chunk range -> activations (the density chain starts from what the rgb.0 epilogue parked)
    auto load_g_to_act = [&](int chunk, int ncols) {
#pragma unroll 1
      for (int blk = 0; blk < (ncols >> 5); ++blk) {
        const uint8_t* base = gt_tile + tape_row_off(TL.n_g, chunk + (blk >> 1), row);
        uint32_t o[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 t = *reinterpret_cast<const uint4*>(base + ((((uint32_t)(4 * (blk & 1) + j)) ^ (row & 7u)) << 4));
          o[4 * j] = t.x; o[4 * j + 1] = t.y; o[4 * j + 2] = t.z; o[4 * j + 3] = t.w;
        }
        tmem_st16(tA + 16 * blk, o);
      }
      tmem_st_wait();
    };
    // backward of the Fourier embedding (nnutils/embedding.py:69-125): accumulator columns [0, 3 + 6 nfreq) hold dL/d e;
    // returns dL/dx = g_e[0:3] + sum_k 2^k (g_sin_k cos(2^k x) - g_cos_k sin(2^k x)), same double-angle walk as the forward
    auto pe_backward = [&](const float3& x, int nfreq, float3& gx) { pe_backward_fn(tD, x, nfreq, gx); };

    for (int it = 0; it < iters; ++it) {
      const int tile_raw = (kGroups * it + g) * (int)gridDim.x + (int)blockIdx.x;
      const bool dead_tile = tile_raw >= p.n_tiles;
      const int tile = dead_tile ? p.n_tiles - 1 : tile_raw;
      if constexpr (EIK) {
        // ================================================================ eikonal chains (nnutils/nerf.py:416-453)
        const EikParams& E = p.eik;
        const int pi_raw = tile * kTileRows + (int)row;
        const bool live_p = !dead_tile && pi_raw < E.n_points;
        const int pi = pi_raw < E.n_points ? pi_raw : E.n_points - 1;
        const int DD = p.rays.D;
        const int rsel = pi / DD, kk = pi - rsel * DD;
        const int ray = __ldg(E.rays_sel + rsel);       // f * N + n in the training forward's batch
        const int fm = ray / p.rays.N, r_in = (ray - fm * p.rays.N) * DD + kk;
        const size_t s = (size_t)fm * p.ND + r_in;
        // the sample's ReLU sign words sit in the training tape's tile (frame, r_in / 128), row r_in % 128
        mask_row = p.tape_mask + ((size_t)(fm * p.tiles_per_frame + r_in / kTileRows) * TL.n_mask * kTileRows + (uint32_t)(r_in % kTileRows)) * kMaskWords;
        n_gt = E.n_chunks;
        gt_tile = E.tape + (size_t)(dead_tile ? p.n_tiles + (int)blockIdx.x : tile) * E.n_chunks * kChunkBytes;
        const float3 x = make_float3(__ldg(p.saved.xyz + s * 3), __ldg(p.saved.xyz + s * 3 + 1), __ldg(p.saved.xyz + s * 3 + 2));
        if (E.mode == 1) {
          // ---- reverse chain: a_F = relu'(linear_final) * (scale * w_sdf) -> activations and tape, then the density chain's
          // data-gradient GEMMs; the embedding columns come back as g = E(x)^T u
          const float Sa = live_p ? E.scale_a : 0.f;
          gtape_zero_row(E.head_chunk);
          *reinterpret_cast<uint4*>(gt_tile + tape_row_off(n_gt, E.head_chunk, row) + ((row & 7u) << 4)) = make_uint4(0u, Op::pack2_sat(0.f, Sa), 0u, 0u);
          {
            const uint4* mp = reinterpret_cast<const uint4*>(mask_row + (size_t)TL.m_base[Dn] * (kTileRows * kMaskWords));
            const uint4 ma = __ldg(mp), mb = NBLK > 2 ? __ldg(mp + 1) : make_uint4(0u, 0u, 0u, 0u);
            const uint32_t mws[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
#pragma unroll 1
            for (int blk = 0; blk < WIDTH / 32; ++blk) {
              float v[32];
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(E.sdf_w + 32 * blk + j));
                v[j] = Sa * w4.x; v[j + 1] = Sa * w4.y; v[j + 2] = Sa * w4.z; v[j + 3] = Sa * w4.w;
              }
              uint32_t o[16];
              mask_pack32(v, mws[blk], o);
              tmem_st16(tA + 16 * blk, o);
              gtape_st32(E.out_chunk[Dn], 32 * blk, o);
            }
            tmem_st_wait();
          }
          prefetch_mask(TL.m_base[Dn - 1]);
          float3 gx = make_float3(0.f, 0.f, 0.f);
          arrive_all();
#pragma unroll 1
          for (int i = Dn; i >= 1; --i) {
            if (i == p.desc.skip) {
              wait_all();
              pe_backward(x, p.desc.L_xyz, gx);
              arrive_all();
            }
            wide_dgrad(std::integral_constant<int, 0>{}, TL.m_base[i - 1], E.out_chunk[i - 1], i >= 2 ? TL.m_base[i - 2] : -1);
          }
          wait_all();
          pe_backward(x, p.desc.L_xyz, gx);
          if (live_p) {
            const float inv = 1.0f / E.scale_a;
            E.g_out[(size_t)pi * 3] = gx.x * inv; E.g_out[(size_t)pi * 3 + 1] = gx.y * inv; E.g_out[(size_t)pi * 3 + 2] = gx.z * inv;
          }
        } else {
          // ---- forward chain A (mode 2) / B (mode 3): v_0 = E(x) (scale * gbar) -> 64 operand columns, then the layers
          const float Sv = live_p ? S : 0.f;
          const float3 gb = make_float3(Sv * __ldg(E.gbar + (size_t)pi * 3), Sv * __ldg(E.gbar + (size_t)pi * 3 + 1), Sv * __ldg(E.gbar + (size_t)pi * 3 + 2));
          float e64[64];
          pe_tangent_fn(x, p.desc.L_xyz, gb, e64);
#pragma unroll
          for (int hb = 0; hb < 2; ++hb) {
            uint32_t o[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = Op::pack2_sat(e64[32 * hb + 2 * i], e64[32 * hb + 2 * i + 1]);
            tmem_st16(tA + 16 * hb, o);
            if (E.v0_chunk >= 0) gtape_st32(E.v0_chunk, 32 * hb, o);
          }
          tmem_st_wait();
          prefetch_mask(E.mask_slot[0]);
          arrive_all();
#pragma unroll 1
          for (int l = 0; l < E.n_layers; ++l)
            wide_dgrad(std::integral_constant<int, 0>{}, E.mask_slot[l], E.out_chunk[l], l + 1 < E.n_layers ? E.mask_slot[l + 1] : -1, l + 1 == E.n_layers);
        }
        continue;
      }
      const int f = tile / p.tiles_per_frame;
      const int r_raw = (tile - f * p.tiles_per_frame) * kTileRows + (int)row;
      const bool live = !dead_tile && r_raw < p.ND;
      const int r_in = r_raw < p.ND ? r_raw : p.ND - 1;
      const int n = r_in / p.rays.D;
      const int k = r_in - n * p.rays.D;
      const size_t s = (size_t)f * p.ND + r_in;
      // a dead tile (its partner group still has a real one) walks the same protocol on a scratch tile of the gradient tape
      gt_tile = p.tape_g + (size_t)(dead_tile ? p.n_tiles + (int)blockIdx.x : tile) * TL.n_g * kChunkBytes;
      at_tile = p.tape_a + (size_t)tile * TL.n_a * kChunkBytes;
      mask_row = p.tape_mask + ((size_t)tile * TL.n_mask * kTileRows + row) * kMaskWords;

      named_bar_sync(1 + g, kGroupThreads);
      {
        const float4* src = reinterpret_cast<const float4*>(p.workspace + CL.n_floats + (size_t)f * FL.n_floats);
        float4* dst = reinterpret_cast<float4*>(fblk_g);
        for (int i = gtid; i < FL.n_floats / 4; i += kGroupThreads) dst[i] = __ldg(src + i);
      }
      named_bar_sync(1 + g, kGroupThreads);

      // ------------------------------------------------ geometry of the sample (as the forward)
      const float* hx = WARPONLY ? nullptr : p.rays.hxy + ((size_t)f * p.rays.N + n) * 3;  // point entries have no rays
      const float h0 = WARPONLY ? 0.f : __ldg(hx), h1 = WARPONLY ? 0.f : __ldg(hx + 1), h2 = WARPONLY ? 1.f : __ldg(hx + 2);
      const float* cam = fblk_g + FL.cam;
      const float3 dvec = make_float3(h0 * cam[0] + h1 * cam[1] + h2 * cam[2], h0 * cam[3] + h1 * cam[4] + h2 * cam[5],
                                      h0 * cam[6] + h1 * cam[7] + h2 * cam[8]);
      const float dn = sqrtf(dvec.x * dvec.x + dvec.y * dvec.y + dvec.z * dvec.z);
      float depth;
      {
        const float nearv = cam[9], farv = cam[10];
        const int DD = p.rays.D;
        const float step = 1.0f / (float)(DD - 1);
        const float z = k < DD / 2 ? step * (float)k : 1.0f - step * (float)(DD - 1 - k);
        depth = p.rays.depth ? __ldg(p.rays.depth + ((size_t)f * p.rays.N + n) * DD + k) : nearv * (1.0f - z) + farv * z;
      }
      const float3 xyz_cam = make_float3(dvec.x * depth, dvec.y * depth, dvec.z * depth);
      const float3 dir_cam = make_float3(dvec.x / dn, dvec.y / dn, dvec.z / dn);
      const Q4 qc = {cam[11], cam[12], cam[13], cam[14]};
      const Q4 qi = qconj(qc);
      const float3 ti = qrot(qi, make_float3(-cam[15], -cam[16], -cam[17]));
      float3 xyz_t = qrot(qi, xyz_cam);
      xyz_t.x += ti.x; xyz_t.y += ti.y; xyz_t.z += ti.z;
      float3 xyz = xyz_t;
      if (B > 0) xyz = make_float3(__ldg(p.saved.xyz + s * 3), __ldg(p.saved.xyz + s * 3 + 1), __ldg(p.saved.xyz + s * 3 + 2));

      // cotangents of the per-sample outputs, scaled; dead rows contribute nothing
      const float Sl = live ? S : 0.f;
      auto ld1 = [&](const float* ptr) { return ptr ? Sl * __ldg(ptr + s) : 0.f; };
      auto ld3 = [&](const float* ptr) { return ptr ? make_float3(Sl * __ldg(ptr + s * 3), Sl * __ldg(ptr + s * 3 + 1), Sl * __ldg(ptr + s * 3 + 2)) : make_float3(0.f, 0.f, 0.f); };
      float3 g_xyz = ld3(p.g.xyz);
      // per-frame / global sums of this row: reduced over the tile at the end
      //  0-8 Kinv, 9-12 qi, 13-15 ti, 16-19 partner Kinv[0],[2],[4],[5], 20-23 partner q, 24-26 partner t,
      //  27 logibeta, 28 warp.logibeta, 29 logscale
      float red[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) red[i] = 0.f;

      if constexpr (NORMALS) {
        // ================================================================ unit cotangent on the sdf -> density chain -> dL/d xyz
        const float Sa = live ? p.eik.scale_a : 0.f;
        {
          const uint4* mp = reinterpret_cast<const uint4*>(mask_row + (size_t)TL.m_base[Dn] * (kTileRows * kMaskWords));
          const uint4 ma = __ldg(mp), mb = NBLK > 2 ? __ldg(mp + 1) : make_uint4(0u, 0u, 0u, 0u);
          const uint32_t mws[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
#pragma unroll 1
          for (int blk = 0; blk < WIDTH / 32; ++blk) {
            float v[32];
            const uint32_t wa = cblk_s + 4u * (CL.sdf_w + 32 * blk);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 w4 = lds128(wa + 4u * j);
              v[j] = Sa * w4.x; v[j + 1] = Sa * w4.y; v[j + 2] = Sa * w4.z; v[j + 3] = Sa * w4.w;
            }
            uint32_t o[16];
            mask_pack32(v, mws[blk], o);
            tmem_st16(tA + 16 * blk, o);
            gtape_st32(TL.g_base[Dn], 32 * blk, o);
          }
          tmem_st_wait();
        }
        prefetch_mask(TL.m_base[Dn - 1]);
        arrive_all();
#pragma unroll 1
        for (int i = Dn; i >= 1; --i) {
          if (i == p.desc.skip) {
            wait_all();
            pe_backward(xyz, p.desc.L_xyz, g_xyz);
            arrive_all();
          }
          wide_dgrad(std::integral_constant<int, 0>{}, TL.m_base[i - 1], TL.g_base[i - 1], i >= 2 ? TL.m_base[i - 2] : -1);
        }
        wait_all();
        pe_backward(xyz, p.desc.L_xyz, g_xyz);
      } else if constexpr (!WARPONLY) {
      // ================================================================ rgb head, rgb.0, colour chain, density chain
      {
        prefetch_mask(TL.m_col[2]);
        const float3 g_rgb = ld3(p.g.rgb);
        const float r0 = __ldg(p.saved.rgb + s * 3), r1 = __ldg(p.saved.rgb + s * 3 + 1), r2 = __ldg(p.saved.rgb + s * 3 + 2);
        const float go0 = g_rgb.x * r0 * (1.f - r0), go1 = g_rgb.y * r1 * (1.f - r1), go2 = g_rgb.z * r2 * (1.f - r2);
        // density = (0.5 + 0.5 sign(s) expm1(-|s| ibeta)) ibeta
        const float sdf = __ldg(p.saved.sdf + s), ibeta = lds32(sc_s + 4u * SC_IBETA);
        const float g_den = ld1(p.g.density);
        const float ex = __expf(-fabsf(sdf) * ibeta), sgn = sdf > 0.f ? 1.f : (sdf < 0.f ? -1.f : 0.f);
        g_sdf = g_den * (-0.5f * ibeta * ibeta * ex);
        red[27] = g_den * ibeta * ((0.5f + 0.5f * sgn * expm1f(-fabsf(sdf) * ibeta)) + ibeta * 0.5f * sgn * ex * (-fabsf(sdf)));
        const float g_vis = ld1(p.g.vis);
        {  // head chunk: columns 3 sdf, 4-6 rgb pre-sigmoid, 7 visibility logit (the order of the scalars block's biases)
          gtape_zero_row(TL.g_head);
          uint8_t* base = gt_tile + tape_row_off(TL.n_g, TL.g_head, row);
          *reinterpret_cast<uint4*>(base + ((row & 7u) << 4)) = make_uint4(0u, Op::pack2_sat(0.f, g_sdf), Op::pack2_sat(go0, go1), Op::pack2_sat(go2, g_vis));
        }
        // G of rgb.0's pre-activation: (g_o W2) * relu'
        const uint32_t w2 = cblk_s + 4u * CL.rgb2_w;
        const uint4 mw = __ldg(reinterpret_cast<const uint4*>(mask_row + (size_t)TL.m_rgb0 * (kTileRows * kMaskWords)));
        const uint32_t mws[4] = {mw.x, mw.y, mw.z, mw.w};
        float3 g_dir = make_float3(0.f, 0.f, 0.f);
#pragma unroll 1
        for (int blk = 0; blk < HN / 32; ++blk) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 wr = lds128(w2 + 4u * (32 * blk + j)), wg = lds128(w2 + 4u * (HN + 32 * blk + j)), wb = lds128(w2 + 4u * (2 * HN + 32 * blk + j));
            v[j] = go0 * wr.x + go1 * wg.x + go2 * wb.x;
            v[j + 1] = go0 * wr.y + go1 * wg.y + go2 * wb.y;
            v[j + 2] = go0 * wr.z + go1 * wg.z + go2 * wb.z;
            v[j + 3] = go0 * wr.w + go1 * wg.w + go2 * wb.w;
          }
          uint32_t o[16];
          mask_pack32(v, mws[blk], o);
          tmem_st16(tA + 16 * blk, o);
          gtape_st32(TL.g_rgb0, 32 * blk, o);
          if (p.desc.L_dir == 0) {  // raw view direction columns of rgb.0 (fp32 SIMT in the forward)
            const uint32_t wd = cblk_s + 4u * CL.dir_w;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float2 gz = Op::unpack2(o[i]);
              const uint32_t da = wd + 12u * (32 * blk + 2 * i);
              g_dir.x += gz.x * lds32(da) + gz.y * lds32(da + 12); g_dir.y += gz.x * lds32(da + 4) + gz.y * lds32(da + 16);
              g_dir.z += gz.x * lds32(da + 8) + gz.y * lds32(da + 20);
            }
          }
        }
        tmem_st_wait();
        if (p.desc.L_dir == 0) {  // dir_f = R(qi) dir_cam, dir_cam = d / |d|
          Q4 gq; float3 gdc;
          qrot_bwd(qi, dir_cam, g_dir, gq, gdc);
          red[9] += gq.w; red[10] += gq.x; red[11] += gq.y; red[12] += gq.z;
          const float dt = gdc.x * dir_cam.x + gdc.y * dir_cam.y + gdc.z * dir_cam.z;
          const float3 gd = make_float3((gdc.x - dir_cam.x * dt) / dn, (gdc.y - dir_cam.y * dt) / dn, (gdc.z - dir_cam.z * dt) / dn);
          red[0] += gd.x * h0; red[1] += gd.x * h1; red[2] += gd.x * h2; red[3] += gd.y * h0; red[4] += gd.y * h1; red[5] += gd.y * h2;
          red[6] += gd.z * h0; red[7] += gd.z * h1; red[8] += gd.z * h2;
        }
        arrive_all();
        wide_dgrad(std::integral_constant<int, 1>{}, TL.m_col[2], TL.g_col[2], TL.m_col[1]);   // through rgb.0: colour + density branches
        wide_dgrad(std::integral_constant<int, 0>{}, TL.m_col[1], TL.g_col[1], TL.m_col[0]);   // colorfield.linear_final
        wide_dgrad(std::integral_constant<int, 0>{}, TL.m_col[0], TL.g_col[0], TL.m_base[Dn - 1]);  // colorfield.linear_2
        wait_all();                                                                 // colorfield.linear_1 -> embedding columns
        pe_backward(xyz, p.desc.L_xyz + 2, g_xyz);
        // density chain
        load_g_to_act(TL.g_base[Dn], WIDTH);
        arrive_all();
#pragma unroll 1
        for (int i = Dn; i >= 1; --i) {
          if (i == p.desc.skip) {  // the skip layer also feeds the embedding: its columns come back first
            wait_all();
            pe_backward(xyz, p.desc.L_xyz, g_xyz);
            arrive_all();
          }
          wide_dgrad(std::integral_constant<int, 0>{}, TL.m_base[i - 1], TL.g_base[i - 1], i >= 2 ? TL.m_base[i - 2] : -1);
        }
        wait_all();
        pe_backward(xyz, p.desc.L_xyz, g_xyz);
      }

      // ================================================================ feature field
      if (p.desc.has_feature) {
        float o16[16], gf[16];
        const float inv = __ldg(p.saved.feat_norm + s);
        float dt = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          o16[j] = __ldg(p.saved.feature + s * 16 + j);
          gf[j] = p.g.feature ? Sl * __ldg(p.g.feature + s * 16 + j) : 0.f;
          dt += gf[j] * o16[j];
        }
        uint32_t o[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = Op::pack2_sat((gf[2 * j] - o16[2 * j] * dt) * inv, (gf[2 * j + 1] - o16[2 * j + 1] * dt) * inv);
#pragma unroll
        for (int j = 8; j < 16; ++j) o[j] = 0u;
        tmem_st16(tA, o);
        tmem_st_wait();
        gtape_st32(TL.g_feat[5], 0, o);
        gemm();
        seq_dgrad(128, TL.m_feat[4], TL.g_feat[4]);
        gemm();
        pe_backward(xyz, 6, g_xyz);
        gemm();
        seq_dgrad(128, TL.m_feat[3], TL.g_feat[3]);
#pragma unroll 1
        for (int i = 3; i >= 1; --i) {
          gemm();
          seq_dgrad(128, TL.m_feat[i - 1], TL.g_feat[i - 1]);
        }
        gemm();
        pe_backward(xyz, 6, g_xyz);
      }

      // ================================================================ visibility MLP
      {
        const float g_vis = ld1(p.g.vis);
        const uint32_t vw = cblk_s + 4u * CL.vis_w;
        const uint2 mw = __ldg(reinterpret_cast<const uint2*>(mask_row + (size_t)TL.m_vis[1] * (kTileRows * kMaskWords)));
        const uint32_t mws[2] = {mw.x, mw.y};
#pragma unroll 1
        for (int blk = 0; blk < 2; ++blk) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 w4 = lds128(vw + 4u * (32 * blk + j));
            v[j] = g_vis * w4.x; v[j + 1] = g_vis * w4.y; v[j + 2] = g_vis * w4.z; v[j + 3] = g_vis * w4.w;
          }
          uint32_t o[16];
          mask_pack32(v, mws[blk], o);
          tmem_st16(tA + 16 * blk, o);
          gtape_st32(TL.g_vis[1], 32 * blk, o);
        }
        tmem_st_wait();
        gemm();
        seq_dgrad(64, TL.m_vis[0], TL.g_vis[0]);
        gemm();
        pe_backward(xyz, 10, g_xyz);
      }

      }  // !WARPONLY
      float3 g_xyz_t = make_float3(0.f, 0.f, 0.f);
      // DenseWarp.forward backward (nnutils/warping.py:143-170): x' = x + 0.1 CondMLP([PE6(x), t, inst]); map m = 0 forward_map,
      // 1 backward_map; w = tape slot of the stage.  Returns dL/dx for the cotangent g_out of x'.
      auto dense_backward = [&](int m, int w, const float3& x_in, const float3& g_out) -> float3 {
        const float gm0 = 0.1f * g_out.x, gm1 = 0.1f * g_out.y, gm2 = 0.1f * g_out.z;
        gtape_zero_row(TL.g_d3[w]);
        *reinterpret_cast<uint2*>(gt_tile + tape_row_off(TL.n_g, TL.g_d3[w], row) + ((row & 7u) << 4)) = make_uint2(Op::pack2_sat(gm0, gm1), Op::pack2_sat(gm2, 0.f));
        const float* w3 = p.dense_w3[m];  // (3, 256) head weight, read through L1 (every row reads the same addresses)
        const uint4 ma = __ldg(reinterpret_cast<const uint4*>(mask_row + (size_t)TL.m_dh2[w] * (kTileRows * kMaskWords)));
        const uint4 mb = __ldg(reinterpret_cast<const uint4*>(mask_row + (size_t)TL.m_dh2[w] * (kTileRows * kMaskWords)) + 1);
        const uint32_t mws[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
#pragma unroll 1
        for (int blk = 0; blk < 8; ++blk) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(w3 + 32 * blk + j)), b = __ldg(reinterpret_cast<const float4*>(w3 + 256 + 32 * blk + j)),
                         c = __ldg(reinterpret_cast<const float4*>(w3 + 512 + 32 * blk + j));
            v[j] = gm0 * a.x + gm1 * b.x + gm2 * c.x; v[j + 1] = gm0 * a.y + gm1 * b.y + gm2 * c.y;
            v[j + 2] = gm0 * a.z + gm1 * b.z + gm2 * c.z; v[j + 3] = gm0 * a.w + gm1 * b.w + gm2 * c.w;
          }
          uint32_t o[16];
          mask_pack32(v, mws[blk], o);
          tmem_st16(tA + 16 * blk, o);
          gtape_st32(TL.g_d2[w], 32 * blk, o);
        }
        tmem_st_wait();
        arrive_all();
        prefetch_mask(TL.m_dh1[w]);
        wide_dgrad(std::integral_constant<int, 0>{}, TL.m_dh1[w], TL.g_d1[w], -1);  // linear_2
        wait_all();                                                                // linear_1 -> embedding columns
        float3 g_in = g_out;
        pe_backward(x_in, 6, g_in);
        return g_in;
      };
      // partner camera: flow = project(K', R(qn) x_next + tn) - hxy  (nnutils/nerf.py:948-997)
      auto flow_backward = [&](const float3& x_next) -> float3 {
        const float* cn = fblk_g + FL.cam_partner;
        const Q4 qn = {cn[11], cn[12], cn[13], cn[14]};
        float3 xc = qrot(qn, x_next);
        xc.x += cn[15]; xc.y += cn[16]; xc.z += cn[17];
        const float k0 = cn[0], k1 = cn[4], k2 = cn[2], k3 = cn[5];
        const float fx = 1.0f / k0, fy = 1.0f / k1, cx = -k2 / k0, cy = -k3 / k1;
        const float hz = xc.z + 1e-6f, nx = fx * xc.x + cx * xc.z, ny = fy * xc.y + cy * xc.z;
        const float3 gfl = ld3(p.g.flow);
        const float gnx = gfl.x / hz, gny = gfl.y / hz, ghz = -(gfl.x * nx + gfl.y * ny) / (hz * hz);
        const float3 g_xc = make_float3(gnx * fx, gny * fy, gnx * cx + gny * cy + ghz);
        const float gfx = gnx * xc.x, gcx = gnx * xc.z, gfy = gny * xc.y, gcy = gny * xc.z;
        red[16] += -gfx / (k0 * k0) + gcx * k2 / (k0 * k0);  // Kinv'[0]
        red[17] += -gcx / k0;                                // Kinv'[2]
        red[18] += -gfy / (k1 * k1) + gcy * k3 / (k1 * k1);  // Kinv'[4]
        red[19] += -gcy / k1;                                // Kinv'[5]
        red[24] += g_xc.x; red[25] += g_xc.y; red[26] += g_xc.z;
        Q4 gq; float3 gx;
        qrot_bwd(qn, x_next, g_xc, gq, gx);
        red[20] += gq.w; red[21] += gq.x; red[22] += gq.y; red[23] += gq.z;
        return gx;
      };

      if constexpr (B > 0) {
        // ================================================================ Gaussian bone density
        if constexpr (!WARPONLY && !NORMALS) {
          float best = INFINITY;
          int sel = 0;
          const uint32_t ctr = cblk_s + 4u * CL.center;
#pragma unroll 5
          for (int b = 0; b < B; ++b) {
            const float4 c = lds128(ctr + 16u * b);
            const float dx = xyz.x - c.x, dy = xyz.y - c.y, dz = xyz.z - c.z;
            const float d2 = dx * dx + dy * dy + dz * dz;
            if (d2 < best) { best = d2; sel = b; }
          }
          const float ggd = ld1(p.g.gauss_density) * expf(-0.5f * (best / (0.01f * 0.01f))) * lds32(sc_s + 4u * SC_WARP_IBETA);
          const float4 c = lds128(ctr + 16u * sel);
          const float kx = ggd * (xyz.x - c.x) * 1e4f, ky = ggd * (xyz.y - c.y) * 1e4f, kz = ggd * (xyz.z - c.z) * 1e4f;
          g_xyz.x -= kx; g_xyz.y -= ky; g_xyz.z -= kz;
          red[28] = ggd;
          if (ggd != 0.f) {
            float* gc = p.g_cblk + CL.center + 4 * sel;
            atomicAdd(gc, kx / S); atomicAdd(gc + 1, ky / S); atomicAdd(gc + 2, kz / S);
          }
        }
        // ================================================================ skinning warps, last first
        const float g_ent = 0.5f * ld1(p.g.skin_entropy), g_dsk = 0.5f * ld1(p.g.delta_skin), g_cyc = ld1(p.g.cyc_dist);
        float3 x_soft[3] = {xyz, xyz, xyz};  // ComposedWarp: [skinned point before the soft deformation, flow input, cycle input]
        if constexpr (DENSE) {
#pragma unroll
          for (int i = 0; i < 3; ++i) x_soft[i] = make_float3(__ldg(p.saved.warp_pts + s * 9 + 3 * i), __ldg(p.saved.warp_pts + s * 9 + 3 * i + 1), __ldg(p.saved.warp_pts + s * 9 + 3 * i + 2));
        }
#pragma unroll 1
        for (int w = NORMALS ? 0 : 2; w >= (WARPONLY ? 2 : 0); --w) {
          const float3 x = w == 0 ? xyz_t : (DENSE ? x_soft[w] : xyz);
          const uint32_t binv = fblk_s + 4u * (w == 0 ? FL.binv_t : (w == 1 ? FL.binv_rest_partner : FL.binv_rest));
          const uint32_t se3 = fblk_s + 4u * (w == 0 ? FL.se3_bwd : (w == 1 ? FL.se3_fwd_partner : FL.se3_fwd));
          // ---- recompute the blend (SkinningWarp.forward) from the taped delta-MLP outputs
          float lw[B];   // logits -> softmax weights
          float zr[B];   // raw delta-MLP outputs
          {
            const uint8_t* zb = at_tile + tape_row_off(TL.n_a, TL.a_z[w], row);
#pragma unroll
            for (int j = 0; j < (B + 7) / 8; ++j) {
              const uint4 t = __ldg(reinterpret_cast<const uint4*>(zb + (((uint32_t)j ^ (row & 7u)) << 4)));
              const uint32_t tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float2 fz = Op::unpack2(tw[u]);
                if (8 * j + 2 * u < B) zr[8 * j + 2 * u] = fz.x;
                if (8 * j + 2 * u + 1 < B) zr[8 * j + 2 * u + 1] = fz.y;
              }
            }
          }
          float mx = -INFINITY;
          int amax = 0;
#pragma unroll
          for (int b = 0; b < B; ++b) {
            const uint32_t ba = binv + 48u * b;
            const float4 r0 = lds128(ba), r1 = lds128(ba + 16), r2 = lds128(ba + 32);
            const float v0 = r0.x * x.x + r0.y * x.y + r0.z * x.z + r0.w, v1 = r1.x * x.x + r1.y * x.y + r1.z * x.z + r1.w,
                        v2 = r2.x * x.x + r2.y * x.y + r2.z * x.z + r2.w;
            const float lg = -(v0 * v0 + v1 * v1 + v2 * v2 + 0.1f * fmaxf(zr[b], 0.f));
            lw[b] = lg;
            if (lg > mx) { mx = lg; amax = b; }
          }
          const float4 qa = lds128(se3 + 32u * amax);
          float se = 0.f;
#pragma unroll
          for (int b = 0; b < B; ++b) { lw[b] = __expf(lw[b] - mx); se += lw[b]; }
          const float ise = 1.0f / se;
          Q4 qhr = {0.f, 0.f, 0.f, 0.f}, qhd = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int b = 0; b < B; ++b) {
            lw[b] *= ise;
            const float4 r = lds128(se3 + 32u * b), dq = lds128(se3 + 32u * b + 16);
            const float ws = (qa.x * r.x + qa.y * r.y + qa.z * r.z + qa.w * r.w) > 0.f ? lw[b] : -lw[b];
            qhr.w += ws * r.x; qhr.x += ws * r.y; qhr.y += ws * r.z; qhr.z += ws * r.w;
            qhd.w += ws * dq.x; qhd.x += ws * dq.y; qhd.y += ws * dq.z; qhd.z += ws * dq.w;
          }
          const float nn2 = qhr.w * qhr.w + qhr.x * qhr.x + qhr.y * qhr.y + qhr.z * qhr.z;
          const float inn = rsqrtf(nn2);
          const Q4 Qr = {qhr.w * inn, qhr.x * inn, qhr.y * inn, qhr.z * inn}, Qd = {qhd.w * inn, qhd.x * inn, qhd.y * inn, qhd.z * inn};
          float3 xo = qrot(Qr, x);
          {
            const Q4 tq = qmul(Qd, qconj(Qr));
            xo.x += 2.f * tq.x; xo.y += 2.f * tq.y; xo.z += 2.f * tq.z;
          }
          // ---- cotangent of the warped point
          float3 g_xo;
          float ge = g_ent, gk = g_dsk;
          if (WARPONLY) {  // the warped point's own cotangent
            g_xo = make_float3(Sl * __ldg(p.g_points + s * 3), Sl * __ldg(p.g_points + s * 3 + 1), Sl * __ldg(p.g_points + s * 3 + 2));
          } else if (w == 2) {  // cycle: |x_cyc - xyz_t|
            const float dx = xo.x - xyz_t.x, dy = xo.y - xyz_t.y, dz = xo.z - xyz_t.z;
            const float cyc = sqrtf(dx * dx + dy * dy + dz * dz);
            const float gs = cyc > 0.f ? g_cyc / cyc : 0.f;
            g_xo = make_float3(gs * dx, gs * dy, gs * dz);
            g_xyz_t.x -= g_xo.x; g_xyz_t.y -= g_xo.y; g_xyz_t.z -= g_xo.z;
          } else if (w == 1) {
            g_xo = flow_backward(xo);
            ge = 0.f; gk = 0.f;
          } else {
            g_xo = g_xyz;
            if constexpr (DENSE) g_xo = dense_backward(1, 0, x_soft[0], g_xyz);  // canonical = soft deformation of the skinned point
          }
          // ---- blend backward (oracle/skin_backward.py)
          const Q4 G = {0.f, g_xo.x, g_xo.y, g_xo.z}, Pq = {0.f, x.x, x.y, x.z};
          const Q4 u = qmul(Qr, Pq), g_u = qmul(G, Qr);
          Q4 g_Qr = qadd(qmul(qconj(G), u), qmul(g_u, qconj(Pq)));
          const Q4 gxq = qmul(qconj(Qr), g_u);
          float3 g_x = make_float3(gxq.x, gxq.y, gxq.z);
          Q4 g_Qd = qmul(G, Qr);
          g_Qd = {2.f * g_Qd.w, 2.f * g_Qd.x, 2.f * g_Qd.y, 2.f * g_Qd.z};
          {
            const Q4 t = qmul(qconj(G), Qd);
            g_Qr = {g_Qr.w + 2.f * t.w, g_Qr.x + 2.f * t.x, g_Qr.y + 2.f * t.y, g_Qr.z + 2.f * t.z};
          }
          const Q4 g_qhd = {g_Qd.w * inn, g_Qd.x * inn, g_Qd.y * inn, g_Qd.z * inn};
          const float radial = g_Qr.w * qhr.w + g_Qr.x * qhr.x + g_Qr.y * qhr.y + g_Qr.z * qhr.z + g_Qd.w * qhd.w + g_Qd.x * qhd.x + g_Qd.y * qhd.y + g_Qd.z * qhd.z;
          const float rn3 = radial * inn * inn * inn;
          const Q4 g_qhr = {g_Qr.w * inn - qhr.w * rn3, g_Qr.x * inn - qhr.x * rn3, g_Qr.y * inn - qhr.y * rn3, g_Qr.z * inn - qhr.z * rn3};
          float gd2[B];  // becomes dL/d dist2_b
          float gwsum = 0.f;
#pragma unroll
          for (int b = 0; b < B; ++b) {
            const float4 r = lds128(se3 + 32u * b), dq = lds128(se3 + 32u * b + 16);
            const float sg = (qa.x * r.x + qa.y * r.y + qa.z * r.z + qa.w * r.w) > 0.f ? 1.f : -1.f;
            const float gw = sg * (g_qhr.w * r.x + g_qhr.x * r.y + g_qhr.y * r.z + g_qhr.z * r.w + g_qhd.w * dq.x + g_qhd.x * dq.y + g_qhd.y * dq.z + g_qhd.z * dq.w);
            gd2[b] = gw;
            gwsum += lw[b] * gw;
          }
          // operands of the bone-table gradients (wgrad kernel): [x y z 1 | g_qhr | g_qhd] and the signed weights
          {
            uint8_t* xg = gt_tile + tape_row_off(TL.n_g, TL.g_xg[w], row);
            gtape_zero_row(TL.g_xg[w]);
            *reinterpret_cast<uint4*>(xg + ((row & 7u) << 4)) = make_uint4(Op::pack2(x.x, x.y), Op::pack2(x.z, 1.0f), Op::pack2_sat(g_qhr.w, g_qhr.x), Op::pack2_sat(g_qhr.y, g_qhr.z));
            *reinterpret_cast<uint4*>(xg + (((row & 7u) ^ 1u) << 4)) = make_uint4(Op::pack2_sat(g_qhd.w, g_qhd.x), Op::pack2_sat(g_qhd.y, g_qhd.z), 0u, 0u);
          }
          {
            uint32_t gz[16], wsp[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float zz[2], ww[2];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int b = 2 * j + h;
                zz[h] = 0.f; ww[h] = 0.f;
                if (b < B) {
                  const float glog = lw[b] * (gd2[b] - gwsum) + ge * (lw[b] - (b == amax ? 1.f : 0.f));
                  const float dlt = 0.1f * fmaxf(zr[b], 0.f);
                  zz[h] = zr[b] > 0.f ? 0.1f * (-glog + gk * 2.f * dlt / (float)B) : 0.f;
                  gd2[b] = -glog;
                  const float4 r = lds128(se3 + 32u * b);
                  ww[h] = (qa.x * r.x + qa.y * r.y + qa.z * r.z + qa.w * r.w) > 0.f ? lw[b] : -lw[b];
                }
              }
              gz[j] = Op::pack2_sat(zz[0], zz[1]);
              wsp[j] = Op::pack2(ww[0], ww[1]);
            }
            tmem_st16(tA, gz);
            tmem_st_wait();
            gtape_zero_row(TL.g_z[w]);
            gtape_st32(TL.g_z[w], 0, gz);
            gtape_st32(TL.g_xbw[w], 96, wsp);  // columns 96.. of the [g_xb | ws] operand
          }
          // ---- delta MLP in reverse on the tensor pipe
          gemm();
          seq_dgrad(64, TL.m_h2[w], TL.g_z2[w]);
          gemm();
          seq_dgrad(64, TL.m_h1[w], TL.g_z1[w]);
          gemm();
          // accumulator: dL/d bone coordinates through the MLP; add the dist2 path, fold back onto the point
#pragma unroll
          for (int blk = 0; blk < 3; ++blk) {  // unrolled: the bone of every column is then a compile-time index
            float v[32];
            uint32_t o[16];
            tmem_ld32(tD + 32 * blk, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int c = 32 * blk + j;
              float gv = 0.f;
              if (c < 3 * B) {
                const float4 rr = lds128(binv + 16u * c);
                const float xb = rr.x * x.x + rr.y * x.y + rr.z * x.z + rr.w;
                gv = v[j] + 2.f * xb * gd2[c / 3 < B ? c / 3 : 0];
                g_x.x += rr.x * gv; g_x.y += rr.y * gv; g_x.z += rr.z * gv;
              }
              v[j] = gv;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = Op::pack2_sat(v[2 * i], v[2 * i + 1]);
            gtape_st32(TL.g_xbw[w], 32 * blk, o);
          }
          if (w == 0) {
            g_xyz_t.x += g_x.x; g_xyz_t.y += g_x.y; g_xyz_t.z += g_x.z;
          } else {
            if constexpr (DENSE) g_x = dense_backward(0, w, xyz, g_x);  // forward warps deform the canonical point first
            g_xyz.x += g_x.x; g_xyz.y += g_x.y; g_xyz.z += g_x.z;
          }
        }
      } else {
        // rigid field: canonical point = time-t point; the flow sees it through the partner camera only
        if constexpr (NORMALS) {
          g_xyz_t = g_xyz;
        } else {
          const float3 gx = flow_backward(xyz);
          g_xyz_t = make_float3(g_xyz.x + gx.x, g_xyz.y + gx.y, g_xyz.z + gx.z);
        }
      }

      if constexpr (WARPONLY) {  // gradient w.r.t. the given point; no camera, no ray
        if (live) {
          const float inv = 1.0f / S;
          p.g_points_out[s * 3] = g_xyz.x * inv; p.g_points_out[s * 3 + 1] = g_xyz.y * inv; p.g_points_out[s * 3 + 2] = g_xyz.z * inv;
        }
        continue;
      }
      if constexpr (NORMALS) {  // xyz_t = R(qi) xyz_cam + ti: the gradient w.r.t. the camera-space point
        Q4 gq; float3 gxc;
        qrot_bwd(qi, xyz_cam, g_xyz_t, gq, gxc);
        if (live) {
          const float inv = 1.0f / p.eik.scale_a;
          p.g_points_out[s * 3] = gxc.x * inv; p.g_points_out[s * 3 + 1] = gxc.y * inv; p.g_points_out[s * 3 + 2] = gxc.z * inv;
        }
        continue;
      }
      // ================================================================ camera -> field, sample placement
      {
        red[13] += g_xyz_t.x; red[14] += g_xyz_t.y; red[15] += g_xyz_t.z;
        Q4 gq; float3 gxc;
        qrot_bwd(qi, xyz_cam, g_xyz_t, gq, gxc);
        red[9] += gq.w; red[10] += gq.x; red[11] += gq.y; red[12] += gq.z;
        const float3 gc = ld3(p.g.xyz_cam);
        const float3 gd = make_float3((gxc.x + gc.x) * depth, (gxc.y + gc.y) * depth, (gxc.z + gc.z) * depth);
        red[0] += gd.x * h0; red[1] += gd.x * h1; red[2] += gd.x * h2; red[3] += gd.y * h0; red[4] += gd.y * h1; red[5] += gd.y * h2;
        red[6] += gd.z * h0; red[7] += gd.z * h1; red[8] += gd.z * h2;
        red[29] = -ld1(p.g.depth) * depth * lds32(sc_s + 4u * SC_INV_SCALE);  // depth / exp(logscale)
      }
      // ---- tile reduction: butterfly over the warp (lane i ends with the warp's sum of value i), then atomics
      {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
          const bool hi = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < off; ++i) {
            const float send = hi ? red[i] : red[i + off];
            const float keep = hi ? red[i + off] : red[i];
            red[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
        const float val = red[0] / S;
        if (!dead_tile && val != 0.f) {
          float* gfb = p.g_fblk + (size_t)f * FL.n_floats;
          float* dst = nullptr;
          if (lane < 9) dst = gfb + FL.cam + lane;                       // Kinv
          else if (lane < 16) dst = gfb + FL.cam + 11 + (lane - 9);      // qi (4), ti (3): gradients w.r.t. the INVERSE camera
          else if (lane < 20) dst = gfb + FL.cam_partner + (lane == 16 ? 0 : (lane == 17 ? 2 : (lane == 18 ? 4 : 5)));
          else if (lane < 27) dst = gfb + FL.cam_partner + 11 + (lane - 20);
          else if (lane == 27) dst = p.g_cblk + CL.scalars + SC_IBETA;
          else if (lane == 28) dst = p.g_cblk + CL.scalars + SC_WARP_IBETA;
          else if (lane == 29) dst = p.g_cblk + CL.scalars + SC_INV_SCALE;
          if (dst) atomicAdd(dst, val);
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();
  if (warp == 9) {
    tc_fence_after_sync();
    tmem_dealloc(0, kTmemCols);
  }
}

template <class Op, int B, int WIDTH, bool DENSE, bool EIK = false, bool WARPONLY = false, bool NORMALS = false>
static cudaError_t launch_one(const BwdKernelParams& p, int n_sm, cudaStream_t stream) {
  auto kern = field_bwd_kernel<Op, B, WIDTH, DENSE, EIK, WARPONLY, NORMALS>;
  const int smem = 1024 + kSmemRing + (p.prog.cl.n_floats + kGroups * p.prog.fl.n_floats) * 4 + 256;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  int grid = (p.n_tiles + 1) / 2 < n_sm ? (p.n_tiles + 1) / 2 : n_sm;
  grid = (grid + kCluster - 1) / kCluster * kCluster;
  if (grid > n_sm) grid -= kCluster;
  if (grid < kCluster) grid = kCluster;
  if (grid > kMaxCtas) return cudaErrorInvalidValue;
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

}  // namespace bwd

cudaError_t launch_field_bwd(const BwdKernelParams& p, int n_sm, cudaStream_t stream) {
  const bool bf = p.desc.operand_dtype == 1;
  if (p.eik.mode != 0) {  // eikonal chains: the basefield only, whatever warps the field has
    if (p.desc.W == 256) return bf ? bwd::launch_one<OpBF16, 0, 256, false, true>(p, n_sm, stream) : bwd::launch_one<OpF16, 0, 256, false, true>(p, n_sm, stream);
    if (p.desc.W == 128) return bf ? bwd::launch_one<OpBF16, 0, 128, false, true>(p, n_sm, stream) : bwd::launch_one<OpF16, 0, 128, false, true>(p, n_sm, stream);
    return cudaErrorInvalidValue;
  }
  if (p.normals) {  // d sdf / d xyz_cam of every sample (b200r_field_normals)
#define B200R_NCASE(BN, WD, DN)                                                                                          \
  if (p.desc.n_bones == BN && p.desc.W == WD && (p.desc.dense != 0) == DN)                                                \
    return bf ? bwd::launch_one<OpBF16, BN, WD, DN, false, false, true>(p, n_sm, stream) : bwd::launch_one<OpF16, BN, WD, DN, false, false, true>(p, n_sm, stream);
    B200R_NCASE(0, 128, false)
    B200R_NCASE(0, 256, false)
    B200R_NCASE(18, 256, false)
    B200R_NCASE(25, 256, false)
    B200R_NCASE(18, 256, true)
    B200R_NCASE(25, 256, true)
#undef B200R_NCASE
    return cudaErrorInvalidValue;
  }
  if (p.g_points_out) {  // one forward warp of given points (b200r_warp_bwd)
#define B200R_WCASE(BN, DN)                                                                                              \
  if (p.desc.n_bones == BN && p.desc.W == 256 && (p.desc.dense != 0) == DN)                                               \
    return bf ? bwd::launch_one<OpBF16, BN, 256, DN, false, true>(p, n_sm, stream) : bwd::launch_one<OpF16, BN, 256, DN, false, true>(p, n_sm, stream);
    B200R_WCASE(18, false)
    B200R_WCASE(25, false)
    B200R_WCASE(18, true)
    B200R_WCASE(25, true)
#undef B200R_WCASE
    return cudaErrorInvalidValue;
  }
#define B200R_CASE(BN, WD, DN)                                                 \
  if (p.desc.n_bones == BN && p.desc.W == WD && (p.desc.dense != 0) == DN)     \
    return bf ? bwd::launch_one<OpBF16, BN, WD, DN>(p, n_sm, stream) : bwd::launch_one<OpF16, BN, WD, DN>(p, n_sm, stream);
  B200R_CASE(0, 128, false)
  B200R_CASE(0, 256, false)
  B200R_CASE(18, 256, false)
  B200R_CASE(25, 256, false)
  B200R_CASE(18, 256, true)
  B200R_CASE(25, 256, true)
#undef B200R_CASE
  return cudaErrorInvalidValue;
}

}  // namespace b200r
