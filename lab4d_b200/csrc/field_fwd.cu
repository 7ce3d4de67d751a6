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
//               multicast to both CTAs of the cluster) through a 3-stage ring.
//   warp 9    : tcgen05.mma issuer (one elected lane) + TMEM owner.
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

template <int V>
using IC = std::integral_constant<int, V>;

// ---------------------------------------------------------------------------------- the kernel
template <class Op, int B, int LMAX>
__global__ void __launch_bounds__(kThreads, 1) field_fwd_kernel(const __grid_constant__ FieldKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* arena = smem;
  uint8_t* ring = smem + kSmemArena;
  float* cblk = reinterpret_cast<float*>(ring + kSmemRing);
  float* fblk = cblk + p.prog.cl.n_floats;
  uint64_t* bars = reinterpret_cast<uint64_t*>(fblk + p.prog.fl.n_floats);
  uint64_t* full_bar = bars;                       // [kNumStages]
  uint64_t* empty_bar = bars + kNumStages;         // [kNumStages]
  uint64_t* a_ready = bars + 2 * kNumStages;       // compute warps -> MMA warp
  uint64_t* acc_full = bars + 2 * kNumStages + 1;  // MMA warp -> compute warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kNumStages + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kNumStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], kCluster); }
    mbar_init(a_ready, kComputeThreads);
    mbar_init(acc_full, 1);
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
        for (int g = 0; g < P.n_seq; ++g) {
          const GemmDesc& G = P.seq[g];
          const uint32_t bytes = (uint32_t)G.n_pad * 128u;
          const uint32_t part = bytes / kCluster;
          for (int c = 0; c < G.n_chunks; ++c) {
            mbar_wait(&empty_bar[stage], phase ^ 1);  // both CTAs' MMAs are done with this slot
            mbar_arrive_expect_tx(&full_bar[stage], bytes);
            const uint8_t* src = p.packed + G.w_off + (uint32_t)c * bytes + cta_rank * part;
            uint8_t* dst = ring + stage * kWStageBytes + cta_rank * part;
            if (kCluster > 1) tma_bulk_g2s_mcast(dst, src, part, &full_bar[stage], cmask);
            else tma_bulk_g2s(dst, src, part, &full_bar[stage]);
            if (++stage == kNumStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 9) {
    // =============================================================== MMA issuer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, a_phase = 0;
      const uint32_t arena_addr = smem_u32(arena), ring_addr = smem_u32(ring);
      for (int it = 0; it < iters; ++it) {
        for (int g = 0; g < P.n_seq; ++g) {
          const GemmDesc& G = P.seq[g];
          const uint32_t idesc = umma_idesc_f16(Op::kFmt, G.n_pad);
          mbar_wait(a_ready, a_phase);
          a_phase ^= 1;
          tc_fence_after_sync();
          uint32_t acc = G.accumulate;
          for (int c = 0; c < G.n_chunks; ++c) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after_sync();
            const uint64_t adesc = umma_desc_k_sw128(arena_addr + G.a_chunk[c] * kAChunkBytes);
            const uint64_t bdesc = umma_desc_k_sw128(ring_addr + stage * kWStageBytes);
            for (int k = 0; k < G.ksteps[c]; ++k) {
              umma_f16_ss(tmem_base + G.tmem_col, umma_desc_advance_k(adesc, k), umma_desc_advance_k(bdesc, k), idesc, acc);
              acc = 1;
            }
            // frees the ring slot (in both CTAs) once these MMAs have read it
            if (kCluster > 1) umma_commit_mcast(&empty_bar[stage], cmask);
            else umma_commit(&empty_bar[stage]);
            if (++stage == kNumStages) { stage = 0; phase ^= 1; }
          }
          umma_commit(acc_full);
        }
      }
    }
  } else {
    // =============================================================== compute / epilogue warps
    const int q = warp & 3, hsel = warp >> 2;
    const uint32_t row = (uint32_t)(q * 32 + lane);  // tile row == TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint8_t* const extra = arena + CH_EXTRA * kAChunkBytes;
    uint32_t acc_phase = 0;
    const int W = p.desc.W;
    const ConstLayout& CL = P.cl;
    const FrameLayout& FL = P.fl;
    const float* const sc = cblk + CL.scalars;

    // stage the constant block once (made visible by the first named barrier of the tile loop)
    {
      const float4* src = reinterpret_cast<const float4*>(p.workspace);
      float4* dst = reinterpret_cast<float4*>(cblk);
      for (int i = threadIdx.x; i < CL.n_floats / 4; i += kComputeThreads) dst[i] = __ldg(src + i);
    }

    // Exchange between the two threads of a row: 12 floats per thread inside the unused part of the
    // CH_EXTRA rows (the MMA only reads their first 32 B).  Every exchange round uses its own floats
    // and rounds that reuse an address are separated by a run_gemm() (a barrier of all compute threads).
    float* const my_x0 = reinterpret_cast<float*>(extra + sw128_off(row, 2 + 3 * hsel));
    float* const my_x1 = reinterpret_cast<float*>(extra + sw128_off(row, 3 + 3 * hsel));
    float* const my_x2 = reinterpret_cast<float*>(extra + sw128_off(row, 4 + 3 * hsel));
    const float* const pr_x0 = reinterpret_cast<const float*>(extra + sw128_off(row, 2 + 3 * (hsel ^ 1)));
    const float* const pr_x1 = reinterpret_cast<const float*>(extra + sw128_off(row, 3 + 3 * (hsel ^ 1)));
    const float* const pr_x2 = reinterpret_cast<const float*>(extra + sw128_off(row, 4 + 3 * (hsel ^ 1)));
    auto pair_sync = [&]() { named_bar_sync(1 + q, 64); };

    // hand the operands to the MMA warp, then wait for the layer's accumulator
    auto run_gemm = [&]() {
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(a_ready);
      mbar_wait(acc_full, acc_phase);
      acc_phase ^= 1;
      tc_fence_after_sync();
    };
    auto bias_of = [&](int s_idx) -> const float* {
      const GemmDesc& G = P.seq[s_idx];
      return (G.bias_frame ? fblk : cblk) + G.bias_off;
    };
    // relu(acc + bias) -> 16-bit operand rows; this thread covers its half of the columns
    auto epi_relu_store = [&](int s_idx, int dst_chunk) {
      const GemmDesc& G = P.seq[s_idx];
      const float* bias = bias_of(s_idx);
      const int ncols = G.n_pad >> 1, cb = hsel * ncols;
      for (int c0 = cb; c0 < cb + ncols; c0 += 32) {
        float v[32];
        tmem_ld32(t_lane + G.tmem_col + c0, v);
        uint8_t* chunk = arena + (dst_chunk + (c0 >> 6)) * kAChunkBytes;
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const float4 b0 = *reinterpret_cast<const float4*>(bias + c0 + g8 * 8);
          const float4 b1 = *reinterpret_cast<const float4*>(bias + c0 + g8 * 8 + 4);
          uint4 o;
          o.x = Op::pack2_relu(v[g8 * 8 + 0] + b0.x, v[g8 * 8 + 1] + b0.y);
          o.y = Op::pack2_relu(v[g8 * 8 + 2] + b0.z, v[g8 * 8 + 3] + b0.w);
          o.z = Op::pack2_relu(v[g8 * 8 + 4] + b1.x, v[g8 * 8 + 5] + b1.y);
          o.w = Op::pack2_relu(v[g8 * 8 + 6] + b1.z, v[g8 * 8 + 7] + b1.w);
          *reinterpret_cast<uint4*>(chunk + sw128_off(row, ((c0 & 63) >> 3) + g8)) = o;
        }
      }
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

      // ------------------------------------------------ skinning warp (SkinningWarp.forward)
      // bone coordinates -> delta MLP on the tensor pipe -> softmax -> dual-quaternion blend.
      // Half 0 owns bones [0,BS), half 1 bones [BS,B); operand groups are split at a 16-B boundary.
      constexpr int BS = B == 25 ? 13 : (B == 18 ? 8 : 0);
      constexpr int XTRA = (8 - (3 * BS) % 8) % 8;  // values of bone BS that complete half 0's last group
      constexpr int I0 = 3 * BS + XTRA;             // first operand column written by half 1
      constexpr int BH = BS > B - BS ? BS : B - BS;
      auto skin_warp = [&](auto half_tag, const float3& x, const float* binv, const float* se3, int s_first,
                           float& entropy, float& delta_skin) -> float3 {
        constexpr int HALF = decltype(half_tag)::value;
        constexpr int b_lo = HALF == 0 ? 0 : BS;
        constexpr int b_hi = HALF == 0 ? BS : B;
        constexpr int NB = b_hi - b_lo;
        constexpr int NV = HALF == 0 ? 3 * BS + XTRA : 3 * (B - BS);
        float dist2[BH > 0 ? BH : 1];
        {
          float v[NV > 0 ? NV : 1];
          const float* ig = cblk + CL.inv_gauss;
#pragma unroll
          for (int j = 0; j < (NV + 2) / 3; ++j) {
            const int b = b_lo + j;
            const Q4 qb = ldq(binv + b * 8);
            const float4 t4 = *reinterpret_cast<const float4*>(binv + b * 8 + 4);
            const float4 g4 = *reinterpret_cast<const float4*>(ig + b * 4);
            float3 xb = qrot(qb, x);
            xb.x = (xb.x + t4.x) * g4.x; xb.y = (xb.y + t4.y) * g4.y; xb.z = (xb.z + t4.z) * g4.z;
            if (j < NB) dist2[j] = xb.x * xb.x + xb.y * xb.y + xb.z * xb.z;
            if (3 * j < NV) v[3 * j] = xb.x;
            if (3 * j + 1 < NV) v[3 * j + 1] = xb.y;
            if (3 * j + 2 < NV) v[3 * j + 2] = xb.z;
          }
          if (HALF == 0) {
#pragma unroll
            for (int g = 0; g < NV / 8; ++g) store_group<Op>(arena + CH_H0 * kAChunkBytes, row, g, v + 8 * g);
          } else {
            constexpr int END = (3 * B + 15) / 16 * 16;  // zero-padded to whole UMMA_K steps
#pragma unroll
            for (int idx = I0; idx < END; idx += 8) {
              float w8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) w8[j] = (idx + j < 3 * B) ? v[idx + j - 3 * BS] : 0.f;
              store_group<Op>(arena + (idx < 64 ? CH_H0 : CH_H1) * kAChunkBytes, row, (idx & 63) >> 3, w8);
            }
          }
        }
        // delta_field.linear_1 / linear_2 (ReLU) and linear_final
        run_gemm();
        epi_relu_store(s_first, CH_H2);
        run_gemm();
        epi_relu_store(s_first + 1, CH_H2);
        run_gemm();
        float dl[32];
        tmem_ld32(t_lane + P.seq[s_first + 2].tmem_col, dl);
        const float* b3 = bias_of(s_first + 2);
        float mx = -INFINITY, dsum = 0.f;
        int amax = 0;
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const float dv = 0.1f * fmaxf(dl[b_lo + j] + b3[b_lo + j], 0.f);
          dsum += dv * dv;
          const float lg = -(dist2[j] + dv);
          dist2[j] = lg;
          if (lg > mx) { mx = lg; amax = b_lo + j; }
        }
        // round 1: global max / anchor bone (first maximum wins, like argmax)
        my_x0[2] = mx;
        my_x0[3] = __int_as_float(amax);
        pair_sync();
        {
          const float omx = pr_x0[2];
          const int oam = __float_as_int(pr_x0[3]);
          const bool take = HALF == 0 ? (omx > mx) : (omx >= mx);
          if (take) { mx = omx; amax = oam; }
        }
        const Q4 qa = ldq(se3 + amax * 8);
        float se = 0.f;
        Q4 qr = {0, 0, 0, 0}, qd = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const int b = b_lo + j;
          const float e = expf(dist2[j] - mx);
          se += e;
          const Q4 r = ldq(se3 + b * 8), dq = ldq(se3 + b * 8 + 4);
          const float dot = qa.w * r.w + qa.x * r.x + qa.y * r.y + qa.z * r.z;
          const float wgt = dot > 0.f ? e : -e;  // the softmax denominator cancels in the normalisation below
          qr.w += wgt * r.w; qr.x += wgt * r.x; qr.y += wgt * r.y; qr.z += wgt * r.z;
          qd.w += wgt * dq.w; qd.x += wgt * dq.x; qd.y += wgt * dq.y; qd.z += wgt * dq.z;
        }
        // round 2: partial sums
        my_x0[0] = se;
        my_x0[1] = dsum;
        *reinterpret_cast<float4*>(my_x1) = make_float4(qr.w, qr.x, qr.y, qr.z);
        *reinterpret_cast<float4*>(my_x2) = make_float4(qd.w, qd.x, qd.y, qd.z);
        pair_sync();
        {
          const float ose = pr_x0[0], ods = pr_x0[1];
          const float4 o1 = *reinterpret_cast<const float4*>(pr_x1), o2 = *reinterpret_cast<const float4*>(pr_x2);
          // add in bone order (half 0 first) so both threads of the row get bit-identical results
          if (HALF == 0) {
            se = se + ose; dsum = dsum + ods;
            qr = {qr.w + o1.x, qr.x + o1.y, qr.y + o1.z, qr.z + o1.w};
            qd = {qd.w + o2.x, qd.x + o2.y, qd.y + o2.z, qd.z + o2.w};
          } else {
            se = ose + se; dsum = ods + dsum;
            qr = {o1.x + qr.w, o1.y + qr.x, o1.z + qr.y, o1.w + qr.z};
            qd = {o2.x + qd.w, o2.y + qd.x, o2.z + qd.y, o2.w + qd.z};
          }
        }
        entropy = logf(se);  // logsumexp - max  (cross_entropy_skin_loss)
        delta_skin = dsum / (float)(B > 0 ? B : 1);
        const float inv = 1.0f / sqrtf(qr.w * qr.w + qr.x * qr.x + qr.y * qr.y + qr.z * qr.z);
        qr = {qr.w * inv, qr.x * inv, qr.y * inv, qr.z * inv};
        qd = {qd.w * inv, qd.x * inv, qd.y * inv, qd.z * inv};
        const Q4 tq = qmul(qd, qconj(qr));
        float3 o = qrot(qr, x);
        o.x += 2.f * tq.x; o.y += 2.f * tq.y; o.z += 2.f * tq.z;
        return o;
      };

      float3 xyz = xyz_t;
      float ent_b = 0.f, dsk_b = 0.f, ent_out = 0.f, dsk_out = 0.f;
      if constexpr (B > 0) {
        xyz = hsel == 0 ? skin_warp(IC<0>{}, xyz_t, fblk + FL.binv_t, fblk + FL.se3_bwd, P.seq_delta_bwd, ent_b, dsk_b)
                        : skin_warp(IC<1>{}, xyz_t, fblk + FL.binv_t, fblk + FL.se3_bwd, P.seq_delta_bwd, ent_b, dsk_b);
      }

      // ------------------------------------------------ positional embedding (PosEmbedding.forward)
      // half 0: x and frequencies 0..5 (+ the first value of frequency 6) = 40 columns;
      // half 1: frequencies 6..LMAX-1 -> columns 40..62, zero column 63, CH_EXTRA columns 0..15
      if (hsel == 0) {
        float v[40];
        v[0] = xyz.x; v[1] = xyz.y; v[2] = xyz.z;
        float fr = 1.0f;
#pragma unroll
        for (int kf = 0; kf < 6; ++kf) {
          sincosf(fr * xyz.x, &v[3 + 6 * kf + 0], &v[3 + 6 * kf + 3]);
          sincosf(fr * xyz.y, &v[3 + 6 * kf + 1], &v[3 + 6 * kf + 4]);
          sincosf(fr * xyz.z, &v[3 + 6 * kf + 2], &v[3 + 6 * kf + 5]);
          fr *= 2.0f;
        }
        v[39] = sinf(64.0f * xyz.x);
#pragma unroll
        for (int g = 0; g < 5; ++g) store_group<Op>(arena + CH_PE * kAChunkBytes, row, g, v + 8 * g);
      } else {
        constexpr int NF = LMAX - 6;
        float v[6 * NF];  // embedding columns 39 .. 39+6*NF-1
        float fr = 64.0f;
#pragma unroll
        for (int kf = 0; kf < NF; ++kf) {
          sincosf(fr * xyz.x, &v[6 * kf + 0], &v[6 * kf + 3]);
          sincosf(fr * xyz.y, &v[6 * kf + 1], &v[6 * kf + 4]);
          sincosf(fr * xyz.z, &v[6 * kf + 2], &v[6 * kf + 5]);
          fr *= 2.0f;
        }
        // columns 40..63 of CH_PE (column 63 is the zero pad)
#pragma unroll
        for (int g = 5; g < 8; ++g) {
          float w8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) w8[j] = (8 * g + j < 63 && 8 * g + j - 39 < 6 * NF) ? v[(8 * g + j - 39) < 6 * NF ? (8 * g + j - 39) : 0] : 0.f;
          store_group<Op>(arena + CH_PE * kAChunkBytes, row, g, w8);
        }
        if (LMAX > 10) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            float w8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w8[j] = (24 + 8 * g + j < 6 * NF) ? v[(24 + 8 * g + j) < 6 * NF ? (24 + 8 * g + j) : 0] : 0.f;
            store_group<Op>(extra, row, g, w8);
          }
        }
      }

      // ------------------------------------------------ visibility MLP (VisField.forward)
      int seq = P.seq_vis;
      run_gemm();
      epi_relu_store(seq, CH_H0);
      run_gemm();
      float vis_out;
      {
        const float* b2 = bias_of(seq + 1);
        const float* vw = cblk + CL.vis_w;
        float v[32];
        tmem_ld32(t_lane + P.seq[seq + 1].tmem_col + 32 * hsel, v);
        float accv = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) accv += fmaxf(v[j] + b2[32 * hsel + j], 0.f) * vw[32 * hsel + j];
        my_x0[0] = accv;
        pair_sync();
        const float other = pr_x0[0];
        vis_out = (hsel == 0 ? accv + other : other + accv) + sc[SC_VIS_B];
      }

      // ------------------------------------------------ density branch (NeRF.forward, basefield + sdf)
      seq = P.seq_base;
      for (int i = 0; i < p.desc.D; ++i) {
        run_gemm();
        epi_relu_store(seq + i, CH_H0);
      }
      run_gemm();
      float sdf;
      {
        const int sf = seq + p.desc.D;
        const GemmDesc& G = P.seq[sf];
        const float* bb = bias_of(sf);
        const float* sw = cblk + CL.sdf_w;
        const int ncols = W >> 1, cb = hsel * ncols;
        float accs = 0.f;
        for (int c0 = cb; c0 < cb + ncols; c0 += 32) {
          float v[32];
          tmem_ld32(t_lane + G.tmem_col + c0, v);
          uint8_t* chunk = arena + (CH_H0 + (c0 >> 6)) * kAChunkBytes;
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              y[j] = fmaxf(v[g8 * 8 + j] + bb[c0 + g8 * 8 + j], 0.f);
              accs += y[j] * sw[c0 + g8 * 8 + j];
            }
            store_group<Op>(chunk, row, ((c0 & 63) >> 3) + g8, y);
          }
        }
        my_x1[0] = accs;
        pair_sync();
        const float other = pr_x1[0];
        sdf = (hsel == 0 ? accs + other : other + accs) + sc[SC_SDF_B];
      }
      const float ibeta = sc[SC_IBETA];
      const float sgn = sdf > 0.f ? 1.f : (sdf < 0.f ? -1.f : 0.f);
      const float density = (0.5f + 0.5f * sgn * expm1f(-fabsf(sdf) * ibeta)) * ibeta;

      // ------------------------------------------------ colour branch: rgb.0 is linear in (base + colour)
      run_gemm();  // base features x rgb.0 -> TMEM[kTmemRgb..)
      seq = P.seq_color;
      run_gemm();
      epi_relu_store(seq, CH_H0);
      run_gemm();
      epi_relu_store(seq + 1, CH_H0);
      run_gemm();
      epi_relu_store(seq + 2, CH_H0);
      seq = P.seq_rgb2;
      run_gemm();  // + colour features x rgb.0
      float rgb[3];
      {
        const float* b0 = bias_of(seq);
        const int H = W / 2, ncols = H >> 1, cb = hsel * ncols;
        const float* w2 = cblk + CL.rgb2_w;
        const float* wd = cblk + CL.dir_w;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int c0 = cb; c0 < cb + ncols; c0 += 32) {
          float v[32];
          tmem_ld32(t_lane + kTmemRgb + c0, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float pre = v[j] + b0[c0 + j];
            if (p.desc.L_dir == 0) pre += wd[(c0 + j) * 3] * dir_f.x + wd[(c0 + j) * 3 + 1] * dir_f.y + wd[(c0 + j) * 3 + 2] * dir_f.z;
            const float hh = fmaxf(pre, 0.f);
            a0 += hh * w2[c0 + j];
            a1 += hh * w2[H + c0 + j];
            a2 += hh * w2[2 * H + c0 + j];
          }
        }
        *reinterpret_cast<float4*>(my_x2) = make_float4(a0, a1, a2, 0.f);
        pair_sync();
        const float4 o = *reinterpret_cast<const float4*>(pr_x2);
        if (hsel == 0) { a0 = a0 + o.x; a1 = a1 + o.y; a2 = a2 + o.z; }
        else { a0 = o.x + a0; a1 = o.y + a1; a2 = o.z + a2; }
        a0 += sc[SC_RGB2_B0]; a1 += sc[SC_RGB2_B1]; a2 += sc[SC_RGB2_B2];
        rgb[0] = 1.f / (1.f + expf(-a0)); rgb[1] = 1.f / (1.f + expf(-a1)); rgb[2] = 1.f / (1.f + expf(-a2));
      }

      // ------------------------------------------------ feature field (FeatureNeRF.compute_feat)
      float feat[16];
      if (p.desc.has_feature) {
        seq = P.seq_feat;
        for (int i = 0; i < 5; ++i) {
          run_gemm();
          epi_relu_store(seq + i, CH_H0);
        }
        run_gemm();
        if (hsel == 0) {  // warp-uniform: 16 outputs, one thread per row
          float v16[16];
          tmem_ld16(t_lane + P.seq[seq + 5].tmem_col, v16);
          const float* bf = bias_of(seq + 5);
          float nn = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) { feat[j] = v16[j] + bf[j]; nn += feat[j] * feat[j]; }
          const float inv = 1.0f / sqrtf(nn);
#pragma unroll
          for (int j = 0; j < 16; ++j) feat[j] *= inv;
        }
      }

      // ------------------------------------------------ flow + cycle warps (compute_flow, cycle_loss)
      float flow[3] = {0.f, 0.f, 0.f};
      float cyc = 0.f;
      float3 x_next = xyz;
      if constexpr (B > 0) {
        float e1, d1, e2, d2;
        x_next = hsel == 0 ? skin_warp(IC<0>{}, xyz, fblk + FL.binv_rest_partner, fblk + FL.se3_fwd_partner, P.seq_delta_flow, e1, d1)
                           : skin_warp(IC<1>{}, xyz, fblk + FL.binv_rest_partner, fblk + FL.se3_fwd_partner, P.seq_delta_flow, e1, d1);
        const float3 xc = hsel == 0 ? skin_warp(IC<0>{}, xyz, fblk + FL.binv_rest, fblk + FL.se3_fwd, P.seq_delta_cyc, e2, d2)
                                    : skin_warp(IC<1>{}, xyz, fblk + FL.binv_rest, fblk + FL.se3_fwd, P.seq_delta_cyc, e2, d2);
        const float dx = xc.x - xyz_t.x, dy = xc.y - xyz_t.y, dz = xc.z - xyz_t.z;
        cyc = sqrtf(dx * dx + dy * dy + dz * dz);
        ent_out = 0.5f * (e2 + ent_b);
        dsk_out = 0.5f * (d2 + dsk_b);
      }
      if (hsel == 1 || !live) continue;  // half 0 writes the sample's outputs

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
      float gdens = 0.f;
      if constexpr (B > 0) {
        float best = -INFINITY;
        const float* ctr = cblk + CL.center;
#pragma unroll
        for (int b = 0; b < B; ++b) {
          const float4 c = *reinterpret_cast<const float4*>(ctr + b * 4);
          const float dx = xyz.x - c.x, dy = xyz.y - c.y, dz = xyz.z - c.z;
          const float d2 = (dx * dx + dy * dy + dz * dz) / (0.01f * 0.01f);
          best = fmaxf(best, expf(-0.5f * d2));
        }
        gdens = best * sc[SC_WARP_IBETA];
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
        st1(p.out.depth, depth * sc[SC_INV_SCALE]);
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

template <class Op, int B, int LMAX>
static cudaError_t launch_one(const FieldKernelParams& p, int n_sm, cudaStream_t stream) {
  auto kern = field_fwd_kernel<Op, B, LMAX>;
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
#define B200R_CASE(BN, LM)                                                        \
  if (p.desc.n_bones == BN && p.Lmax == LM)                                       \
    return bf ? launch_one<OpBF16, BN, LM>(p, n_sm, stream) : launch_one<OpF16, BN, LM>(p, n_sm, stream);
  B200R_CASE(0, 10)
  B200R_CASE(0, 12)
  B200R_CASE(18, 12)
  B200R_CASE(25, 12)
#undef B200R_CASE
  return cudaErrorInvalidValue;
}

}  // namespace b200r
