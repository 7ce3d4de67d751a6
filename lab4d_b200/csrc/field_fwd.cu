// Fused per-ray-sample forward of one Lab4D field (training-mode query_field) for sm_100a.
//
// One persistent CTA per SM walks 128-sample tiles.  Warp roles:
//   warps 0-3 (128 threads): one thread per sample = one TMEM lane.  They place the sample on its ray,
//       move it camera -> field space, run dual-quaternion blend skinning, write the 16-bit operand
//       rows (bone coordinates, positional embedding, hidden activations) into swizzled shared
//       memory, and run every layer's epilogue straight out of TMEM.
//   warp 4: TMA producer - streams pre-packed weight chunks (cp.async.bulk) through a 3-stage ring.
//   warp 5: tcgen05.mma issuer (one elected lane) + TMEM owner.
// Hidden activations never leave the SM; HBM sees O(100 B) per sample of outputs.
//
// Restates (not ports) lab4d/nnutils/{nerf,deformable,feature,warping,skinning,embedding,visibility}.py
// and lab4d/utils/{render_utils,geom_utils,quat_transform}.py - see include/b200r.h for file:line.
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include "kernels.h"
#include "ptx.cuh"

namespace b200r {

constexpr int kNumStages = 3;
constexpr int kComputeThreads = 128;
constexpr int kThreads = 192;
constexpr int kSmemArena = kArenaChunks * kAChunkBytes;          //  96 KB
constexpr int kSmemRing = kNumStages * kWStageBytes;             //  96 KB
constexpr int kSmemBytes = 1024 + kSmemArena + kSmemRing + 256;  // + alignment slack + barriers

struct FieldKernelParams {
  Program prog;
  b200r_field_desc desc;
  b200r_field_args a;
  const uint8_t* packed;
  const float* bias_seq[kMaxSeq];  // bias rows of each seq entry
  int32_t bias_stride_seq[kMaxSeq];
  int32_t S;          // M*N*D
  int32_t n_tiles;
  int32_t Lmax;       // frequencies of the shared embedding chunk(s)
};

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
  float4 v = __ldg(reinterpret_cast<const float4*>(p));
  return {v.x, v.y, v.z, v.w};
}

// Row writer: pushes fp32 values, converts to the 16-bit operand type and stores 16-byte groups
// into the swizzled [128 x 64] chunk(s).  All indices resolve at compile time once unrolled.
template <class Op>
struct RowWriter {
  uint8_t* chunk;  // generic pointer to the chunk base
  uint32_t row;
  uint32_t w[4];
  int cnt, grp;
  float hold;
  __device__ __forceinline__ RowWriter(uint8_t* c, uint32_t r) : chunk(c), row(r), cnt(0), grp(0), hold(0.f) {
    w[0] = w[1] = w[2] = w[3] = 0;
  }
  __device__ __forceinline__ void flush() {
    *reinterpret_cast<uint4*>(chunk + sw128_off(row, grp)) = make_uint4(w[0], w[1], w[2], w[3]);
    w[0] = w[1] = w[2] = w[3] = 0;
    cnt = 0;
    ++grp;
  }
  __device__ __forceinline__ void push(float v) {
    if (cnt & 1) w[cnt >> 1] = Op::pack2(hold, v);
    else hold = v;
    if (++cnt == 8) flush();
  }
  // zero-fill up to a multiple of `halves` columns (16 = one UMMA_K step)
  __device__ __forceinline__ void pad_to(int halves) {
    if (cnt & 1) { w[cnt >> 1] = Op::pack2(hold, 0.f); ++cnt; }
    if (cnt == 8) flush();
    while (((grp * 8 + cnt) % halves) != 0 || cnt != 0) {
      if (cnt == 0 && ((grp * 8) % halves) == 0) break;
      cnt = 8;  // remaining words are already zero
      flush();
    }
  }
};

// ---------------------------------------------------------------------------------- the kernel
template <class Op, int B, int LMAX>
__global__ void __launch_bounds__(kThreads, 1) field_fwd_kernel(const __grid_constant__ FieldKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* arena = smem;
  uint8_t* ring = smem + kSmemArena;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + kSmemRing);
  uint64_t* full_bar = bars;                   // [kNumStages]
  uint64_t* empty_bar = bars + kNumStages;     // [kNumStages]
  uint64_t* a_ready = bars + 2 * kNumStages;   // compute warps -> MMA warp
  uint64_t* acc_full = bars + 2 * kNumStages + 1;  // MMA warp -> compute warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kNumStages + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kNumStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(a_ready, kComputeThreads);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const Program& P = p.prog;

  if (warp == 4) {
    // =============================================================== TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int g = 0; g < P.n_seq; ++g) {
          const GemmDesc& G = P.seq[g];
          const uint32_t bytes = (uint32_t)G.n_pad * 128u;
          for (int c = 0; c < G.n_chunks; ++c) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], bytes);
            tma_bulk_g2s(ring + stage * kWStageBytes, p.packed + G.w_off + (uint32_t)c * bytes, bytes, &full_bar[stage]);
            if (++stage == kNumStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 5) {
    // =============================================================== MMA issuer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, a_phase = 0;
      const uint32_t arena_addr = smem_u32(arena), ring_addr = smem_u32(ring);
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
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
            umma_commit(&empty_bar[stage]);  // frees the ring slot once these MMAs have read it
            if (++stage == kNumStages) { stage = 0; phase ^= 1; }
          }
          umma_commit(acc_full);
        }
      }
    }
  } else {
    // =============================================================== compute / epilogue warps
    const uint32_t row = threadIdx.x;  // tile row == TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint32_t acc_phase = 0;
    int seq = 0;  // position in P.seq, advanced in lock-step with the other roles
    const int ND = p.a.N * p.a.D;
    const int W = p.desc.W;

    // hand the operands to the MMA warp, then wait for the layer's accumulator
    auto run_gemm = [&]() {
      fence_proxy_async_smem();
      tc_fence_before_sync();
      mbar_arrive(a_ready);
      mbar_wait(acc_full, acc_phase);
      acc_phase ^= 1;
      tc_fence_after_sync();
    };
    auto bias_ptr = [&](int s_idx, int f) { return p.bias_seq[s_idx] + (size_t)f * p.bias_stride_seq[s_idx]; };
    // relu(acc + bias) -> 16-bit operand rows of arena chunks dst_chunk, dst_chunk+1, ...
    auto epi_relu_store = [&](int s_idx, const float* bias, int dst_chunk) {
      const GemmDesc& G = P.seq[s_idx];
      for (int c0 = 0; c0 < G.n_pad; c0 += 32) {
        float v[32];
        tmem_ld32(t_lane + G.tmem_col + c0, v);
        uint8_t* chunk = arena + (dst_chunk + (c0 >> 6)) * kAChunkBytes;
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c0 + g8 * 8));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c0 + g8 * 8 + 4));
          uint4 o;
          o.x = Op::pack2(fmaxf(v[g8 * 8 + 0] + b0.x, 0.f), fmaxf(v[g8 * 8 + 1] + b0.y, 0.f));
          o.y = Op::pack2(fmaxf(v[g8 * 8 + 2] + b0.z, 0.f), fmaxf(v[g8 * 8 + 3] + b0.w, 0.f));
          o.z = Op::pack2(fmaxf(v[g8 * 8 + 4] + b1.x, 0.f), fmaxf(v[g8 * 8 + 5] + b1.y, 0.f));
          o.w = Op::pack2(fmaxf(v[g8 * 8 + 6] + b1.z, 0.f), fmaxf(v[g8 * 8 + 7] + b1.w, 0.f));
          *reinterpret_cast<uint4*>(chunk + sw128_off(row, ((c0 & 63) >> 3) + g8)) = o;
        }
      }
    };

    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      seq = 0;
      int s_raw = tile * kTileRows + (int)row;
      const bool live = s_raw < p.S;
      const int s = live ? s_raw : p.S - 1;
      const int f = s / ND;
      const int r_in = s - f * ND;
      const int n = r_in / p.a.D;
      const int k = r_in - n * p.a.D;
      const int fn = (p.a.M >= 2) ? (f ^ 1) : f;  // flip_pair partner frame

      // ------------------------------------------------ sample placement (sample_cam_rays)
      const float* hx = p.a.hxy + ((size_t)f * p.a.N + n) * 3;
      const float h0 = __ldg(hx), h1 = __ldg(hx + 1), h2 = __ldg(hx + 2);
      const float* Ki = p.a.Kinv + (size_t)f * 9;
      float3 d = make_float3(h0 * __ldg(Ki + 0) + h1 * __ldg(Ki + 1) + h2 * __ldg(Ki + 2),
                             h0 * __ldg(Ki + 3) + h1 * __ldg(Ki + 4) + h2 * __ldg(Ki + 5),
                             h0 * __ldg(Ki + 6) + h1 * __ldg(Ki + 7) + h2 * __ldg(Ki + 8));
      const float dn = sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
      const float nearv = __ldg(p.a.near_far + 2 * f), farv = __ldg(p.a.near_far + 2 * f + 1);
      const int Dn = p.a.D;
      const float step = 1.0f / (float)(Dn - 1);
      auto lin = [&](int i) { return i < Dn / 2 ? step * (float)i : 1.0f - step * (float)(Dn - 1 - i); };
      auto depth_at = [&](int i) { float z = lin(i); return nearv * (1.0f - z) + farv * z; };
      const float depth = depth_at(k);
      const float delta = (k + 1 < Dn ? depth_at(k + 1) - depth : depth - depth_at(k - 1)) * dn;
      const float3 xyz_cam = make_float3(d.x * depth, d.y * depth, d.z * depth);
      const float3 dir_cam = make_float3(d.x / dn, d.y / dn, d.z / dn);

      // ------------------------------------------------ camera -> field (cam_to_field)
      const Q4 qc = ldq(p.a.field2cam + (size_t)f * 8);
      const float4 tc4 = __ldg(reinterpret_cast<const float4*>(p.a.field2cam + (size_t)f * 8 + 4));
      const Q4 qi = qconj(qc);
      const float3 ti = qrot(qi, make_float3(-tc4.x, -tc4.y, -tc4.z));
      float3 xyz_t = qrot(qi, xyz_cam);
      xyz_t.x += ti.x; xyz_t.y += ti.y; xyz_t.z += ti.z;
      const float3 dir_f = qrot(qi, dir_cam);

      // ------------------------------------------------ skinning warp (SkinningWarp.forward)
      // bone coordinates -> delta MLP on the tensor pipe -> softmax -> dual-quaternion blend
      float ent_out = 0.f, dsk_out = 0.f;
      auto skin_warp = [&](const float3& x, const float* binv, const float* se3, int s_first, const float* bias1,
                           float& entropy, float& delta_skin) -> float3 {
        float dist2[B > 0 ? B : 1];
        {
          RowWriter<Op> wr(arena + CH_H0 * kAChunkBytes, row);
#pragma unroll
          for (int b = 0; b < B; ++b) {
            const Q4 q = ldq(binv + b * 8);
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(binv + b * 8 + 4));
            const float4 ig = __ldg(reinterpret_cast<const float4*>(p.a.inv_gauss + b * 4));
            float3 xb = qrot(q, x);
            xb.x = (xb.x + t4.x) * ig.x; xb.y = (xb.y + t4.y) * ig.y; xb.z = (xb.z + t4.z) * ig.z;
            dist2[b] = xb.x * xb.x + xb.y * xb.y + xb.z * xb.z;
            if (3 * b == 63) { wr.push(xb.x); wr = RowWriter<Op>(arena + CH_H1 * kAChunkBytes, row); wr.push(xb.y); wr.push(xb.z); }
            else { wr.push(xb.x); wr.push(xb.y); wr.push(xb.z); }
          }
          wr.pad_to(16);
        }
        // delta_field.linear_1 / linear_2 (ReLU) and linear_final
        run_gemm();
        epi_relu_store(s_first, bias1, CH_H2);
        run_gemm();
        epi_relu_store(s_first + 1, bias_ptr(s_first + 1, f), CH_H2);
        run_gemm();
        float dl[32];
        tmem_ld32(t_lane + P.seq[s_first + 2].tmem_col, dl);
        const float* b3 = bias_ptr(s_first + 2, f);
        float mx = -INFINITY, dsum = 0.f;
        int amax = 0;
#pragma unroll
        for (int b = 0; b < B; ++b) {
          const float dv = 0.1f * fmaxf(dl[b] + __ldg(b3 + b), 0.f);
          dsum += dv * dv;
          const float lg = -(dist2[b] + dv);
          dist2[b] = lg;
          if (lg > mx) { mx = lg; amax = b; }
        }
        float se = 0.f;
#pragma unroll
        for (int b = 0; b < B; ++b) { dist2[b] = expf(dist2[b] - mx); se += dist2[b]; }
        entropy = logf(se);  // logsumexp - max
        delta_skin = dsum / (float)(B > 0 ? B : 1);
        const float inv_se = 1.0f / se;
        const Q4 qa = ldq(se3 + amax * 8);
        Q4 qr = {0, 0, 0, 0}, qd = {0, 0, 0, 0};
#pragma unroll
        for (int b = 0; b < B; ++b) {
          const Q4 r = ldq(se3 + b * 8), dq = ldq(se3 + b * 8 + 4);
          const float dot = qa.w * r.w + qa.x * r.x + qa.y * r.y + qa.z * r.z;
          const float wgt = (dot > 0.f ? 1.f : -1.f) * dist2[b] * inv_se;
          qr.w += wgt * r.w; qr.x += wgt * r.x; qr.y += wgt * r.y; qr.z += wgt * r.z;
          qd.w += wgt * dq.w; qd.x += wgt * dq.x; qd.y += wgt * dq.y; qd.z += wgt * dq.z;
        }
        const float inv = 1.0f / sqrtf(qr.w * qr.w + qr.x * qr.x + qr.y * qr.y + qr.z * qr.z);
        qr = {qr.w * inv, qr.x * inv, qr.y * inv, qr.z * inv};
        qd = {qd.w * inv, qd.x * inv, qd.y * inv, qd.z * inv};
        const Q4 tq = qmul(qd, qconj(qr));
        float3 o = qrot(qr, x);
        o.x += 2.f * tq.x; o.y += 2.f * tq.y; o.z += 2.f * tq.z;
        return o;
      };

      float3 xyz = xyz_t;
      float ent_b = 0.f, dsk_b = 0.f;
      if (B > 0) {
        seq = P.seq_delta_bwd;
        xyz = skin_warp(xyz_t, p.a.bone_inv_t + (size_t)f * B * 8, p.a.se3_bwd + (size_t)f * B * 8, seq,
                        bias_ptr(seq, f), ent_b, dsk_b);
      }

      // ------------------------------------------------ positional embedding (PosEmbedding.forward)
      {
        RowWriter<Op> wr(arena + CH_PE * kAChunkBytes, row);
        wr.push(xyz.x); wr.push(xyz.y); wr.push(xyz.z);
        float fr = 1.0f;
#pragma unroll
        for (int kf = 0; kf < LMAX; ++kf) {
          float s0, c0, s1, c1, s2, c2;
          sincosf(fr * xyz.x, &s0, &c0);
          sincosf(fr * xyz.y, &s1, &c1);
          sincosf(fr * xyz.z, &s2, &c2);
          if (kf == 10) { wr.push(0.f); wr = RowWriter<Op>(arena + CH_EXTRA * kAChunkBytes, row); }
          wr.push(s0); wr.push(s1); wr.push(s2); wr.push(c0); wr.push(c1); wr.push(c2);
          fr *= 2.0f;
        }
        if (LMAX <= 10) wr.push(0.f);  // zero column 63 of the PE chunk
        wr.pad_to(16);
      }

      // ------------------------------------------------ visibility MLP (VisField.forward)
      seq = P.seq_vis;
      run_gemm();
      epi_relu_store(seq, bias_ptr(seq, f), CH_H0);
      run_gemm();
      float vis_out;
      {
        const float* b2 = bias_ptr(seq + 1, f);
        float accv = __ldg(p.a.vis_final_b);
        for (int c0 = 0; c0 < 64; c0 += 32) {
          float v[32];
          tmem_ld32(t_lane + P.seq[seq + 1].tmem_col + c0, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) accv += fmaxf(v[j] + __ldg(b2 + c0 + j), 0.f) * __ldg(p.a.vis_final_w + c0 + j);
        }
        vis_out = accv;
      }

      // ------------------------------------------------ density branch (NeRF.forward, basefield + sdf)
      seq = P.seq_base;
      for (int i = 0; i < p.desc.D; ++i) {
        run_gemm();
        epi_relu_store(seq + i, bias_ptr(seq + i, f), CH_H0);
      }
      run_gemm();
      float sdf;
      {
        const int sf = seq + p.desc.D;
        const float* bb = bias_ptr(sf, f);
        float accs = __ldg(p.a.sdf_b);
        for (int c0 = 0; c0 < W; c0 += 32) {
          float v[32];
          tmem_ld32(t_lane + P.seq[sf].tmem_col + c0, v);
          uint8_t* chunk = arena + (CH_H0 + (c0 >> 6)) * kAChunkBytes;
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              y[j] = fmaxf(v[g8 * 8 + j] + __ldg(bb + c0 + g8 * 8 + j), 0.f);
              accs += y[j] * __ldg(p.a.sdf_w + c0 + g8 * 8 + j);
            }
            uint4 o = make_uint4(Op::pack2(y[0], y[1]), Op::pack2(y[2], y[3]), Op::pack2(y[4], y[5]), Op::pack2(y[6], y[7]));
            *reinterpret_cast<uint4*>(chunk + sw128_off(row, ((c0 & 63) >> 3) + g8)) = o;
          }
        }
        sdf = accs;
      }
      const float ibeta = expf(__ldg(p.a.logibeta));
      const float sgn = sdf > 0.f ? 1.f : (sdf < 0.f ? -1.f : 0.f);
      const float density = (0.5f + 0.5f * sgn * expm1f(-fabsf(sdf) * ibeta)) * ibeta;

      // ------------------------------------------------ colour branch: rgb.0 is linear in (base + colour)
      seq = P.seq_rgb1;
      run_gemm();  // base features x rgb.0 -> TMEM[kTmemRgb..)
      seq = P.seq_color;
      run_gemm();
      epi_relu_store(seq, bias_ptr(seq, f), CH_H0);
      run_gemm();
      epi_relu_store(seq + 1, bias_ptr(seq + 1, f), CH_H0);
      run_gemm();
      epi_relu_store(seq + 2, bias_ptr(seq + 2, f), CH_H0);
      seq = P.seq_rgb2;
      run_gemm();  // + colour features x rgb.0
      float rgb[3];
      {
        const float* b0 = bias_ptr(seq, f);
        const int H = W / 2;
        float a0 = __ldg(p.a.rgb2_b), a1 = __ldg(p.a.rgb2_b + 1), a2 = __ldg(p.a.rgb2_b + 2);
        for (int c0 = 0; c0 < H; c0 += 32) {
          float v[32];
          tmem_ld32(t_lane + kTmemRgb + c0, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float pre = v[j] + __ldg(b0 + c0 + j);
            if (p.desc.L_dir == 0) {
              const float* wd = p.a.rgb0_dir_w + (c0 + j) * 3;
              pre += __ldg(wd) * dir_f.x + __ldg(wd + 1) * dir_f.y + __ldg(wd + 2) * dir_f.z;
            }
            const float hh = fmaxf(pre, 0.f);
            a0 += hh * __ldg(p.a.rgb2_w + c0 + j);
            a1 += hh * __ldg(p.a.rgb2_w + H + c0 + j);
            a2 += hh * __ldg(p.a.rgb2_w + 2 * H + c0 + j);
          }
        }
        rgb[0] = 1.f / (1.f + expf(-a0)); rgb[1] = 1.f / (1.f + expf(-a1)); rgb[2] = 1.f / (1.f + expf(-a2));
      }

      // ------------------------------------------------ feature field (FeatureNeRF.compute_feat)
      float feat[16];
      if (p.desc.has_feature) {
        seq = P.seq_feat;
        for (int i = 0; i < 5; ++i) {
          run_gemm();
          epi_relu_store(seq + i, bias_ptr(seq + i, f), CH_H0);
        }
        run_gemm();
        float v16[16];
        tmem_ld16(t_lane + P.seq[seq + 5].tmem_col, v16);
        const float* bf = bias_ptr(seq + 5, f);
        float nn = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) { feat[j] = v16[j] + __ldg(bf + j); nn += feat[j] * feat[j]; }
        const float inv = 1.0f / sqrtf(nn);
#pragma unroll
        for (int j = 0; j < 16; ++j) feat[j] *= inv;
      }

      // ------------------------------------------------ flow + cycle warps (compute_flow, cycle_loss)
      float flow[3] = {0.f, 0.f, 0.f};
      float cyc = 0.f;
      float3 x_next = xyz;
      if (B > 0) {
        float e1, d1;
        seq = P.seq_delta_flow;
        x_next = skin_warp(xyz, p.a.bone_inv_rest + (size_t)fn * B * 8, p.a.se3_fwd + (size_t)fn * B * 8, seq,
                           p.a.delta1_bias_fwd + (size_t)f * 64, e1, d1);
        float e2, d2;
        seq = P.seq_delta_cyc;
        const float3 xc = skin_warp(xyz, p.a.bone_inv_rest + (size_t)f * B * 8, p.a.se3_fwd + (size_t)f * B * 8, seq,
                                    p.a.delta1_bias_fwd + (size_t)f * 64, e2, d2);
        const float dx = xc.x - xyz_t.x, dy = xc.y - xyz_t.y, dz = xc.z - xyz_t.z;
        cyc = sqrtf(dx * dx + dy * dy + dz * dz);
        ent_out = 0.5f * (e2 + ent_b);
        dsk_out = 0.5f * (d2 + dsk_b);
      }
      {
        // field_to_cam with the partner frame's camera, pinhole projection, flow (nerf.py:948-997)
        const Q4 qn = ldq(p.a.field2cam + (size_t)fn * 8);
        const float4 tn = __ldg(reinterpret_cast<const float4*>(p.a.field2cam + (size_t)fn * 8 + 4));
        float3 xc = qrot(qn, x_next);
        xc.x += tn.x; xc.y += tn.y; xc.z += tn.z;
        const float* Kn = p.a.Kinv + (size_t)fn * 9;
        const float k0 = __ldg(Kn + 0), k1 = __ldg(Kn + 4), k2 = __ldg(Kn + 2), k3 = __ldg(Kn + 5);
        const float fx = 1.0f / k0, fy = 1.0f / k1, cx = -k2 / k0, cy = -k3 / k1;
        const float hz = xc.z;
        const float hxn = (fx * xc.x + cx * xc.z) / (hz + 1e-6f);
        const float hyn = (fy * xc.y + cy * xc.z) / (hz + 1e-6f);
        flow[0] = hxn - h0;
        flow[1] = hyn - h1;
        bool valid = xc.z > 1e-6f;
        if (p.a.flow_thresh >= 0.f) valid = valid && (sqrtf(flow[0] * flow[0] + flow[1] * flow[1]) < p.a.flow_thresh);
        flow[2] = valid ? 1.f : 0.f;
      }

      // ------------------------------------------------ Gaussian bone density (compute_gauss_density)
      float gdens = 0.f;
      if (B > 0) {
        float best = -INFINITY;
#pragma unroll
        for (int b = 0; b < B; ++b) {
          const float4 c = __ldg(reinterpret_cast<const float4*>(p.a.bone_center + b * 4));
          const float dx = xyz.x - c.x, dy = xyz.y - c.y, dz = xyz.z - c.z;
          const float d2 = (dx * dx + dy * dy + dz * dz) / (0.01f * 0.01f);
          best = fmaxf(best, expf(-0.5f * d2));
        }
        gdens = best * expf(__ldg(p.a.warp_logibeta));
      }

      // ------------------------------------------------ per-sample outputs
      if (live) {
        const size_t o = (size_t)s;
        auto st3 = [&](float* dst, float a, float b, float c) { if (dst) { dst[o * 3] = a; dst[o * 3 + 1] = b; dst[o * 3 + 2] = c; } };
        auto st1 = [&](float* dst, float a) { if (dst) dst[o] = a; };
        st3(p.a.rgb, rgb[0], rgb[1], rgb[2]);
        st1(p.a.density, density);
        st1(p.a.sdf, sdf);
        st1(p.a.vis, vis_out);
        st3(p.a.xyz, xyz.x, xyz.y, xyz.z);
        st3(p.a.xyz_cam, xyz_cam.x, xyz_cam.y, xyz_cam.z);
        st3(p.a.xyz_t, xyz_t.x, xyz_t.y, xyz_t.z);
        st3(p.a.dir, dir_f.x, dir_f.y, dir_f.z);
        st1(p.a.depth, depth / expf(__ldg(p.a.logscale)));
        st1(p.a.deltas, delta);
        st3(p.a.flow, flow[0], flow[1], flow[2]);
        st1(p.a.cyc_dist, cyc);
        st1(p.a.delta_skin, dsk_out);
        st1(p.a.skin_entropy, ent_out);
        st1(p.a.gauss_density, gdens);
        if (p.a.feature && p.desc.has_feature) {
          float4* fo = reinterpret_cast<float4*>(p.a.feature + o * 16);
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
  if (warp == 5) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <class Op, int B, int LMAX>
static cudaError_t launch_one(const FieldKernelParams& p, int n_sm, cudaStream_t stream) {
  auto kern = field_fwd_kernel<Op, B, LMAX>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (e != cudaSuccess) return e;
  const int grid = p.n_tiles < n_sm ? p.n_tiles : n_sm;
  kern<<<grid, kThreads, kSmemBytes, stream>>>(p);
  return cudaGetLastError();
}

static cudaError_t launch_field_fwd(const FieldKernelParams& p, int n_sm, cudaStream_t stream) {
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

cudaError_t launch_field_fwd_desc(const b200r_field_desc& desc, const Program& prog, const b200r_field_args& args,
                                  const void* packed, int n_sm, cudaStream_t stream) {
  FieldKernelParams p;
  memset(&p, 0, sizeof(p));
  p.prog = prog;
  p.desc = desc;
  p.a = args;
  p.packed = reinterpret_cast<const uint8_t*>(packed);
  for (int i = 0; i < prog.n_seq; ++i) {
    p.bias_seq[i] = args.bias[prog.seq[i].layer];
    p.bias_stride_seq[i] = args.bias_stride[prog.seq[i].layer];
  }
  p.S = args.M * args.N * args.D;
  p.n_tiles = (p.S + kTileRows - 1) / kTileRows;
  p.Lmax = desc.L_xyz + 2 > 10 ? 12 : 10;
  if (desc.L_xyz + 2 > 10 && desc.L_xyz != 10) return cudaErrorInvalidValue;  // only L_xyz <= 8 or == 10
  return launch_field_fwd(p, n_sm, stream);
}

}  // namespace b200r
