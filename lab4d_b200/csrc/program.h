// Host+device description of the per-tile tensor-core program of one field and of the scratch
// blocks (constant block, per-frame blocks) the prologue kernel builds for it.
//
// A "tile" is 128 consecutive ray-samples of ONE frame.  Its activations live in an arena of K-major
// SWIZZLE_128B operand chunks ([128 rows x 64 halves] = 16 KB each) in shared memory; every dense
// layer is one GemmDesc: D[128 x n_pad] (+)= sum over K chunks A_chunk[128 x 16*ksteps] * W_chunk^T.
// Weights are pre-packed by b200r_pack_weights into [n_pad x 64] chunks with the same swizzle, so a
// plain 1-D TMA bulk copy lands them in shared memory ready for tcgen05.mma.
#pragma once
#include <stdint.h>

#include <vector>

#include "../../include/b200r.h"

namespace b200r {

constexpr int kTileRows = 128;
constexpr int kAChunkBytes = kTileRows * 128;  // 16 KB
constexpr int kMaxN = 256;
constexpr int kWStageBytes = kMaxN * 128;      // 32 KB weight ring stage
constexpr int kMaxSeq = 48;
constexpr int kMaxKChunks = 6;
constexpr int kMaxCond = 8;

// arena chunk ids
enum : int { CH_PE = 0, CH_EXTRA = 1, CH_H0 = 2, CH_H1 = 3, CH_H2 = 4, CH_H3 = 5, kArenaChunks = 6 };

// TMEM columns
constexpr int kTmemCols = 512;
constexpr int kTmemMain = 0;
constexpr int kTmemRgb = 256;

struct GemmDesc {
  uint32_t w_off;      // byte offset of the first packed chunk
  uint16_t n_pad;      // UMMA N (multiple of 16, <= 256)
  uint16_t tmem_col;   // accumulator column offset
  uint16_t bias_off;   // float offset of the bias row inside the constant block or the frame block
  uint8_t bias_frame;  // 1: bias row lives in the per-frame block
  uint8_t n_chunks;
  uint8_t accumulate;  // 1: first MMA adds onto the existing accumulator
  uint8_t layer;       // canonical layer id
  uint8_t a_chunk[kMaxKChunks];
  uint8_t ksteps[kMaxKChunks];
};

struct LayerIds {
  int delta[3];  // -1 when absent
  int vis[2];
  int base[10];  // linear_1..D, final
  int rgb0;
  int color[3];
  int feat[6];
  int count;
};

// code tables a conditioned layer can take its per-frame code from
enum : int { CODE_INST_BASE = 0, CODE_INST_COLOR, CODE_INST_VIS, CODE_APPR, CODE_INST_SKIN, CODE_T_EMBED, CODE_T_EMBED_MEAN, kNumCodes };

// bias row that depends on the frame: b + sum_seg W[:, col0:col0+C] @ code_seg[frame]
struct CondRow {
  int16_t layer;      // canonical layer (weight / bias source)
  int16_t n;          // rows
  int16_t in_dim;
  int16_t frame_off;  // float offset inside the frame block
  int16_t n_seg;
  int16_t col0[2], width[2], code[2];
};

// float offsets inside the constant block
struct ConstLayout {
  int16_t sdf_w, rgb2_w, vis_w, dir_w, inv_gauss, center, scalars;  // scalars: see SC_* below
  int16_t n_floats;
  int16_t plain_off[B200R_MAX_LAYERS];  // bias row of non-conditioned layers, -1 otherwise
};
enum : int { SC_IBETA = 0, SC_INV_SCALE, SC_WARP_IBETA, SC_SDF_B, SC_RGB2_B0, SC_RGB2_B1, SC_RGB2_B2, SC_VIS_B, kNumScalars = 8 };

// float offsets inside one frame block
struct FrameLayout {
  int16_t cam, cam_partner;  // 24 floats each: Kinv[9], near, far, q[4], t[3], pad
  // bone tables: binv_* = scaled inverse bone transform as 3 rows (R'_i0 R'_i1 R'_i2 t'_i), B*12 floats;
  // se3_* = blend transform as dual quaternion (real, dual), B*8 floats
  int16_t binv_t, se3_bwd, binv_rest, se3_fwd, binv_rest_partner, se3_fwd_partner;
  int16_t delta1_fwd;        // bias row of delta_field.linear_1 with the MEAN time code
  int16_t n_cond;
  int16_t n_floats;          // multiple of 4
  int16_t pad_;
  CondRow cond[kMaxCond];
};

struct Program {
  int32_t n_seq;
  // positions in seq[] of the phases the compute warps walk through
  int32_t seq_delta_bwd, seq_vis, seq_base, seq_rgb1, seq_color, seq_rgb2, seq_feat, seq_delta_flow, seq_delta_cyc;
  ConstLayout cl;
  FrameLayout fl;
  GemmDesc seq[kMaxSeq];
};

// one source slice of a weight matrix that fills one packed K chunk
struct PackSlice {
  int layer;      // canonical layer id
  int n;          // valid output rows
  int n_pad;
  int in_dim;     // leading dimension of the fp32 source
  int col0;       // first source column
  int ncols;      // valid columns (<= 64), rest zero
  int pe_window;  // 0: none, 1: basefield window (L_xyz freqs), 2: colorfield window (L_xyz+2)
  int pe_col0;    // index of this slice's first column inside the positional embedding
  uint32_t dst_off;
};

inline int pe_dim(int L) { return L < 0 ? 0 : 3 * (2 * L + 1); }
inline int pad16(int n) { return (n + 15) / 16 * 16; }
inline int pad4(int n) { return (n + 3) / 4 * 4; }

inline LayerIds layer_ids(const b200r_field_desc& d) {
  LayerIds L;
  int c = 0;
  for (int i = 0; i < 3; ++i) L.delta[i] = d.n_bones > 0 ? c++ : -1;
  for (int i = 0; i < 2; ++i) L.vis[i] = c++;
  for (int i = 0; i < 10; ++i) L.base[i] = i <= d.D ? c++ : -1;
  L.rgb0 = c++;
  for (int i = 0; i < 3; ++i) L.color[i] = c++;
  for (int i = 0; i < 6; ++i) L.feat[i] = d.has_feature ? c++ : -1;
  L.count = c;
  return L;
}

struct BuiltProgram {
  Program prog;
  std::vector<PackSlice> slices;
  std::vector<int> layer_out;  // N of each canonical layer
  std::vector<int> layer_in;   // fp32 source in_dim of each canonical layer
  size_t packed_bytes;
  bool ok;
  const char* err;
};

inline BuiltProgram build_program(const b200r_field_desc& d) {
  BuiltProgram bp;
  bp.ok = false;
  bp.err = "";
  bp.packed_bytes = 0;
  Program& P = bp.prog;
  P = Program{};
  if (!(d.W == 256 || d.W == 128)) { bp.err = "W must be 128 or 256"; return bp; }
  if (d.D < 2 || d.D > 9) { bp.err = "D out of range"; return bp; }
  if (d.skip < 1 || d.skip >= d.D) { bp.err = "skip must satisfy 1 <= skip < D"; return bp; }
  if (!(d.L_xyz == 10 || (d.L_xyz >= 1 && d.L_xyz <= 8))) { bp.err = "L_xyz must be 10 or in [1,8]"; return bp; }
  if (!(d.L_dir == -1 || d.L_dir == 0)) { bp.err = "L_dir must be -1 or 0"; return bp; }
  if (!(d.n_bones == 0 || d.n_bones == 18 || d.n_bones == 25)) { bp.err = "n_bones must be 0, 18 or 25"; return bp; }
  if (d.n_bones > 0 && d.L_xyz != 10) { bp.err = "skinned fields need L_xyz == 10"; return bp; }
  if (d.appr_channels < 0 || d.appr_channels > 64) { bp.err = "appr_channels out of range"; return bp; }
  if (d.operand_dtype != 0 && d.operand_dtype != 1) { bp.err = "operand_dtype must be 0 or 1"; return bp; }
  const LayerIds L = layer_ids(d);
  bp.layer_out.assign(L.count, 0);
  bp.layer_in.assign(L.count, 0);
  const int INST = 32, TEMB = 128, B = d.n_bones;
  const int pe_b = pe_dim(d.L_xyz), pe_c = pe_dim(d.L_xyz + 2), pe_v = pe_dim(10), pe_f = pe_dim(6);
  const int hw = d.W / 64;  // hidden chunks
  uint32_t off = 0;
  std::vector<uint32_t> layer_off(L.count, 0);
  std::vector<std::vector<PackSlice>> layer_slices(L.count);

  auto add_layer = [&](int id, int n, int in_dim, std::vector<PackSlice> sl) {
    bp.layer_out[id] = n;
    bp.layer_in[id] = in_dim;
    layer_off[id] = off;
    for (auto& s : sl) {
      s.layer = id;
      s.n = n;
      s.n_pad = pad16(n);
      s.in_dim = in_dim;
      s.dst_off = off;
      off += (uint32_t)s.n_pad * 128u;
      bp.slices.push_back(s);
    }
    layer_slices[id] = sl;
  };
  auto sl = [](int col0, int ncols, int win = 0, int pe_col0 = 0) {
    PackSlice s{};
    s.col0 = col0; s.ncols = ncols; s.pe_window = win; s.pe_col0 = pe_col0;
    return s;
  };
  auto hidden = [&](int col0, int width) {
    std::vector<PackSlice> v;
    for (int j = 0; j < width / 64; ++j) v.push_back(sl(col0 + 64 * j, 64));
    return v;
  };
  auto pe_slices = [&](int pe_n, int win) {  // embedding columns [0,pe_n) -> CH_PE (first 63) + CH_EXTRA
    std::vector<PackSlice> v;
    v.push_back(sl(0, pe_n < 63 ? pe_n : 63, win, 0));
    if (pe_n > 63) v.push_back(sl(63, pe_n - 63, win, 63));
    return v;
  };
  auto cat = [](std::vector<PackSlice> a, const std::vector<PackSlice>& b) {
    a.insert(a.end(), b.begin(), b.end());
    return a;
  };

  if (B > 0) {
    const int xb = 3 * B, in1 = xb + TEMB + INST;
    std::vector<PackSlice> v;
    v.push_back(sl(0, xb < 64 ? xb : 64));
    if (xb > 64) v.push_back(sl(64, xb - 64));
    add_layer(L.delta[0], 64, in1, v);
    add_layer(L.delta[1], 64, 64, hidden(0, 64));
    add_layer(L.delta[2], B, 64, hidden(0, 64));
  }
  add_layer(L.vis[0], 64, pe_v + INST, pe_slices(pe_v, 0));
  add_layer(L.vis[1], 64, 64, hidden(0, 64));
  for (int i = 0; i < d.D; ++i) {
    if (i == 0) add_layer(L.base[i], d.W, pe_b + INST, pe_slices(pe_b, 1));
    else if (i == d.skip) add_layer(L.base[i], d.W, pe_b + INST + d.W, cat(pe_slices(pe_b, 1), hidden(pe_b + INST, d.W)));
    else add_layer(L.base[i], d.W, d.W, hidden(0, d.W));
  }
  add_layer(L.base[d.D], d.W, d.W, hidden(0, d.W));
  add_layer(L.rgb0, d.W / 2, d.W + pe_dim(d.L_dir) + d.appr_channels, hidden(0, d.W));
  add_layer(L.color[0], d.W, pe_c + INST, pe_slices(pe_c, 2));
  add_layer(L.color[1], d.W, d.W, hidden(0, d.W));
  add_layer(L.color[2], d.W, d.W, hidden(0, d.W));
  if (d.has_feature) {
    for (int i = 0; i < 5; ++i) {
      if (i == 0) add_layer(L.feat[i], 128, pe_f, pe_slices(pe_f, 0));
      else if (i == 4) add_layer(L.feat[i], 128, pe_f + 128, cat(pe_slices(pe_f, 0), hidden(pe_f, 128)));
      else add_layer(L.feat[i], 128, 128, hidden(0, 128));
    }
    add_layer(L.feat[5], 16, 128, hidden(0, 128));
  }
  bp.packed_bytes = off;

  // ---- frame block: cameras, conditioned bias rows, bone tables
  FrameLayout& F = P.fl;
  int fo = 0;
  F.cam = (int16_t)fo; fo += 24;
  F.cam_partner = (int16_t)fo; fo += 24;
  std::vector<int> cond_off(L.count, -1);
  int nc = 0;
  auto add_cond = [&](int layer, int c0a, int wa, int codea, int c0b = 0, int wb = 0, int codeb = 0) {
    CondRow& c = F.cond[nc++];
    c.layer = (int16_t)layer; c.n = (int16_t)bp.layer_out[layer]; c.in_dim = (int16_t)bp.layer_in[layer];
    c.frame_off = (int16_t)fo;
    c.n_seg = (int16_t)(wb > 0 ? 2 : 1);
    c.col0[0] = (int16_t)c0a; c.width[0] = (int16_t)wa; c.code[0] = (int16_t)codea;
    c.col0[1] = (int16_t)c0b; c.width[1] = (int16_t)wb; c.code[1] = (int16_t)codeb;
    cond_off[layer] = fo;
    fo += pad4(bp.layer_out[layer]);
  };
  F.delta1_fwd = -1;
  if (B > 0) {
    add_cond(L.delta[0], 3 * B, TEMB, CODE_T_EMBED, 3 * B + TEMB, INST, CODE_INST_SKIN);
    F.delta1_fwd = (int16_t)fo;
    // forward-warp variant of the same layer (mean time code); not referenced through cond_off
    CondRow& c = F.cond[nc++];
    c = F.cond[nc - 2];
    c.frame_off = (int16_t)fo;
    c.code[0] = CODE_T_EMBED_MEAN;
    fo += 64;
  }
  add_cond(L.vis[0], pe_v, INST, CODE_INST_VIS);
  add_cond(L.base[0], pe_b, INST, CODE_INST_BASE);
  add_cond(L.base[d.skip], pe_b, INST, CODE_INST_BASE);
  if (d.appr_channels > 0) add_cond(L.rgb0, d.W + pe_dim(d.L_dir), d.appr_channels, CODE_APPR);
  add_cond(L.color[0], pe_c, INST, CODE_INST_COLOR);
  F.n_cond = (int16_t)nc;
  auto bones = [&](int per) { int o = fo; fo += B * per; return (int16_t)o; };
  F.binv_t = bones(12); F.se3_bwd = bones(8); F.binv_rest = bones(12); F.se3_fwd = bones(8);
  F.binv_rest_partner = bones(12); F.se3_fwd_partner = bones(8);
  F.n_floats = (int16_t)pad4(fo);

  // ---- constant block: plain bias rows, head weights, Gaussian scales, scalars
  ConstLayout& C = P.cl;
  int co = 0;
  for (int i = 0; i < B200R_MAX_LAYERS; ++i) C.plain_off[i] = -1;
  for (int i = 0; i < L.count; ++i)
    if (cond_off[i] < 0) { C.plain_off[i] = (int16_t)co; co += pad16(bp.layer_out[i]); }
  C.sdf_w = (int16_t)co; co += d.W;
  C.rgb2_w = (int16_t)co; co += 3 * (d.W / 2);
  C.vis_w = (int16_t)co; co += 64;
  C.dir_w = (int16_t)co; co += (d.L_dir == 0) ? pad4(3 * (d.W / 2)) : 0;
  C.inv_gauss = (int16_t)co; co += B * 4;
  C.center = (int16_t)co; co += B * 4;
  C.scalars = (int16_t)co; co += kNumScalars;
  C.n_floats = (int16_t)pad4(co);

  // ---- per-tile sequence
  int ns = 0;
  auto emit = [&](int id, const std::vector<int>& a_chunks, int tmem_col, int accumulate, int bias_override = -1) {
    GemmDesc& g = P.seq[ns++];
    g = GemmDesc{};
    g.w_off = layer_off[id];
    g.n_pad = (uint16_t)pad16(bp.layer_out[id]);
    g.tmem_col = (uint16_t)tmem_col;
    g.accumulate = (uint8_t)accumulate;
    g.layer = (uint8_t)id;
    if (bias_override >= 0) { g.bias_frame = 1; g.bias_off = (uint16_t)bias_override; }
    else if (cond_off[id] >= 0) { g.bias_frame = 1; g.bias_off = (uint16_t)cond_off[id]; }
    else { g.bias_frame = 0; g.bias_off = (uint16_t)C.plain_off[id]; }
    const auto& S = layer_slices[id];
    g.n_chunks = (uint8_t)S.size();
    for (size_t c = 0; c < S.size(); ++c) {
      g.a_chunk[c] = (uint8_t)a_chunks[c];
      g.ksteps[c] = (uint8_t)((S[c].ncols + 15) / 16);
    }
  };
  auto hch = [&](int first, int n) { std::vector<int> v; for (int j = 0; j < n; ++j) v.push_back(first + j); return v; };
  auto pe_ch = [&](int pe_n) { std::vector<int> v{CH_PE}; if (pe_n > 63) v.push_back(CH_EXTRA); return v; };
  auto catv = [](std::vector<int> a, const std::vector<int>& b) { a.insert(a.end(), b.begin(), b.end()); return a; };
  auto emit_delta = [&](bool fwd) {
    const int xb = 3 * B;
    emit(L.delta[0], xb > 64 ? std::vector<int>{CH_H0, CH_H1} : std::vector<int>{CH_H0}, kTmemMain, 0, fwd ? F.delta1_fwd : -1);
    emit(L.delta[1], {CH_H2}, kTmemMain, 0);
    emit(L.delta[2], {CH_H2}, kTmemMain, 0);
  };
  P.seq_delta_bwd = ns;
  if (B > 0) emit_delta(false);
  P.seq_delta_flow = ns;
  if (B > 0) emit_delta(true);
  P.seq_delta_cyc = ns;
  if (B > 0) emit_delta(true);
  P.seq_vis = ns;
  emit(L.vis[0], pe_ch(pe_v), kTmemMain, 0);
  emit(L.vis[1], {CH_H0}, kTmemMain, 0);
  P.seq_base = ns;
  for (int i = 0; i <= d.D; ++i) {
    if (i == 0) emit(L.base[i], pe_ch(pe_b), kTmemMain, 0);
    else if (i == d.skip) emit(L.base[i], catv(pe_ch(pe_b), hch(CH_H0, hw)), kTmemMain, 0);
    else emit(L.base[i], hch(CH_H0, hw), kTmemMain, 0);
  }
  P.seq_rgb1 = ns;
  emit(L.rgb0, hch(CH_H0, hw), kTmemRgb, 0);
  P.seq_color = ns;
  emit(L.color[0], pe_ch(pe_c), kTmemMain, 0);
  emit(L.color[1], hch(CH_H0, hw), kTmemMain, 0);
  emit(L.color[2], hch(CH_H0, hw), kTmemMain, 0);
  P.seq_rgb2 = ns;
  emit(L.rgb0, hch(CH_H0, hw), kTmemRgb, 1);
  P.seq_feat = ns;
  if (d.has_feature) {
    for (int i = 0; i < 5; ++i) {
      if (i == 0) emit(L.feat[i], pe_ch(pe_f), kTmemMain, 0);
      else if (i == 4) emit(L.feat[i], catv(pe_ch(pe_f), hch(CH_H0, 2)), kTmemMain, 0);
      else emit(L.feat[i], hch(CH_H0, 2), kTmemMain, 0);
    }
    emit(L.feat[5], hch(CH_H0, 2), kTmemMain, 0);
  }
  P.n_seq = ns;
  bp.ok = true;
  return bp;
}

inline size_t workspace_floats(const Program& P, int M) { return (size_t)P.cl.n_floats + (size_t)M * P.fl.n_floats; }

}  // namespace b200r
