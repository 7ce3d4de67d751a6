// Host+device description of the per-tile tensor-core program of one field.
//
// A "tile" is 128 consecutive ray-samples.  Its activations live in an arena of K-major
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
constexpr int kChunkK = 64;                        // halves per operand row (128 B)
constexpr int kAChunkBytes = kTileRows * 128;      // 16 KB
constexpr int kMaxN = 256;
constexpr int kWStageBytes = kMaxN * 128;          // 32 KB weight ring stage
constexpr int kMaxSeq = 48;
constexpr int kMaxKChunks = 6;

// arena chunk ids
enum : int { CH_PE = 0, CH_EXTRA = 1, CH_H0 = 2, CH_H1 = 3, CH_H2 = 4, CH_H3 = 5, kArenaChunks = 6 };

// TMEM columns
constexpr int kTmemCols = 512;
constexpr int kTmemMain = 0;
constexpr int kTmemRgb = 256;

struct GemmDesc {
  uint32_t w_off;     // byte offset of the first packed chunk
  uint16_t n_pad;     // UMMA N (multiple of 16, <= 256)
  uint16_t tmem_col;  // accumulator column offset
  uint8_t n_chunks;
  uint8_t accumulate;  // 1: first MMA adds onto the existing accumulator
  uint8_t layer;       // canonical layer id (bias lookup)
  uint8_t pad_;
  uint8_t a_chunk[kMaxKChunks];
  uint8_t ksteps[kMaxKChunks];
};

// canonical layer ids inside a Program (indices into Program::layer_of)
struct LayerIds {
  int delta[3];    // -1 when absent
  int vis[2];
  int base[10];    // linear_1..D, final
  int rgb0;
  int color[3];
  int feat[6];
  int count;
};

struct Program {
  int32_t n_seq;
  // positions in seq[] of the phases the compute warps walk through
  int32_t seq_delta_bwd, seq_vis, seq_base, seq_rgb1, seq_color, seq_rgb2, seq_feat, seq_delta_flow, seq_delta_cyc;
  GemmDesc seq[kMaxSeq];
};

// one source slice of a weight matrix that fills one packed K chunk
struct PackSlice {
  int layer;     // canonical layer id
  int n;         // valid output rows
  int n_pad;
  int in_dim;    // leading dimension of the fp32 source
  int col0;      // first source column
  int ncols;     // valid columns (<= 64), rest zero
  int pe_window; // 0: none, 1: basefield window (L_xyz freqs), 2: colorfield window (L_xyz+2); applies to PE columns
  int pe_col0;   // index of this slice's first column inside the positional embedding (for the window)
  uint32_t dst_off;
};

inline int pe_dim(int L) { return L < 0 ? 0 : 3 * (2 * L + 1); }
inline int pad16(int n) { return (n + 15) / 16 * 16; }

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
  if (d.L_xyz < 1 || d.L_xyz > 10) { bp.err = "L_xyz must be in [1,10]"; return bp; }
  if (!(d.L_dir == -1 || d.L_dir == 0)) { bp.err = "L_dir must be -1 or 0"; return bp; }
  if (!(d.n_bones == 0 || d.n_bones == 18 || d.n_bones == 25)) { bp.err = "n_bones must be 0, 18 or 25"; return bp; }
  if (d.operand_dtype != 0 && d.operand_dtype != 1) { bp.err = "operand_dtype must be 0 or 1"; return bp; }
  const LayerIds L = layer_ids(d);
  bp.layer_out.assign(L.count, 0);
  bp.layer_in.assign(L.count, 0);
  const int INST = 32, TEMB = 128;
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

  if (d.n_bones > 0) {
    const int xb = 3 * d.n_bones, in1 = xb + TEMB + INST;
    std::vector<PackSlice> v;
    v.push_back(sl(0, xb < 64 ? xb : 64));
    if (xb > 64) v.push_back(sl(64, xb - 64));
    add_layer(L.delta[0], 64, in1, v);
    add_layer(L.delta[1], 64, 64, hidden(0, 64));
    add_layer(L.delta[2], d.n_bones, 64, hidden(0, 64));
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

  // ---- per-tile sequence
  int ns = 0;
  auto emit = [&](int id, const std::vector<int>& a_chunks, int tmem_col, int accumulate) {
    GemmDesc& g = P.seq[ns++];
    g = GemmDesc{};
    g.w_off = layer_off[id];
    g.n_pad = (uint16_t)pad16(bp.layer_out[id]);
    g.tmem_col = (uint16_t)tmem_col;
    g.accumulate = (uint8_t)accumulate;
    g.layer = (uint8_t)id;
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
  auto emit_delta = [&]() {
    const int xb = 3 * d.n_bones;
    emit(L.delta[0], xb > 64 ? std::vector<int>{CH_H0, CH_H1} : std::vector<int>{CH_H0}, kTmemMain, 0);
    emit(L.delta[1], {CH_H2}, kTmemMain, 0);
    emit(L.delta[2], {CH_H2}, kTmemMain, 0);
  };
  P.seq_delta_bwd = ns;
  if (d.n_bones > 0) emit_delta();
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
  P.seq_delta_flow = ns;
  if (d.n_bones > 0) emit_delta();
  P.seq_delta_cyc = ns;
  if (d.n_bones > 0) emit_delta();
  P.n_seq = ns;
  bp.ok = true;
  return bp;
}

}  // namespace b200r
