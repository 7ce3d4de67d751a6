// Host+device description of the per-tile tensor-core program of one field and of the scratch
// blocks (constant block, per-frame blocks) the prologue kernel builds for it.
//
// A "tile" is 128 consecutive ray-samples of ONE frame; a CTA keeps two tiles in flight (tile groups 0 and 1).
// Per group: the Fourier embedding of the tile lives in two K-major SWIZZLE_128B operand chunks in shared memory
// ([128 rows x 64 halves] = 16 KB each), hidden activations live in TMEM (128 columns = 256 halves per row) next
// to a 128-column fp32 accumulator.  Every dense layer is a list of MMA steps D[128 x n] (+)= A * W_chunk^T.
// Weights are pre-packed by b200r_pack_weights into [n x 64] chunks with the same swizzle, so a plain 1-D TMA
// bulk copy lands them in shared memory ready for tcgen05.mma.
#pragma once
#include <stdint.h>

#include <vector>

#include "../../include/b200r.h"

namespace b200r {

constexpr int kTileRows = 128;
constexpr int kAChunkBytes = kTileRows * 128;  // 16 KB
constexpr int kMaxN = 256;
constexpr int kWStageBytes = 2 * 128 * 128;    // 32 KB weight ring stage (two [<=128 rows x 64] tiles)
constexpr int kMaxSteps = 208;  // fg + dense warp in split mode: one ring slot per 64-wide K chunk
constexpr int kMaxCond = 12;

// embedding operand chunks of one tile group (embedding columns 0..62 and 63..)
enum : int { CH_PE = 0, CH_EXTRA = 1, kArenaChunks = 2 };

constexpr int kTmemCols = 512;  // the CTA owns all of TMEM: per group 128 accumulator + 128 activation columns

// One MMA step = one ring stage = one or two packed weight tiles [n x 64] (= up to 8 UMMA_K steps of
// D[128 x n] (+)= A * W^T).  The A operand of each tile is either an arena chunk in shared memory (SS) or
// 32 TMEM columns of 16-bit activations written by the previous epilogue (TS; the second tile's columns follow).
struct MmaStep {
  uint32_t w_off;        // byte offset of the first packed tile (the second follows it)
  uint16_t n;            // UMMA N (multiple of 16, <= 256)
  uint16_t d_col;        // unused (each group has one accumulator)
  uint16_t a_tmem_col;   // TS: column inside the group's activation buffer of the operand's first k-step
  uint8_t a_kind;        // 0 = shared-memory chunk, 1 = TMEM
  uint8_t n_sub;         // 1 or 2 weight tiles in this step
  uint8_t a_chunk;       // SS: arena chunk id of tile 0
  uint8_t a_chunk2;      // SS: arena chunk id of tile 1
  uint8_t ksteps;        // UMMA_K steps of tile 0 (1..4)
  uint8_t ksteps2;       // UMMA_K steps of tile 1
  uint8_t pad_[2];
  uint8_t accumulate;    // 1: first k-step adds onto the existing accumulator
  uint8_t wait;          // BAR_* the MMA thread waits on before issuing this step (0 = none)
  uint8_t commit;        // BAR_* committed (arrive when the MMAs so far are done) after this step (0 = none)
};
// barriers between the compute warps and the MMA thread (ids shared by both directions)
enum : int { BAR_NONE = 0, BAR_ALL = 1, BAR_H0 = 2, BAR_H1 = 3 };

struct LayerBias { uint16_t off; uint8_t frame; uint8_t pad_; };  // float offset in const / frame block

struct LayerIds {
  int delta[3];  // -1 when absent
  int vis[2];
  int base[10];  // linear_1..D, final
  int rgb0;
  int color[3];
  int feat[6];
  int dense[6];  // forward_map L1, L2, final; backward_map L1, L2, final
  int count;
};

// code tables a conditioned layer can take its per-frame code from
enum : int { CODE_INST_BASE = 0, CODE_INST_COLOR, CODE_INST_VIS, CODE_APPR, CODE_INST_SKIN, CODE_T_EMBED, CODE_T_EMBED_MEAN,
              CODE_DENSE_T, CODE_DENSE_T_PARTNER, CODE_INST_DENSE_FWD, CODE_INST_DENSE_BWD, kNumCodes };

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
  int16_t sdf_w, rgb2_w, vis_w, dir_w, center, scalars, pad_;  // scalars: see SC_* below
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
  int16_t dense1_partner;    // bias row of post_warp.forward_map.linear_1 with the PARTNER frame's time code
  int16_t n_cond;
  int16_t n_floats;          // multiple of 4
  CondRow cond[kMaxCond];
};

// One block = one GEMM (or one N-half of a wide layer) as the MMA issuers see it: an optional first ring slot
// whose A operand is one or two embedding chunks in shared memory, then `ts_slots` slots whose A operand is the
// group's activation buffer in TMEM, read front to back (4 + 4 k-steps per slot; the last slot has 4 + ts_ks2_last).
// Split-operand mode (operand_dtype 2): every 64-wide K chunk is its own slot holding the weight tile's fp16 head
// and tail ([n x 64] each); the embedding chunks are one slot each, ts_slots counts chunks and ts_ks2_last is the
// number of k-steps of the LAST chunk.
struct MmaBlock {
  uint8_t n16;          // UMMA N / 16
  uint8_t ss;           // 0: no shared-memory slot; else 0x80 | ksteps | ksteps2 << 3
  uint8_t ss_chunks;    // arena chunk of tile 0 | chunk of tile 1 << 4
  uint8_t ts_slots;
  uint8_t ts_ks2_last;
  uint8_t wait, commit;
  uint8_t pad_;
};

struct Program {
  int32_t n_steps;
  int32_t n_blocks;
  // first step of each phase, in execution order: 3 warps (dense MLP + skinning delta MLP, see the kernel),
  // vis, feature, base chain, colour chain, final rgb.0
  int32_t st_delta[3], st_vis, st_feat, st_base, st_color, st_rgb;
  ConstLayout cl;
  FrameLayout fl;
  LayerBias bias[B200R_MAX_LAYERS];
  MmaStep steps[kMaxSteps];
  MmaBlock blocks[kMaxSteps];
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
  int row0;       // first output row of this chunk (N-halves of the pipelined layers)
  uint32_t dst_off;
  int transpose;  // 1 (dgrad operands): tile element (row r, col c) = W[col0 + c][row0 + r] - rows are IN-features
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
  for (int i = 0; i < 6; ++i) L.dense[i] = d.dense ? c++ : -1;
  L.count = c;
  return L;
}

// one packed K chunk of a layer: offset of its (first) tile in the operand buffer, padded N, UMMA_K steps
struct ChunkRef { uint32_t w_off; int n; int ksteps; };

struct BuiltProgram {
  Program prog;
  std::vector<PackSlice> slices;
  std::vector<std::vector<ChunkRef>> fwd_chunks, fwd_chunks_h1;  // per canonical layer (N-half 1 of the pipelined layers in _h1)
  std::vector<int> layer_out;  // N of each canonical layer
  std::vector<int> layer_in;   // fp32 source in_dim of each canonical layer
  size_t packed_bytes;
  bool ok;
  const char* err;
};

// what the per-tile step list evaluates (the packed operands are the same for every mode)
enum : int { MODE_FIELD = 0, MODE_POINTS = 1, MODE_WARP_BWD = 2, MODE_WARP_FWD = 3 };

inline BuiltProgram build_program(const b200r_field_desc& d, int mode = MODE_FIELD) {
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
  if (d.operand_dtype < 0 || d.operand_dtype > 2) { bp.err = "operand_dtype must be 0 (fp16), 1 (bf16) or 2 (fp16 head+tail)"; return bp; }
  const bool split = d.operand_dtype == 2;  // every packed tile is followed by the fp16 tail of its rounding error
  if (d.dense != 0 && d.dense != 1) { bp.err = "dense must be 0 or 1"; return bp; }
  if (d.dense && d.n_bones == 0) { bp.err = "dense (ComposedWarp) needs a skinned field"; return bp; }
  if (d.n_bones > 0 && d.W != 256) { bp.err = "skinned fields are built for W == 256 only"; return bp; }
  const LayerIds L = layer_ids(d);
  bp.layer_out.assign(L.count, 0);
  bp.layer_in.assign(L.count, 0);
  const int INST = 32, TEMB = 128, B = d.n_bones;
  const int pe_b = pe_dim(d.L_xyz), pe_c = pe_dim(d.L_xyz + 2), pe_v = pe_dim(10), pe_f = pe_dim(6);
  const int W = d.W, HN = W / 2, KC = W / 64;  // half width of the pipelined layers, hidden K chunks
  uint32_t off = 0;
  // chunk lists per layer; pipelined layers keep one list per N-half
  using Chunk = ChunkRef;
  std::vector<std::vector<Chunk>> chunks(L.count), chunks_h1(L.count);

  auto sl = [](int col0, int ncols, int win = 0) {
    PackSlice s{};
    s.col0 = col0; s.ncols = ncols; s.pe_window = win; s.pe_col0 = col0;
    return s;
  };
  auto hidden = [&](int col0, int width) {
    std::vector<PackSlice> v;
    for (int j = 0; j < width / 64; ++j) v.push_back(sl(col0 + 64 * j, 64));
    return v;
  };
  auto pe_slices = [&](int pe_n, int win) {  // embedding columns [0,pe_n) -> CH_PE (first 63) + CH_EXTRA
    std::vector<PackSlice> v;
    v.push_back(sl(0, pe_n < 63 ? pe_n : 63, win));
    if (pe_n > 63) v.push_back(sl(63, pe_n - 63, win));
    return v;
  };
  auto cat = [](std::vector<PackSlice> a, const std::vector<PackSlice>& b) {
    a.insert(a.end(), b.begin(), b.end());
    return a;
  };
  // pack rows [row0, row0+rows) of layer `id` as one chunk per slice
  auto add_rows = [&](int id, int n_total, int in_dim, int row0, int rows, std::vector<PackSlice> sls, std::vector<Chunk>& out) {
    bp.layer_out[id] = n_total;
    bp.layer_in[id] = in_dim;
    for (auto& s : sls) {
      s.layer = id;
      s.row0 = row0;
      s.n = (n_total - row0) < rows ? (n_total - row0) : rows;  // valid rows in this chunk
      s.n_pad = pad16(rows);
      s.in_dim = in_dim;
      s.dst_off = off;
      out.push_back({off, s.n_pad, (s.ncols + 15) / 16});
      off += (uint32_t)s.n_pad * 128u * (split ? 2u : 1u);
      bp.slices.push_back(s);
    }
  };
  auto add_layer = [&](int id, int n, int in_dim, const std::vector<PackSlice>& sls) { add_rows(id, n, in_dim, 0, n, sls, chunks[id]); };
  auto add_split = [&](int id, int in_dim, const std::vector<PackSlice>& sls) {  // two N-halves of a W-wide layer
    add_rows(id, W, in_dim, 0, HN, sls, chunks[id]);
    add_rows(id, W, in_dim, HN, HN, sls, chunks_h1[id]);
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
    if (i == 0) add_split(L.base[i], pe_b + INST, pe_slices(pe_b, 1));
    else if (i == d.skip) add_split(L.base[i], pe_b + INST + W, cat(pe_slices(pe_b, 1), hidden(pe_b + INST, W)));
    else add_split(L.base[i], W, hidden(0, W));
  }
  add_split(L.base[d.D], W, hidden(0, W));
  add_layer(L.rgb0, HN, W + pe_dim(d.L_dir) + d.appr_channels, hidden(0, W));
  add_split(L.color[0], pe_c + INST, pe_slices(pe_c, 2));
  add_split(L.color[1], W, hidden(0, W));
  add_split(L.color[2], W, hidden(0, W));
  if (d.has_feature) {
    for (int i = 0; i < 5; ++i) {
      if (i == 0) add_layer(L.feat[i], 128, pe_f, pe_slices(pe_f, 0));
      else if (i == 4) add_layer(L.feat[i], 128, pe_f + 128, cat(pe_slices(pe_f, 0), hidden(pe_f, 128)));
      else add_layer(L.feat[i], 128, 128, hidden(0, 128));
    }
    add_layer(L.feat[5], 16, 128, hidden(0, 128));
  }
  const int pe_d = pe_dim(6), DW = 256;  // DenseWarp: 6 frequencies, two hidden layers of 256
  auto add_split_w = [&](int id, int width, int in_dim, const std::vector<PackSlice>& sls) {
    add_rows(id, width, in_dim, 0, width / 2, sls, chunks[id]);
    add_rows(id, width, in_dim, width / 2, width / 2, sls, chunks_h1[id]);
  };
  if (d.dense) {
    for (int m = 0; m < 2; ++m) {
      add_split_w(L.dense[3 * m + 0], DW, pe_d + TEMB + INST, pe_slices(pe_d, 0));
      add_split_w(L.dense[3 * m + 1], DW, DW, hidden(0, DW));
      add_layer(L.dense[3 * m + 2], 3, DW, hidden(0, DW));
    }
  }
  bp.packed_bytes = off;
  bp.fwd_chunks = chunks;
  bp.fwd_chunks_h1 = chunks_h1;

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
    CondRow& c = F.cond[nc++];  // forward-warp variant of the same layer (mean time code)
    c = F.cond[nc - 2];
    c.frame_off = (int16_t)fo;
    c.code[0] = CODE_T_EMBED_MEAN;
    fo += 64;
  }
  add_cond(L.vis[0], pe_v, INST, CODE_INST_VIS);
  add_cond(L.base[0], pe_b, INST, CODE_INST_BASE);
  add_cond(L.base[d.skip], pe_b, INST, CODE_INST_BASE);
  if (d.appr_channels > 0) add_cond(L.rgb0, W + pe_dim(d.L_dir), d.appr_channels, CODE_APPR);
  add_cond(L.color[0], pe_c, INST, CODE_INST_COLOR);
  F.dense1_partner = -1;
  if (d.dense) {
    add_cond(L.dense[3], pe_d, TEMB, CODE_DENSE_T, pe_d + TEMB, INST, CODE_INST_DENSE_BWD);
    add_cond(L.dense[0], pe_d, TEMB, CODE_DENSE_T, pe_d + TEMB, INST, CODE_INST_DENSE_FWD);
    F.dense1_partner = (int16_t)fo;
    CondRow& c = F.cond[nc++];  // flow: the partner frame's time code (warping.py:459-463 with frame_id_next)
    c = F.cond[nc - 2];
    c.frame_off = (int16_t)fo;
    c.code[0] = CODE_DENSE_T_PARTNER;
    fo += DW;
  }
  F.n_cond = (int16_t)nc;
  auto bones = [&](int per) { int o = fo; fo += B * per; return (int16_t)o; };
  F.binv_t = bones(12); F.se3_bwd = bones(8); F.binv_rest = bones(12); F.se3_fwd = bones(8);
  F.binv_rest_partner = bones(12); F.se3_fwd_partner = bones(8);
  F.n_floats = (int16_t)pad4(fo);

  // ---- constant block: plain bias rows, head weights, rest bone centres, scalars
  ConstLayout& C = P.cl;
  int co = 0;
  for (int i = 0; i < B200R_MAX_LAYERS; ++i) C.plain_off[i] = -1;
  for (int i = 0; i < L.count; ++i)
    if (cond_off[i] < 0) { C.plain_off[i] = (int16_t)co; co += pad16(bp.layer_out[i]); }
  C.sdf_w = (int16_t)co; co += W;
  C.rgb2_w = (int16_t)co; co += 3 * HN;
  C.vis_w = (int16_t)co; co += 64;
  C.dir_w = (int16_t)co; co += (d.L_dir == 0) ? pad4(3 * HN) : 0;
  C.center = (int16_t)co; co += B * 4;
  C.scalars = (int16_t)co; co += kNumScalars;
  C.n_floats = (int16_t)pad4(co);
  for (int i = 0; i < L.count; ++i) {
    P.bias[i].frame = cond_off[i] >= 0;
    P.bias[i].off = (uint16_t)(cond_off[i] >= 0 ? cond_off[i] : C.plain_off[i]);
  }

  // ---- per-tile MMA step list
  int ns = 0;
  // emit one step from 1 or 2 consecutive chunks (same operand kind)
  auto step = [&](const Chunk* c, int n_sub, int a_kind, int a_chunk, int a_chunk2, int a_tmem_col, int d_col, int acc, int wait,
                  int commit) {
    if (ns >= kMaxSteps) { ++ns; return; }  // reported by the caller
    MmaStep& s = P.steps[ns++];
    s = MmaStep{};
    s.w_off = c[0].w_off; s.n = (uint16_t)c[0].n; s.d_col = (uint16_t)d_col; s.a_tmem_col = (uint16_t)a_tmem_col;
    s.a_kind = (uint8_t)a_kind; s.n_sub = (uint8_t)n_sub; s.a_chunk = (uint8_t)a_chunk; s.a_chunk2 = (uint8_t)a_chunk2;
    s.ksteps = (uint8_t)c[0].ksteps; s.ksteps2 = (uint8_t)(split ? c[0].ksteps : (n_sub > 1 ? c[1].ksteps : 0));
    s.accumulate = (uint8_t)acc; s.wait = (uint8_t)wait; s.commit = (uint8_t)commit;
  };
  {
    // ------------------------------------------------------------------ two tiles in flight per CTA.
    // Each tile group owns 128 accumulator columns and 128 columns of 16-bit activations in TMEM; the only
    // shared-memory operands are the embedding chunks CH_PE / CH_EXTRA.  Steps are grouped in blocks (a block ends
    // at the step that commits); the MMA warp issues block b for group 0, then for group 1, then block b+1, ...
    struct Opnd { int kind; int where; };  // kind 0: arena chunk id, kind 1: activation column (16-bit pairs)
    auto ss = [](int chunk) { return Opnd{0, chunk}; };
    auto ts = [](int kc) { return Opnd{1, kc * 32}; };
    auto tsn = [&](int n) { std::vector<Opnd> v; for (int j = 0; j < n; ++j) v.push_back(ts(j)); return v; };
    auto pe5 = [&](int pe_n) { std::vector<Opnd> v{ss(CH_PE)}; if (pe_n > 63) v.push_back(ss(CH_EXTRA)); return v; };
    auto cat5 = [](std::vector<Opnd> a, const std::vector<Opnd>& b) { a.insert(a.end(), b.begin(), b.end()); return a; };
    // one block: D (+)= sum over the chunks of `cs`; consecutive chunks of the same kind share a step when they fit a stage
    bool shape_ok = true;
    auto block = [&](const std::vector<Chunk>& cs, const std::vector<Opnd>& ops, int wait, int commit) {
      const size_t n = cs.size();
      MmaBlock& Bk = P.blocks[P.n_blocks++];
      Bk = MmaBlock{};
      Bk.n16 = (uint8_t)(cs[0].n / 16); Bk.wait = (uint8_t)wait; Bk.commit = (uint8_t)commit;
      int next_col = 0;
      for (size_t c = 0; c < n;) {
        if (split) {  // one chunk per slot: [head tile][tail tile]
          step(&cs[c], 2, ops[c].kind, ops[c].kind == 0 ? ops[c].where : 0, 0, ops[c].kind == 1 ? ops[c].where : 0, 0, c > 0,
               c == 0 ? wait : BAR_NONE, c + 1 == n ? commit : BAR_NONE);
          const int ks = cs[c].ksteps;
          if (ops[c].kind == 0) {  // embedding chunks: the first one or two slots of a block
            shape_ok = shape_ok && c <= 1 && ks <= 7 && (c == 0 || ops[0].kind == 0);
            if (c == 0) { Bk.ss = (uint8_t)(0x80 | ks); Bk.ss_chunks = (uint8_t)ops[c].where; }
            else { Bk.ss |= (uint8_t)(ks << 3); Bk.ss_chunks |= (uint8_t)(ops[c].where << 4); }
          } else {
            shape_ok = shape_ok && ops[c].where == next_col && (ks == 4 || c + 1 == n);
            next_col += 32;
            Bk.ts_slots++;
            Bk.ts_ks2_last = (uint8_t)ks;
          }
          c += 1;
          continue;
        }
        int nsub = 1;
        if (c + 1 < n && ops[c].kind == ops[c + 1].kind && 2 * cs[c].n * 128 <= kWStageBytes &&
            (ops[c].kind == 0 || ops[c + 1].where == ops[c].where + 8 * cs[c].ksteps))
          nsub = 2;
        step(&cs[c], nsub, ops[c].kind, ops[c].kind == 0 ? ops[c].where : 0, (nsub > 1 && ops[c].kind == 0) ? ops[c + 1].where : 0,
             ops[c].kind == 1 ? ops[c].where : 0, 0, c > 0, c == 0 ? wait : BAR_NONE, c + (size_t)nsub == n ? commit : BAR_NONE);
        const int ks = cs[c].ksteps, ks2 = nsub > 1 ? cs[c + 1].ksteps : 0;
        if (ops[c].kind == 0) {  // shared-memory slot: only as the first slot of a block
          shape_ok = shape_ok && c == 0 && ks <= 7 && ks2 <= 7;
          Bk.ss = (uint8_t)(0x80 | ks | (ks2 << 3));
          Bk.ss_chunks = (uint8_t)(ops[c].where | ((nsub > 1 ? ops[c + 1].where : 0) << 4));
        } else {  // activation slot: columns are consumed front to back, 4 k-steps per tile except in the last slot
          shape_ok = shape_ok && ops[c].where == next_col && ks == 4 && (c + (size_t)nsub == n || ks2 == 4);
          next_col += 8 * (ks + ks2);
          Bk.ts_slots++;
          Bk.ts_ks2_last = (uint8_t)ks2;
        }
        c += (size_t)nsub;
      }
    };
    auto seq5 = [&](int id, const std::vector<Opnd>& ops, int wait = BAR_ALL) { block(chunks[id], ops, wait, BAR_ALL); };
    // split layer: N-half 0 then N-half 1 on the same accumulator (drained in between by the epilogue)
    auto pipe5 = [&](int id, const std::vector<Opnd>& ops, int first_wait) {
      block(chunks[id], ops, first_wait, BAR_H0);
      block(chunks_h1[id], ops, BAR_H0, BAR_H1);
    };
    auto delta5 = [&]() {
      seq5(L.delta[0], 3 * B > 64 ? tsn(2) : tsn(1));
      seq5(L.delta[1], tsn(1));
      seq5(L.delta[2], tsn(1));
    };
    auto dense5 = [&](int m) {
      pipe5(L.dense[3 * m + 0], {ss(CH_PE)}, BAR_ALL);
      pipe5(L.dense[3 * m + 1], tsn(4), BAR_H1);
      seq5(L.dense[3 * m + 2], tsn(4), BAR_H1);
    };
    // MODE_POINTS (b200r_points_fwd): only the density + colour chains; MODE_WARP_* (b200r_warp_fwd): one warp only
    const bool points_only = mode == MODE_POINTS, warp_only = mode == MODE_WARP_BWD || mode == MODE_WARP_FWD;
    if (warp_only && B == 0) { bp.err = "warp entry needs a skinned field"; return bp; }
    for (int w = 0; w < 3 && !points_only; ++w) {
      P.st_delta[w] = ns;
      if (warp_only && w != (mode == MODE_WARP_BWD ? 0 : 2)) continue;
      if (d.dense && w > 0) dense5(0);
      if (B > 0) delta5();
      if (d.dense && w == 0) dense5(1);
    }
    if (warp_only) {
      P.n_steps = ns;
      if (ns > kMaxSteps) { bp.err = "too many MMA steps"; return bp; }
      if (!shape_ok) { bp.err = "internal: unsupported block shape"; return bp; }
      bp.ok = true;
      return bp;
    }
    P.st_vis = ns;
    if (!points_only) {
      seq5(L.vis[0], pe5(pe_v));
      seq5(L.vis[1], tsn(1));
    }
    P.st_feat = ns;
    if (d.has_feature && !points_only) {
      seq5(L.feat[0], pe5(pe_f));
      for (int i = 1; i < 4; ++i) seq5(L.feat[i], tsn(2));
      seq5(L.feat[4], cat5(pe5(pe_f), tsn(2)));
      seq5(L.feat[5], tsn(2));
    }
    P.st_base = ns;
    for (int i = 0; i <= d.D; ++i) {
      if (i == 0) pipe5(L.base[i], pe5(pe_b), BAR_ALL);
      else if (i == d.skip) pipe5(L.base[i], cat5(pe5(pe_b), tsn(KC)), BAR_H1);
      else pipe5(L.base[i], tsn(KC), BAR_H1);
    }
    P.st_color = ns;
    pipe5(L.color[0], pe5(pe_c), BAR_H1);
    pipe5(L.color[1], tsn(KC), BAR_H1);
    pipe5(L.color[2], tsn(KC), BAR_H1);
    P.st_rgb = ns;
    seq5(L.rgb0, tsn(KC), BAR_H1);
    P.n_steps = ns;
    if (ns > kMaxSteps) { bp.err = "too many MMA steps"; return bp; }
    if (!shape_ok) { bp.err = "internal: unsupported block shape"; return bp; }
    bp.ok = true;
    return bp;
  }
}

// ------------------------------------------------------------------------------------------------ training tape
// A training-mode forward (b200r_field_fwd with a tape) records, per 128-sample tile, every MMA operand it produced
// as [128 rows x 64] 16-bit chunks in the K-major SWIZZLE_128B image (16 KB each; row r at r * 128 B, 16-B group g at
// ((g ^ (r & 7)) << 4)), stored per tile as [64-row half][chunk][64 rows x 128 B] (ptx.cuh tape_row_off: the chunks of one
// half are adjacent), plus one word of ReLU sign bits per (row, 32 columns).  The backward (b200r_field_bwd) adds
// its masked gradients in the same format; the weight-gradient kernel then reads both straight into shared memory
// with bulk copies and multiplies them as MN-major UMMA operands (reduction over the tile's rows).
constexpr int kChunkBytes = kAChunkBytes;
constexpr int kMaskWords = 8;  // sign words per (row, slot): up to 256 columns
struct TapeLayout {
  int16_t n_a, n_g, n_mask;  // chunks per tile written by the forward / by the backward; mask slots per row
  // ---- forward-written operand chunks (first chunk id of each group, -1 = absent)
  int16_t a_xb[3], a_h1[3], a_h2[3], a_z[3];  // skinning warps w = 0 (backward), 1 (flow), 2 (cycle): bone coords, delta MLP hidden, raw output
  int16_t a_pe, a_extra;                       // Fourier embedding of the canonical point: columns 0..62, 63..74
  int16_t a_vis[2], a_feat[5], a_base[10], a_col[2], a_f2, a_rgb0;  // hidden activations; a_base[D] = base features; a_f2 = input of rgb.0
  int16_t a_dpe[3], a_dh1[3], a_dh2[3];        // dense warp stages (backward map, forward map partner, forward map own)
  int16_t a_dir;                               // raw view direction (3 columns), L_dir == 0
  // ---- sign-bit slots (bit = 1: pre-activation <= 0)
  int16_t m_h1[3], m_h2[3], m_z[3], m_vis[2], m_feat[5], m_base[10], m_col[3], m_rgb0, m_dh1[3], m_dh2[3];
  // ---- backward-written gradient chunks
  int16_t g_z1[3], g_z2[3], g_z[3], g_xbw[3], g_xg[3];
  int16_t g_vis[2], g_feat[6], g_base[10], g_col[3], g_rgb0, g_head;
  int16_t g_d1[3], g_d2[3], g_d3[3];
};
// columns of the head chunk (g_head): gradients of the fp32 heads' pre-activations
enum : int { GH_SDF = 0, GH_RGB = 1, GH_VIS = 4, kHeadCols = 16 };

inline TapeLayout tape_layout(const b200r_field_desc& d) {
  TapeLayout T;
  int16_t* p = reinterpret_cast<int16_t*>(&T);
  for (size_t i = 0; i < sizeof(T) / sizeof(int16_t); ++i) p[i] = -1;
  const int B = d.n_bones, KC = d.W / 64;
  int a = 0, g = 0, m = 0;
  auto A = [&](int n) { int o = a; a += n; return (int16_t)o; };
  auto G = [&](int n) { int o = g; g += n; return (int16_t)o; };
  auto Mk = [&]() { return (int16_t)(m++); };
  for (int w = 0; w < 3 && B > 0; ++w) {
    T.a_xb[w] = A((3 * B + 63) / 64); T.a_h1[w] = A(1); T.a_h2[w] = A(1); T.a_z[w] = A(1);
    T.m_h1[w] = Mk(); T.m_h2[w] = Mk(); T.m_z[w] = Mk();
    T.g_z1[w] = G(1); T.g_z2[w] = G(1); T.g_z[w] = G(1); T.g_xbw[w] = G(2); T.g_xg[w] = G(1);
    if (d.dense) {
      T.a_dpe[w] = A(1); T.a_dh1[w] = A(4); T.a_dh2[w] = A(4);
      T.m_dh1[w] = Mk(); T.m_dh2[w] = Mk();
      T.g_d1[w] = G(4); T.g_d2[w] = G(4); T.g_d3[w] = G(1);
    }
  }
  T.a_pe = A(1); T.a_extra = A(1);
  if (d.L_dir == 0) T.a_dir = A(1);
  for (int i = 0; i < 2; ++i) { T.a_vis[i] = A(1); T.m_vis[i] = Mk(); T.g_vis[i] = G(1); }
  if (d.has_feature) {
    for (int i = 0; i < 5; ++i) { T.a_feat[i] = A(2); T.m_feat[i] = Mk(); T.g_feat[i] = G(2); }
    T.g_feat[5] = G(1);
  }
  for (int i = 0; i <= d.D; ++i) { T.a_base[i] = A(KC); T.m_base[i] = Mk(); T.g_base[i] = G(KC); }
  for (int i = 0; i < 2; ++i) { T.a_col[i] = A(KC); T.m_col[i] = Mk(); T.g_col[i] = G(KC); }
  T.m_col[2] = Mk(); T.g_col[2] = G(KC);
  T.a_f2 = A(KC);
  T.a_rgb0 = A(KC / 2); T.m_rgb0 = Mk(); T.g_rgb0 = G(KC / 2);
  T.g_head = G(1);
  T.n_a = (int16_t)a; T.n_g = (int16_t)g; T.n_mask = (int16_t)m;
  return T;
}
inline size_t tape_a_bytes(const TapeLayout& T, int n_tiles) { return (size_t)n_tiles * T.n_a * kChunkBytes; }
inline size_t tape_g_bytes(const TapeLayout& T, int n_tiles) { return (size_t)n_tiles * T.n_g * kChunkBytes; }
inline size_t tape_mask_bytes(const TapeLayout& T, int n_tiles) { return (size_t)n_tiles * kTileRows * T.n_mask * kMaskWords * 4; }

// ------------------------------------------------------------------------------------------------ backward (dgrad) program
// The backward kernel walks the layers in reverse: G_{l-1} = (G_l W_l) * relu'.  Every GEMM reads its A operand (the
// 16-bit gradient rows) from the group's activation columns in TMEM and its B operand from a TRANSPOSED packed tile
// (rows = in-features of the layer, K = out-features), streamed through the same ring by the same producer / issuer
// code as the forward.  Block order = the order of csrc/field_bwd.cu's phases:
//   rgb.0, colorfield (final, 2, 1), basefield (final, D..1; the skip layer first returns its embedding columns),
//   feature field, visibility MLP, then per skinning warp w = 2, 1, 0: [dense map], delta MLP (final, 2, 1).
// `density_only`: the step list holds the density chain alone (basefield final .. linear_1, the reverse chain of the eikonal
// term, csrc/field_bwd.cu mode 1) - the tile offsets are those of the full program, so both read the same W^T buffer.
// `warp_w` in {0, 1, 2}: the step list holds that skinning warp's blocks alone (delta MLP in reverse and its dense map) - the
// backward of one warp of given points (csrc/field_bwd.cu WARPONLY).
// density_only together with warp_w: both block groups (density chain, then that warp) - the normals entry (csrc/field_bwd.cu NORMALS).
inline BuiltProgram build_bwd_program(const b200r_field_desc& d, bool density_only = false, int warp_w = -1) {
  BuiltProgram bp = build_program(d, MODE_FIELD);  // same constant / frame block layouts, same checks
  if (!bp.ok) return bp;
  bp.ok = false;
  bp.slices.clear();
  Program& P = bp.prog;
  P.n_steps = 0;
  P.n_blocks = 0;
  const LayerIds L = layer_ids(d);
  const int B = d.n_bones, W = d.W, HN = W / 2;
  const int pe_b = pe_dim(d.L_xyz), pe_c = pe_dim(d.L_xyz + 2), pe_v = pe_dim(10), pe_f = pe_dim(6), pe_d = pe_dim(6);
  uint32_t off = 0;
  int ns = 0;
  bool ok = true;
  struct Chunk { uint32_t w_off; int n; int ksteps; };
  // tiles of W_layer^T for in-features [r0, r0 + rows): one [pad16(rows) x 64] tile per 64 out-features
  auto chunks_of = [&](int layer, int r0, int rows, int win) {
    std::vector<Chunk> v;
    const int n_out = bp.layer_out[layer];
    for (int c0 = 0; c0 < n_out; c0 += 64) {
      PackSlice s{};
      s.layer = layer; s.transpose = 1;
      s.n = rows; s.n_pad = pad16(rows); s.in_dim = bp.layer_in[layer];
      s.col0 = c0; s.ncols = n_out - c0 < 64 ? n_out - c0 : 64;
      s.row0 = r0; s.pe_window = win; s.pe_col0 = r0; s.dst_off = off;
      v.push_back({off, s.n_pad, (s.ncols + 15) / 16});
      off += (uint32_t)s.n_pad * 128u;
      bp.slices.push_back(s);
    }
    return v;
  };
  bool emit = !density_only && warp_w < 0;  // tiles are laid out for every block; steps only for the emitted ones
  auto block = [&](const std::vector<Chunk>& cs, int wait, int commit) {
    if (!emit) return;
    if (P.n_blocks >= kMaxSteps) { ok = false; return; }
    MmaBlock& Bk = P.blocks[P.n_blocks++];
    Bk = MmaBlock{};
    Bk.n16 = (uint8_t)(cs[0].n / 16); Bk.wait = (uint8_t)wait; Bk.commit = (uint8_t)commit;
    for (size_t c = 0; c < cs.size();) {
      const int nsub = (c + 1 < cs.size() && 2 * cs[c].n * 128 <= kWStageBytes && cs[c].ksteps == 4) ? 2 : 1;
      if (ns >= kMaxSteps) { ok = false; return; }
      MmaStep& S = P.steps[ns++];
      S = MmaStep{};
      S.w_off = cs[c].w_off; S.n = (uint16_t)cs[c].n; S.n_sub = (uint8_t)nsub; S.a_kind = 1;
      S.ksteps = (uint8_t)cs[c].ksteps; S.ksteps2 = (uint8_t)(nsub > 1 ? cs[c + 1].ksteps : 0);
      S.accumulate = c > 0; S.wait = (uint8_t)(c == 0 ? wait : BAR_NONE);
      S.commit = (uint8_t)(c + (size_t)nsub == cs.size() ? commit : BAR_NONE);
      // slots of activation operands: 4 k-steps per tile except in the last slot (field kernels' issuer contract)
      ok = ok && (cs[c].ksteps == 4 || c + (size_t)nsub == cs.size());
      Bk.ts_slots++;
      Bk.ts_ks2_last = (uint8_t)(nsub > 1 ? cs[c + 1].ksteps : 0);
      if (nsub == 1 && cs[c].ksteps != 4) {  // a lone short tile: encode its k-steps as "first tile" of the last slot
        Bk.pad_ = (uint8_t)cs[c].ksteps;
      }
      c += (size_t)nsub;
    }
  };
  auto seq = [&](int layer, int r0, int rows, int win = 0, int wait = BAR_ALL) { block(chunks_of(layer, r0, rows, win), wait, BAR_ALL); };
  auto pipe = [&](int layer, int r0, int width, int first_wait) {
    const auto h0 = chunks_of(layer, r0, width / 2, 0), h1 = chunks_of(layer, r0 + width / 2, width / 2, 0);
    block(h0, first_wait, BAR_H0);
    block(h1, BAR_H0, BAR_H1);
  };
  // ---- rgb.0 and the colour chain
  pipe(L.rgb0, 0, W, BAR_ALL);
  pipe(L.color[2], 0, W, BAR_H1);
  pipe(L.color[1], 0, W, BAR_H1);
  seq(L.color[0], 0, pe_c, 2, BAR_H1);
  // ---- density chain
  if (density_only) emit = true;
  pipe(L.base[d.D], 0, W, BAR_ALL);
  for (int i = d.D - 1; i >= 1; --i) {
    if (i == d.skip) {
      seq(L.base[i], 0, pe_b, 1, BAR_H1);
      pipe(L.base[i], pe_b + 32, W, BAR_ALL);
    } else {
      pipe(L.base[i], 0, W, BAR_H1);
    }
  }
  seq(L.base[0], 0, pe_b, 1, BAR_H1);
  if (density_only) emit = false;
  // ---- feature field
  if (d.has_feature) {
    seq(L.feat[5], 0, 128);
    seq(L.feat[4], 0, pe_f);
    seq(L.feat[4], pe_f, 128);
    for (int i = 3; i >= 1; --i) seq(L.feat[i], 0, 128);
    seq(L.feat[0], 0, pe_f);
  }
  // ---- visibility MLP
  seq(L.vis[1], 0, 64);
  seq(L.vis[0], 0, pe_v);
  // ---- skinning warps, last first: cycle (w = 2), flow (w = 1), backward (w = 0)
  for (int w = 2; w >= 0 && B > 0; --w) {
    auto dense = [&](int m) {  // DenseWarp map m (0 forward_map, 1 backward_map): linear_2, linear_1 (the 3-wide head is SIMT)
      pipe(L.dense[3 * m + 1], 0, 256, BAR_ALL);
      seq(L.dense[3 * m + 0], 0, pe_d, 0, BAR_H1);
    };
    if (warp_w >= 0) emit = w == warp_w;
    if (d.dense && w == 0) dense(1);
    seq(L.delta[2], 0, 64);
    seq(L.delta[1], 0, 64);
    seq(L.delta[0], 0, 3 * B);
    if (d.dense && w > 0) dense(0);
  }
  P.n_steps = ns;
  bp.packed_bytes = off;
  if (!ok) { bp.err = "internal: backward program does not fit"; return bp; }
  bp.ok = true;
  return bp;
}

// ------------------------------------------------------------------------------------------------ eikonal chains
// Forward chains of the eikonal term's backward (csrc/field_bwd.cu modes 2 / 3): v_l = relu'_l * (W_l v_{l-1}) over the
// basefield with the FORWARD operand tiles (heads only in the split mode), every operand in TMEM, one tile per ring slot.
//   chain A: linear_1 (embedding chunk), linear_2 .. final (the skip layer takes its hidden chunks only);
//   chain B: the skip layer's embedding chunk, then the layers after it.
enum : int { EIK_REVERSE = 1, EIK_CHAIN_A = 2, EIK_CHAIN_B = 3 };
inline BuiltProgram build_eik_chain_program(const b200r_field_desc& d, int which) {
  BuiltProgram bp = build_program(d, MODE_FIELD);
  if (!bp.ok) return bp;
  bp.ok = false;
  Program& P = bp.prog;
  P.n_steps = 0;
  P.n_blocks = 0;
  const LayerIds L = layer_ids(d);
  int ns = 0;
  bool ok = true;
  auto block = [&](const std::vector<ChunkRef>& cs, size_t c0, size_t c1, int wait, int commit) {
    if (c1 > cs.size() || c0 >= c1) { ok = false; return; }
    for (size_t c = c0; c < c1; ++c) {
      if (ns >= kMaxSteps) { ok = false; return; }
      MmaStep& S = P.steps[ns++];
      S = MmaStep{};
      S.w_off = cs[c].w_off; S.n = (uint16_t)cs[c].n; S.n_sub = 1; S.a_kind = 1;
      S.ksteps = (uint8_t)cs[c].ksteps;
      S.accumulate = c > c0; S.wait = (uint8_t)(c == c0 ? wait : BAR_NONE); S.commit = (uint8_t)(c + 1 == c1 ? commit : BAR_NONE);
      ok = ok && (cs[c].ksteps == 4 || c + 1 == c1);  // operand columns advance by 64 halves per slot
    }
  };
  auto pipe = [&](int layer, size_t c0, size_t c1, int first_wait) {
    block(bp.fwd_chunks[layer], c0, c1, first_wait, BAR_H0);
    block(bp.fwd_chunks_h1[layer], c0, c1, BAR_H0, BAR_H1);
  };
  const size_t n_pe_chunks = bp.fwd_chunks[L.base[0]].size();  // embedding chunks of the basefield input
  if (n_pe_chunks != 1) { bp.err = "eikonal chains need a one-chunk position embedding"; return bp; }
  if (which == EIK_CHAIN_A) {
    for (int i = 0; i <= d.D; ++i) {
      const size_t n = bp.fwd_chunks[L.base[i]].size();
      if (i == 0) pipe(L.base[i], 0, 1, BAR_ALL);
      else if (i == d.skip) pipe(L.base[i], n_pe_chunks, n, BAR_H1);
      else pipe(L.base[i], 0, n, BAR_H1);
    }
  } else if (which == EIK_CHAIN_B) {
    pipe(L.base[d.skip], 0, 1, BAR_ALL);
    for (int i = d.skip + 1; i <= d.D; ++i) pipe(L.base[i], 0, bp.fwd_chunks[L.base[i]].size(), BAR_H1);
  } else {
    bp.err = "internal: unknown eikonal chain";
    return bp;
  }
  P.n_steps = ns;
  if (!ok) { bp.err = "internal: eikonal chain program does not fit"; return bp; }
  bp.ok = true;
  return bp;
}
// chunk ids inside the two eikonal tapes (per 128-point tile, same [half][chunk] image as the training tape)
struct EikLayout {
  int n_a, n_v;  // chunks per tile of the reverse-chain tape / of the forward-chain tape
  int a_base[10], a_head;            // reverse chain: a_i of basefield layer i (KC chunks each), head chunk
  int v0, vA[10], vB[10];            // forward chains: v_0 (1 chunk), chain A / chain B outputs per layer (-1 = absent)
};
inline EikLayout eik_layout(const b200r_field_desc& d) {
  EikLayout E;
  const int KC = d.W / 64;
  int a = 0, v = 0;
  for (int i = 0; i < 10; ++i) { E.a_base[i] = -1; E.vA[i] = -1; E.vB[i] = -1; }
  for (int i = 0; i <= d.D; ++i) { E.a_base[i] = a; a += KC; }
  E.a_head = a++;
  E.v0 = v++;
  for (int i = 0; i <= d.D; ++i) { E.vA[i] = v; v += KC; }
  for (int i = d.skip; i <= d.D; ++i) { E.vB[i] = v; v += KC; }
  E.n_a = a; E.n_v = v;
  return E;
}

inline size_t workspace_floats(const Program& P, int M) { return (size_t)P.cl.n_floats + (size_t)M * P.fl.n_floats; }
// the field kernel keeps the 256 packed base features of every row of both tiles in flight in a per-CTA scratch (L2 resident)
constexpr int kMaxCtas = 160;
constexpr size_t kScratchPerCta = 2 * (size_t)kTileRows * 512;
inline size_t scratch_offset_bytes(const Program& P, int M) { return (workspace_floats(P, M) * sizeof(float) + 255) / 256 * 256; }
inline size_t workspace_bytes_total(const Program& P, int M) { return scratch_offset_bytes(P, M) + kMaxCtas * kScratchPerCta; }

}  // namespace b200r
