"""CPU: the C-ABI library loads and exports every symbol include/b200r.h declares; host-side
program construction is consistent.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "b200r.h")).read()
    return sorted(set(re.findall(r"\b(b200r_[a-z_]+)\s*\(", src)))


def test_library_exports_declared_symbols():
    from lab4d_b200 import _lib, build

    build.build()
    lib = _lib.load()
    names = _declared()
    assert set(names) == set(_lib.EXPORTS)
    for n in names:
        assert hasattr(lib, n), n


def test_header_is_plain_c_and_struct_sizes_match_ctypes(tmp_path):
    """include/b200r.h compiles as C99 (no C++ or torch types in the ABI), a C program links against the library, and
    the ctypes mirrors in lab4d_b200/_lib.py have the same sizes as the C structs."""
    import shutil
    import subprocess

    from lab4d_b200 import _lib, build

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    build.build()
    src = tmp_path / "abi.c"
    src.write_text(
        '#include <stdio.h>\n#include "b200r.h"\n'
        "int main(void) {\n"
        "  b200r_field_desc d = {0};\n"
        '  printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(b200r_eik_batch), sizeof(b200r_match_args), sizeof(b200r_match_bwd_args), sizeof(b200r_loss_args), sizeof(b200r_loss_bwd_args), sizeof(b200r_compose_bwd_args));\n'
        '  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(b200r_field_desc), sizeof(b200r_field_params), sizeof(b200r_frame_tables),\n'
        "         sizeof(b200r_ray_batch), sizeof(b200r_field_outputs), sizeof(b200r_composite_args), sizeof(b200r_composite_bwd_args),\n"
        "         sizeof(b200r_compose_args), sizeof(b200r_point_batch), sizeof(b200r_importance_args), sizeof(b200r_field_grads), sizeof(b200r_tape),\n"
        "         sizeof(b200r_block_layout), sizeof(b200r_param_grads), sizeof(b200r_frame_grads), (size_t)b200r_layer_count(&d));\n"
        "  return 0;\n}\n")
    exe = tmp_path / "abi"
    libdir = os.path.join(ROOT, "lab4d_b200")
    cuda_lib = "/usr/local/cuda/lib64"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-l:libb200render.so", f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{cuda_lib}", f"-Wl,-rpath-link,{cuda_lib}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    new = [int(x) for x in out[:6]]
    assert new == [C.sizeof(m) for m in (_lib.EikBatch, _lib.MatchArgs, _lib.MatchBwdArgs, _lib.LossArgs, _lib.LossBwdArgs, _lib.ComposeBwdArgs)]
    out = out[6:]
    sizes = [int(x) for x in out[:15]]
    mirrors = [_lib.FieldDesc, _lib.FieldParams, _lib.FrameTables, _lib.RayBatch, _lib.FieldOutputs, _lib.CompositeArgs,
               _lib.CompositeBwdArgs, _lib.ComposeArgs, _lib.PointBatch, _lib.ImportanceArgs, _lib.FieldGrads, _lib.Tape,
               _lib.BlockLayout, _lib.ParamGrads, _lib.FrameGrads]
    assert sizes == [C.sizeof(m) for m in mirrors]


def test_layer_counts_and_packed_sizes():
    from lab4d_b200 import _lib, spec

    lib = _lib.load()
    for cfg, nl in ((spec.FG_BOB, 3 + 2 + 9 + 1 + 3 + 6), (spec.FG_RIGID, 2 + 9 + 1 + 3 + 6), (spec.BG, 2 + 6 + 1 + 3),
                    (spec.FG_COMP_QUAD, 3 + 2 + 9 + 1 + 3 + 6 + 6)):
        d = _lib.FieldDesc(category=0 if cfg.category == "fg" else 1, D=cfg.D, W=cfg.W, L_xyz=cfg.L_xyz, L_dir=cfg.L_dir,
                           appr_channels=cfg.appr_channels, skip=cfg.skip, n_bones=cfg.B if cfg.motion != "rigid" else 0,
                           has_feature=int(cfg.has_feature), operand_dtype=0, dense=int(cfg.dense))
        assert lib.b200r_layer_count(C.byref(d)) == nl
        nbytes = lib.b200r_packed_bytes(C.byref(d))
        assert nbytes > 0 and nbytes % 2048 == 0
        # every parameter the packer reads exists in the parameter spec
        shapes = spec.field_param_shapes(cfg)
        assert all(k in shapes for k in ("sdf.weight", "rgb.2.weight", "vis_mlp.basefield.linear_final.weight"))
    # a dense post-warp needs a skinned field
    bad = _lib.FieldDesc(category=0, D=8, W=256, L_xyz=10, L_dir=-1, appr_channels=32, skip=4, n_bones=0, has_feature=1, operand_dtype=0, dense=1)
    assert lib.b200r_packed_bytes(C.byref(bad)) == 0
    bad = _lib.FieldDesc(category=0, D=8, W=200, L_xyz=10, L_dir=-1, appr_channels=32, skip=4, n_bones=25, has_feature=1, operand_dtype=0)
    assert lib.b200r_packed_bytes(C.byref(bad)) == 0


def test_renderer_refuses_cpu():
    from lab4d_b200 import spec
    from lab4d_b200.render import FieldRenderer

    with pytest.raises(RuntimeError):
        FieldRenderer(spec.FG_BOB, device="cpu")


def test_training_layouts_are_consistent():
    """Host-side training metadata for every field type: tape sizes scale with the tile count, the block layout's offsets are
    inside their blocks and disjoint, the flat gradient buffer of the renderer covers every hot-path parameter once."""
    from lab4d_b200 import _lib, spec

    lib = _lib.load()
    for cfg in (spec.BG, spec.FG_RIGID, spec.FG_BOB, spec.FG_SKEL_HUMAN, spec.FG_COMP_QUAD):
        d = _lib.FieldDesc(category=0 if cfg.category == "fg" else 1, D=cfg.D, W=cfg.W, L_xyz=cfg.L_xyz, L_dir=cfg.L_dir,
                           appr_channels=cfg.appr_channels, skip=cfg.skip, n_bones=cfg.B if cfg.motion != "rigid" else 0,
                           has_feature=int(cfg.has_feature), operand_dtype=0, dense=int(cfg.dense))
        a, g, m = C.c_size_t(), C.c_size_t(), C.c_size_t()
        assert lib.b200r_tape_sizes(C.byref(d), 4, 16, 64, C.byref(a), C.byref(g), C.byref(m)) == 0
        a2, g2, m2 = C.c_size_t(), C.c_size_t(), C.c_size_t()
        assert lib.b200r_tape_sizes(C.byref(d), 8, 16, 64, C.byref(a2), C.byref(g2), C.byref(m2)) == 0
        assert a2.value == 2 * a.value and m2.value == 2 * m.value and a.value % 16384 == 0 and g2.value > g.value
        assert lib.b200r_packed_t_bytes(C.byref(d)) > 0
        # eikonal tapes: (D+1) KC + 1 chunks of reverse chain, 1 + (D+1) KC + (D+1-skip) KC chunks of forward chains per 128-point tile
        ea, ev = C.c_size_t(), C.c_size_t()
        assert lib.b200r_eikonal_sizes(C.byref(d), 8, 64, C.byref(ea), C.byref(ev)) == 0
        KC, tiles = cfg.W // 64, (8 * 64) // 128 + 160
        assert ea.value == tiles * ((cfg.D + 1) * KC + 1) * 16384
        assert ev.value == tiles * (1 + (cfg.D + 1) * KC + (cfg.D + 1 - cfg.skip) * KC) * 16384
        lay = _lib.BlockLayout()
        assert lib.b200r_get_block_layout(C.byref(d), C.byref(lay)) == 0
        spans = []
        for i in range(lay.n_cond):
            c = lay.cond[i]
            assert 0 <= c.frame_off and c.frame_off + c.n <= lay.frame_floats
            spans.append((c.frame_off, c.frame_off + c.n))
        spans += [(lay.f_cam, lay.f_cam + 24), (lay.f_cam_partner, lay.f_cam_partner + 24)]
        if cfg.motion != "rigid":
            B = cfg.B
            for off, per in ((lay.f_binv_t, 12), (lay.f_se3_bwd, 8), (lay.f_binv_rest, 12), (lay.f_se3_fwd, 8),
                             (lay.f_binv_rest_partner, 12), (lay.f_se3_fwd_partner, 8)):
                assert off >= 0
                spans.append((off, off + B * per))
        spans.sort()
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 <= b0, (cfg, spans)
        assert spans[-1][1] <= lay.frame_floats
        assert lay.c_scalars + 8 <= lay.const_floats and lay.c_sdf_w >= 0 and lay.c_rgb2_w >= 0


def test_per_tile_programs_of_every_entry_build():
    """Host logic of the program builders (csrc/program.h) for every field type and every entry: the step lists exist, the
    sub-programs are consistent with the full backward program, the eikonal chains have the expected slot counts."""
    from lab4d_b200 import _lib, spec

    lib = _lib.load()
    comphuman = spec.FieldConfig(motion="skel", B=18, symm_idx=spec.HUMAN_SYMM, dense=True)
    for cfg in (spec.BG, spec.FG_RIGID, spec.FG_BOB, spec.FG_SKEL_HUMAN, spec.FG_COMP_QUAD, comphuman):
        for dtype in (0, 1, 2):
            d = _lib.FieldDesc(category=0 if cfg.category == "fg" else 1, D=cfg.D, W=cfg.W, L_xyz=cfg.L_xyz, L_dir=cfg.L_dir,
                               appr_channels=cfg.appr_channels, skip=cfg.skip, n_bones=cfg.B if cfg.motion != "rigid" else 0,
                               has_feature=int(cfg.has_feature), operand_dtype=dtype, dense=int(cfg.dense))
            n = {k: lib.b200r_program_steps(C.byref(d), k) for k in range(8)}
            skinned = cfg.motion != "rigid"
            assert n[0] > 0 and n[1] > 0 and n[2] > 0, (cfg, dtype, n)
            assert 0 < n[2] < n[1]
            if skinned:
                assert n[3] > 0 and n[7] > 0 and n[4] == n[2] + n[7]
                assert n[1] >= n[2] + n[3] + n[7]
            else:
                assert n[3] == 0 and n[7] == 0 and n[4] == n[2]
            KC = cfg.W // 64
            assert n[5] == 2 * (1 + KC * cfg.D), (cfg, n[5])            # linear_1 (one embedding chunk), D layers of KC hidden chunks, two N-halves
            assert n[6] == 2 * (1 + KC * (cfg.D - cfg.skip)), (cfg, n[6])  # the skip layer's embedding chunk, the layers after it
    assert lib.b200r_program_steps(C.byref(d), 99) < 0
