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
