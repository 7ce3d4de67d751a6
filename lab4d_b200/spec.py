"""Field configurations and hot-path parameter layout.

Names are the reference's state_dict keys (probe dump of
MultiFields(...).field_params[cat].state_dict()), so a reference checkpoint / live module can be
handed to the renderer unchanged.  Field shapes follow MultiFields.define_field
(lab4d/nnutils/multifields.py:60-100): fg = Deformable(D=8, W=256, L_xyz=10, L_dir=-1, appr=32),
bg = NeRF(D=5, W=128, L_xyz=6, L_dir=0, appr=0); VisField D=2 W=64 L=10
(nnutils/visibility.py:25-50); feature field D=5 W=128 L=6 -> 16 (nnutils/feature.py:77-85);
skinning delta field D=2 W=64 (nnutils/skinning.py:44-87).
"""
from dataclasses import dataclass, field as _f
from typing import Optional, Tuple

INST_CH = 32  # CondMLP inst_channels (nnutils/base.py:96-121)
T_EMBED_CH = 128  # TimeEmbedding out_channels (nnutils/embedding.py:146)


def pe_dim(L, C=3):
    return 0 if L == -1 else C * (2 * L + 1)


@dataclass(frozen=True)
class FieldConfig:
    category: str = "fg"  # "fg" | "bg"
    D: int = 8  # basefield depth
    W: int = 256  # basefield width
    L_xyz: int = 10
    L_dir: int = -1
    appr_channels: int = 32
    skip: int = 4
    motion: str = "bob"  # "rigid" | "bob" | "skel"
    B: int = 25  # bones
    has_feature: bool = True
    symm_idx: Optional[Tuple[int, ...]] = None
    dense: bool = False  # ComposedWarp: DenseWarp(D=2, W=256) soft deformation around the skinning warp

    def as_oracle_cfg(self):
        d = dict(category=self.category, D=self.D, W=self.W, L_xyz=self.L_xyz, L_dir=self.L_dir,
                 appr_channels=self.appr_channels, motion=self.motion, B=self.B, has_feature=self.has_feature)
        if self.symm_idx is not None:
            d["symm_idx"] = list(self.symm_idx)
        d["dense"] = self.dense
        return d


FG_BOB = FieldConfig()
FG_RIGID = FieldConfig(motion="rigid", B=0)
HUMAN_SYMM = (0, 1, 2, 3, 8, 9, 10, 11, 4, 5, 6, 7, 15, 16, 17, 12, 13, 14)
FG_SKEL_HUMAN = FieldConfig(motion="skel", B=18, symm_idx=HUMAN_SYMM)  # configs[2]: skel-human
QUAD_SYMM = (0, 1, 2, 3, 8, 9, 10, 11, 4, 5, 6, 7, 12, 13, 14, 15, 16, 21, 22, 23, 24, 17, 18, 19, 20)
FG_COMP_QUAD = FieldConfig(motion="skel", B=25, symm_idx=QUAD_SYMM, dense=True)  # comp_skel-quad_dense
BG = FieldConfig(category="bg", D=5, W=128, L_xyz=6, L_dir=0, appr_channels=0, motion="rigid", B=0, has_feature=False)


def _mlp_shapes(prefix, d_in, W, D, d_out, skip, final_seq):
    s = {}
    for i in range(D):
        k = d_in if i == 0 else (W + d_in if i == skip else W)
        s[f"{prefix}linear_{i+1}.0.weight"] = (W, k)
        s[f"{prefix}linear_{i+1}.0.bias"] = (W,)
    fin = "linear_final.0." if final_seq else "linear_final."
    s[prefix + fin + "weight"] = (d_out, W)
    s[prefix + fin + "bias"] = (d_out,)
    return s


def field_param_shapes(cfg: FieldConfig):
    """name -> shape of every parameter the per-sample hot path reads."""
    W = cfg.W
    s = {"logibeta": (1,), "logscale": (1,)}
    s.update(_mlp_shapes("basefield.", pe_dim(cfg.L_xyz) + INST_CH, W, cfg.D, W, cfg.skip, True))
    s.update(_mlp_shapes("colorfield.", pe_dim(cfg.L_xyz + 2) + INST_CH, W, 2, W, cfg.skip, True))
    s["sdf.weight"], s["sdf.bias"] = (1, W), (1,)
    s["rgb.0.weight"] = (W // 2, W + pe_dim(cfg.L_dir) + cfg.appr_channels)
    s["rgb.0.bias"] = (W // 2,)
    s["rgb.2.weight"], s["rgb.2.bias"] = (3, W // 2), (3,)
    s.update(_mlp_shapes("vis_mlp.basefield.", pe_dim(10) + INST_CH, 64, 2, 1, 4, False))
    if cfg.has_feature:
        s.update(_mlp_shapes("feature_field.", pe_dim(6), 128, 5, 16, 4, False))
    if cfg.motion != "rigid":
        s["warp.logibeta"] = (1,)
        s["warp.skinning_model.log_gauss"] = (cfg.B, 3)
        s.update(_mlp_shapes("warp.skinning_model.delta_field.", 3 * cfg.B + T_EMBED_CH + INST_CH, 64, 2, cfg.B, 4, False))
        if cfg.dense:
            for m in ("forward_map", "backward_map"):
                s.update(_mlp_shapes(f"warp.post_warp.{m}.", pe_dim(6) + T_EMBED_CH + INST_CH, 256, 2, 3, 4, False))
    return s
