"""Shared helpers for the parity tests: golden-fixture loading and synthetic parameters."""
import glob
import os

import numpy as np
import torch

import synth
from lab4d_b200 import spec

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CFG_OF = {"bg_rigid": spec.BG, "fg_rigid": spec.FG_RIGID, "fg_bob": spec.FG_BOB, "fg_compquad": spec.FG_COMP_QUAD,
          "fg_skelhuman": spec.FG_SKEL_HUMAN}


def golden_files(prefix=""):
    return sorted(glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def cfg_for(path, cat=None):
    base = os.path.basename(path)
    if base.startswith("comp"):
        return spec.FG_BOB if cat == "fg" else spec.BG
    for k, v in CFG_OF.items():
        if base.startswith(k):
            return v
    raise KeyError(base)


def load_golden(path):
    z = np.load(path)
    return {k: z[k] for k in z.files}


def sub(pack, prefix, dtype=None, device=None):
    out = {}
    for k, v in pack.items():
        if k.startswith(prefix):
            t = torch.from_numpy(np.asarray(v))
            if dtype is not None and t.dtype.is_floating_point:
                t = t.to(dtype)
            if device is not None:
                t = t.to(device)
            out[k[len(prefix):]] = t
    return out


def synth_params(cfg, seed=0, dtype=torch.float32, device=None):
    st = synth.synth_state(spec.field_param_shapes(cfg), seed, cfg.category)
    return {k: torch.from_numpy(v).to(dtype).to(device or "cpu") for k, v in st.items()}


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))
