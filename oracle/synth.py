"""Deterministic synthetic parameters and ray batches shared by the golden-vector generator,
the tests and bench.py.  TEST/BENCH INFRASTRUCTURE (numpy only, no reference import).

There is no checkpoint offline, and PyTorch-default-initialised Lab4D fields give a spatially
constant SDF, so every harness in this repo uses the same "trained-like"
synthetic rule instead: each tensor is drawn from a numpy RandomState seeded by crc32(name), so
any consumer can regenerate exactly the tensors it needs from (name, shape, seed) alone and the
committed fixtures only have to hold inputs/outputs, not 8.6 MB of weights.
"""
import zlib

import numpy as np


# sdf bias that puts roughly half of the sampled volume inside the surface, per category
SDF_BIAS = {"fg": 0.3, "bg": 0.02}


def synth_tensor(name, shape, seed=0, category="fg"):
    """He-uniform weights, small biases; a few named overrides keep the field well conditioned."""
    rs = np.random.RandomState((zlib.crc32(name.encode()) + 7919 * seed) % (2**31))
    shape = tuple(shape)
    leaf = name.split(".")[-1]
    if name.endswith("log_gauss"):
        return (np.log(0.08) + 0.3 * rs.uniform(-1, 1, shape)).astype(np.float32)
    if name in ("logibeta",):
        return np.full(shape, -np.log(0.05), np.float32)
    if name in ("warp.logibeta",):
        return np.full(shape, -np.log(0.01), np.float32)
    if name in ("logscale",):
        return np.full(shape, np.log(0.2), np.float32)
    if name in ("logsigma",):
        return np.zeros(shape, np.float32)
    if name.endswith("base_quat"):
        q = np.zeros(shape, np.float32)
        q[..., 0] = 1
        return q
    if leaf == "scale":  # ScaleLayer buffer
        return np.full(shape, 0.1, np.float32)
    if len(shape) == 2 and leaf == "weight":
        fan_in = shape[1]
        gain = 1.0
        if "inst_embedding" in name or "mapping.weight" in name:
            return (0.5 * rs.standard_normal(shape)).astype(np.float32)
        if name.startswith("sdf."):
            gain = 0.25
        if ".articulation.so3.2" in name or ".articulation.trans.2" in name:
            gain = 0.6
        if ".post_warp." in name and ".linear_final" in name:
            gain = 0.1  # soft deformation of a few % of the object size (x + 0.1 * mlp, warping.py:165), as after training
        a = gain * np.sqrt(6.0 / fan_in)  # He-uniform: keeps ReLU activations O(1) through 8 layers
        return rs.uniform(-a, a, shape).astype(np.float32)
    if leaf == "bias":
        if name.startswith("sdf."):
            return np.full(shape, SDF_BIAS[category], np.float32)
        return rs.uniform(-0.1, 0.1, shape).astype(np.float32)
    return (0.1 * rs.standard_normal(shape)).astype(np.float32)


def synth_state(shapes, seed=0, category="fg"):
    """shapes: dict name -> shape.  Returns dict name -> float32 ndarray."""
    return {k: synth_tensor(k, s, seed, category) for k, s in shapes.items()}


def synth_rays(M, N, seed=1, T=40, nvid=1, center=128.0, spread=60.0):
    """Per-frame ray batch in the layout of samples_dict (nnutils/nerf.py:530-578)."""
    rs = np.random.RandomState(1000 + seed)
    hxy = np.ones((M, N, 3), np.float32)
    hxy[..., :2] = center + spread * rs.uniform(-1, 1, (M, N, 2))
    fx = 200.0
    Kinv = np.zeros((M, 3, 3), np.float32)
    Kinv[:, 0, 0] = 1 / fx
    Kinv[:, 1, 1] = 1 / fx
    Kinv[:, 0, 2] = -128.0 / fx
    Kinv[:, 1, 2] = -128.0 / fx
    Kinv[:, 2, 2] = 1
    near_far = np.tile(np.array([[0.35, 0.85]], np.float32), (M, 1))
    # frame pairs stay adjacent (flip_pair, nnutils/nerf.py:929-946)
    frame_id = (np.arange(M) % (T * nvid)).astype(np.int64)
    inst_id = (frame_id // T).astype(np.int64)
    # camera: small per-frame rotation about a random axis, object 3 units in front (x logscale)
    ax = rs.standard_normal((M, 3))
    ax /= np.linalg.norm(ax, axis=-1, keepdims=True)
    ang = 0.15 * rs.uniform(-1, 1, (M, 1))
    q = np.concatenate([np.cos(ang / 2), np.sin(ang / 2) * ax], -1)
    t = np.array([[0.0, 0.0, 3.0]]) + 0.1 * rs.uniform(-1, 1, (M, 3))
    field2cam = np.concatenate([q, t], -1).astype(np.float32)
    return dict(hxy=hxy, Kinv=Kinv, near_far=near_far, frame_id=frame_id, inst_id=inst_id, field2cam=field2cam)
