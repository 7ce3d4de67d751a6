"""Per-frame quaternion / dual-quaternion helpers for the host-side prologue (M x B rows per step;
the per-sample quaternion math lives in the CUDA kernels).  Conventions of
lab4d/utils/quat_transform.py: real part first, dual quaternion = (real, dual) pair."""
import torch


def qmul(a, b):
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw], -1)


def qconj(q):
    return torch.cat([q[..., :1], -q[..., 1:]], -1)


def dq_mul(a, b):
    return qmul(a[0], b[0]), qmul(a[0], b[1]) + qmul(a[1], b[0])


def dq_inv(a):
    return qconj(a[0]), qconj(a[1])


def dq_translation(dq):
    """translation of a unit dual quaternion: 2 (q_d q_r*)_xyz"""
    return 2 * qmul(dq[1], qconj(dq[0]))[..., 1:]
