"""Drop-in for the reference's dqtorch extension module `quaternion` (lab4d/third_party/quaternion/quaternion.py:1-140 over
src/quaternion.cu:29-217): `quaternion_mul(a, b)` on (B, 3|4) operands -> (B, 4) and `quaternion_conjugate(q)` on (B, 4),
differentiable twice like the reference's three nested autograd Functions - on the kernels of csrc/quat.cu through the C ABI.
`lab4d/utils/quat_transform.py:10-16` does `from quaternion import quaternion_conjugate, quaternion_mul`: putting this module on
the path under that name (or `nnutils.install(dqtorch=True)`) replaces the extension.  CUDA fp32 only; no CPU path."""
import ctypes as C

import torch

from . import _lib


def _call(fn_name, dev, *args):
    h = _lib.handle_for(dev)
    rc = getattr(h.lib, fn_name)(h.h, *args, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    h.check(rc, fn_name)


def _c(t):
    """fp32 contiguous operand.  An operand that already is one is handed on AS IS (not detached): the Functions save their
    inputs, and autograd reconnects saved inputs to the graph when the backward is differentiated again (the reference saves
    `inputs.contiguous()` the same way, quaternion.py:62-75)."""
    if not t.is_cuda:
        raise RuntimeError("lab4d_b200.quaternion: CUDA tensors only (no CPU path)")
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


class _MulBackward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grad, a, b):
        g, a, b = _c(grad), _c(a), _c(b)
        B, D1, D2 = a.shape[0], a.shape[1], b.shape[1]
        ga, gb = torch.empty(B, D1, device=a.device), torch.empty(B, D2, device=a.device)
        _call("b200r_quat_mul_bwd", a.device, g.data_ptr(), a.data_ptr(), b.data_ptr(), ga.data_ptr(), gb.data_ptr(), B, D1, D2)
        ctx.save_for_backward(g, a, b)
        return ga, gb

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, u1, u2):
        g, a, b = ctx.saved_tensors
        B, D1, D2 = a.shape[0], a.shape[1], b.shape[1]
        u1, u2 = _c(u1), _c(u2)
        gg, gga, ggb = torch.empty(B, 4, device=a.device), torch.empty(B, D1, device=a.device), torch.empty(B, D2, device=a.device)
        _call("b200r_quat_mul_bwd_bwd", a.device, u1.data_ptr(), u2.data_ptr(), g.data_ptr(), a.data_ptr(), b.data_ptr(), gg.data_ptr(),
              gga.data_ptr(), ggb.data_ptr(), B, D1, D2)
        return gg, gga, ggb


class _Mul(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = _c(a), _c(b)
        if a.dim() != 2 or b.dim() != 2 or a.shape[0] != b.shape[0] or a.shape[1] not in (3, 4) or b.shape[1] not in (3, 4):
            raise RuntimeError(f"quaternion_mul: operands must be (B, 3|4) with equal B, got {tuple(a.shape)} and {tuple(b.shape)}")
        out = torch.empty(a.shape[0], 4, device=a.device)
        _call("b200r_quat_mul_fwd", a.device, a.data_ptr(), b.data_ptr(), out.data_ptr(), a.shape[0], a.shape[1], b.shape[1])
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, grad):
        a, b = ctx.saved_tensors
        return _MulBackward.apply(grad, a, b)


class _Conj(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q):
        q = _c(q)
        if q.dim() != 2 or q.shape[1] != 4:
            raise RuntimeError(f"quaternion_conjugate: operand must be (B, 4), got {tuple(q.shape)}")
        out = torch.empty_like(q)
        _call("b200r_quat_conj", q.device, q.data_ptr(), out.data_ptr(), q.shape[0])
        return out

    @staticmethod
    def backward(ctx, grad):
        return _Conj.apply(grad)


quaternion_mul = _Mul.apply
quaternion_conjugate = _Conj.apply
