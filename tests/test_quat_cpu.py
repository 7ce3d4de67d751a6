"""CPU: the Hamilton-product formulas the quaternion kernels implement (csrc/quat.cu: backward and backward-of-backward of
quaternion_mul with 3- and 4-wide operands, lab4d/third_party/quaternion/src/quaternion.cu:67-199) against first- and
second-order autograd of the pure-torch restatement, fp64."""
import pytest
import torch


def _pad(x):
    return torch.cat([torch.zeros_like(x[..., :1]), x], -1) if x.shape[-1] == 3 else x


def qmul(a, b):
    a, b = _pad(a), _pad(b)
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack((aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw), -1)


def conj(q):
    q = _pad(q)
    return torch.cat((q[..., :1], -q[..., 1:]), -1)


def cut(x, D):
    return x[..., 1:] if D == 3 else x


@pytest.mark.parametrize("D1,D2", [(4, 4), (4, 3), (3, 4), (3, 3)])
def test_quaternion_product_derivatives(D1, D2):
    g = torch.Generator().manual_seed(D1 * 10 + D2)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    a, b, G = rn(9, D1).requires_grad_(True), rn(9, D2).requires_grad_(True), rn(9, 4).requires_grad_(True)
    ga, gb = torch.autograd.grad(qmul(a, b), (a, b), G, create_graph=True)
    assert torch.allclose(ga, cut(qmul(G, conj(b)), D1)) and torch.allclose(gb, cut(qmul(conj(a), G), D2))
    u1, u2 = rn(9, D1), rn(9, D2)
    gG, gga, ggb = torch.autograd.grad((ga, gb), (G, a, b), (u1, u2))
    assert torch.allclose(gG, qmul(u1, b) + qmul(a, u2))
    assert torch.allclose(gga, cut(qmul(G, conj(u2)), D1))
    assert torch.allclose(ggb, cut(qmul(conj(u1), G), D2))
