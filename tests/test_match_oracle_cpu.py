"""CPU, build container only: the oracle's global_match against the unmodified reference's FeatureNeRF.global_match
(lab4d/nnutils/feature.py:152-205) under the same random draw - the pin of the checker of the match kernels."""
import os
import sys
import types

import pytest
import torch

import lab4d_oracle as O

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")


def test_global_match_oracle_is_the_reference():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "ref_shims"))
    import _install  # noqa: F401  (stubs of the reference's cold-path imports)
    import ref_harness  # noqa: F401
    from lab4d.nnutils.feature import FeatureNeRF

    g = torch.Generator().manual_seed(3)
    M, N, D = 3, 5, 40
    feat_px = torch.nn.functional.normalize(torch.randn(M, N, 16, generator=g), dim=-1)
    feat_can = torch.nn.functional.normalize(torch.randn(M, N, D, 16, generator=g), dim=-1)
    xyz = 0.2 * torch.randn(M, N, D, 3, generator=g)
    stub = types.SimpleNamespace(logsigma=torch.tensor([0.7]))
    for K in (1024, 128):
        torch.manual_seed(5)
        ref = FeatureNeRF.global_match(stub, feat_px, feat_can, xyz, num_candidates=K)
        torch.manual_seed(5)
        idx = torch.randperm(M * N * D)[:min(K, M * N * D)]
        ours = O.global_match(feat_px, feat_can, xyz, stub.logsigma, idx)
        assert torch.equal(ours, ref)
