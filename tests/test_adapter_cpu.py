"""CPU, build container only (needs /root/reference): the drop-in adapter reads the architecture and the
per-frame tables out of live reference modules exactly as the golden-vector harness does."""
import os
import sys

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")


def test_adapter_reads_reference_modules():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import ref_harness as H
    import synth
    from lab4d_b200 import nnutils, spec

    mf = H.build_field("fg", "bob")
    field = mf.field_params["fg"]
    cfg = nnutils.config_from_module(field)
    assert cfg == spec.FG_BOB
    assert nnutils.config_from_module(H.build_field("bg", "rigid").field_params["bg"]) == spec.BG
    comp = H.build_field("fg", "comp_skel-quad_dense")
    assert nnutils.config_from_module(comp.field_params["fg"]) == spec.FG_COMP_QUAD
    _, _, _, ctabs, cgraph = H.run_field(comp, "fg", synth.synth_rays(2, 4), 4)
    with torch.no_grad():
        ctab = nnutils.tables_from_module(comp.field_params["fg"], cgraph[3])
    for k in ("dense_t_embed", "inst_dense_fwd", "inst_dense_bwd"):
        assert torch.equal(ctab[k], torch.from_numpy(ctabs[k])), k
    rays = synth.synth_rays(4, 4)
    feat, deltas, rend, tabs, graph = H.run_field(mf, "fg", rays, 8)
    samples = graph[3]
    with torch.no_grad():
        tab = nnutils.tables_from_module(field, samples)
    for k, v in tab.items():
        assert torch.allclose(v, torch.from_numpy(tabs[k]), atol=0, rtol=0), k
    # parameter names the renderer reads exist in the module
    names = dict(field.named_parameters())
    for k in spec.field_param_shapes(cfg):
        assert k in names and tuple(names[k].shape) == tuple(spec.field_param_shapes(cfg)[k]), k
    # no CPU fallback: on a CPU device the patched entry point fails loudly
    field.train()
    with pytest.raises(RuntimeError, match="CUDA"):
        nnutils.query_field(field, samples)
