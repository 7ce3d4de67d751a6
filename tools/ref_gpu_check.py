"""GPU box: the UNMODIFIED reference (baseline/_ref through oracle/ref_shims, pure-torch quaternion ops) on CUDA -
sanity of the travelling copy and the 'reference on the B200 itself' number of SURVEY.md 8(d)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch

import ref_harness as H
import synth


def main():
    dev = "cuda"
    M, N, D = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (128, 16, 128)
    mf = H.build_field("fg", "bob", seed=0).to(dev)
    field = mf.field_params["fg"]
    H.set_n_depth(D)
    rays = synth.synth_rays(M, N, seed=10)
    Kinv, batch = H.make_batch(field, rays, dev)
    from lab4d.utils.render_utils import render_pixel
    import lab4d.nnutils.nerf as rnerf

    # the reference's importance_sampling takes n_depth as a keyword default (nerf.py:697): align it with D for the eval timing
    _imp = rnerf.NeRF.importance_sampling
    rnerf.NeRF.importance_sampling = lambda self, *a, **k: _imp(self, *a, **dict(k, n_depth=D))

    def fwd():
        s = field.get_samples(Kinv, batch)
        feat, deltas, aux = field.query_field(s, flow_thresh=None)
        return render_pixel(feat, deltas)

    def step():
        field.zero_grad()
        r = fwd()
        (r["rgb"].sum() + r["mask"].sum() + r["flow"].sum() * 1e-3).backward()

    def step_full():  # what a training step of the reference evaluates on one field: + eikonal term and feature matching in the loss
        field.zero_grad()
        s = field.get_samples(Kinv, batch)
        feat, deltas, aux = field.query_field(s, flow_thresh=None)
        r = render_pixel(feat, deltas)
        loss = r["rgb"].sum() + r["mask"].sum() + r["flow"].sum() * 1e-3 + r["eikonal"].mean()
        if "xy_reproj" in aux:
            loss = loss + 1e-3 * aux["xy_reproj"].mean()
        loss.backward()

    g = torch.Generator().manual_seed(3)
    batch["feature"] = torch.nn.functional.normalize(torch.randn(M, N, 16, generator=g), dim=-1).to(dev)

    def eval_render():  # lab4d/render.py -> dvr_model.evaluate: eval-mode query_field (importance sampling, masking, normals) + render_pixel
        field.eval()
        try:
            s = field.get_samples(Kinv, batch)
            feat, deltas, aux = field.query_field(s, flow_thresh=None)
            return render_pixel(feat, deltas)
        finally:
            field.train()

    def timed(tag):
        for name, fn, ctx in (("forward", fwd, torch.no_grad()), ("forward+backward", step, torch.enable_grad()),
                              ("forward+backward incl. eikonal + matching in the loss", step_full, torch.enable_grad()),
                              ("eval-mode render (importance sampling + normals)", eval_render, torch.no_grad())):
            torch.cuda.reset_peak_memory_stats()
            with ctx:
                for _ in range(2):
                    fn()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                n = 5
                for _ in range(n):
                    fn()
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / n
            print(f"{tag} fg-bob {M}x{N}x{D} {name}: {dt*1e3:.1f} ms -> {M*N*D/dt:.3e} ray-samples/s, "
                  f"peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)

    timed("reference on CUDA")
    # the same calls with the B200 renderer patched in (lab4d_b200.nnutils.install): eager launches through the reference's own
    # entry points, the reference's per-frame modules (cameras, articulations, embeddings) still torch
    sys.path.insert(0, ROOT)
    from lab4d_b200 import nnutils

    for prec, kw in (("fp16x3", {}), ("fp16", {}), ("fp16x3", dict(dqtorch=True, match_rng="device"))):
        undo = nnutils.install(n_depth=D, operand_dtype=prec, **kw)
        try:
            timed(f"patched reference ({prec}{', quaternion kernels + device candidate draw' if kw else ''})")
        finally:
            undo()


if __name__ == "__main__":
    main()
