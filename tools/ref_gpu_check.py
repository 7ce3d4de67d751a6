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

    def fwd():
        s = field.get_samples(Kinv, batch)
        feat, deltas, aux = field.query_field(s, flow_thresh=None)
        return render_pixel(feat, deltas)

    def step():
        field.zero_grad()
        r = fwd()
        (r["rgb"].sum() + r["mask"].sum() + r["flow"].sum() * 1e-3).backward()

    for name, fn, ctx in (("forward", fwd, torch.no_grad()), ("forward+backward", step, torch.enable_grad())):
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
        print(f"reference on CUDA fg-bob {M}x{N}x{D} {name}: {dt*1e3:.1f} ms -> {M*N*D/dt:.3e} ray-samples/s, "
              f"peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")


if __name__ == "__main__":
    main()
