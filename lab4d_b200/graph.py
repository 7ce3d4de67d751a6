"""CUDA-graph capture of a whole renderer step (static shapes): the training step of a 2048-ray batch is a few hundred
launches (two fused kernels, compositing, the per-frame chain's small torch ops); replaying them as one graph keeps the
host out of the step, which matters at the strong-scaling shape (131 k samples per GPU)."""
import torch


class GraphedStep:
    """Capture `fn()` after `warmup` eager calls on a side stream; `replay()` re-runs it and returns the SAME output
    tensors (their contents are overwritten).  `fn` must use static input tensors (update them in place between replays)."""

    def __init__(self, fn, warmup=3, device=None):
        self.fn = fn
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fn()

    def replay(self):
        self.graph.replay()
        return self.out
