"""Data-parallel plumbing: rays shard over ranks, one all-reduce of the flat gradient buffer per step
(the reference's only collective is DDP's bucketed gradient all-reduce, engine/trainer.py:110-115,345).
Frames stay in adjacent pairs on every rank (flip_pair, nnutils/nerf.py:929-946)."""
import torch
import torch.distributed as dist


def shard_frames(M, rank, world):
    """Contiguous block of frame PAIRS for `rank`: returns (start, stop) frame indices, both even."""
    if M % 2:
        raise ValueError("frames come in pairs")
    pairs = M // 2
    base, rem = divmod(pairs, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return 2 * lo, 2 * hi


def shard_batch(rays, tab, rank, world):
    """Slice every per-frame tensor of a ray batch / frame-table dict to this rank's frames."""
    M = rays["hxy"].shape[0]
    lo, hi = shard_frames(M, rank, world)

    def cut(d):
        out = {}
        for k, v in d.items():
            out[k] = v[lo:hi] if (torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == M) else v
        return out

    return cut(rays), cut(tab)


def allreduce_mean_(flat_grad, group=None):
    """DDP semantics: sum over ranks, divide by world size, in place on a flat fp32 buffer."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
        flat_grad.div_(dist.get_world_size(group))
    return flat_grad
