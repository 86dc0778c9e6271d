"""Frame-batch data parallelism across GPUs: frames are independent, so a job is cut into
contiguous per-rank blocks and NO data-path collective exists.  The only cross-rank traffic is the bookkeeping of
a measurement: a barrier, the max of the per-rank times and the sum of the per-rank frame counts.

Used by bench.py (backend nccl, one process per GPU) and covered by tests/test_sharding_gloo.py (backend gloo,
world size 2, CPU)."""


def frame_range(rank, world, frames_per_gpu):
    """Weak scaling: every rank owns `frames_per_gpu` frames; global frame indices are contiguous per rank."""
    if not (0 <= rank < world) or frames_per_gpu < 0:
        raise ValueError("bad rank/world/frames_per_gpu: %r %r %r" % (rank, world, frames_per_gpu))
    first = rank * frames_per_gpu
    return first, first + frames_per_gpu


def split_frames(n_frames, world):
    """Strong scaling: `n_frames` frames over `world` ranks, block sizes differing by at most one."""
    base, extra = divmod(n_frames, world)
    out, first = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((first, first + n))
        first += n
    return out


class Group:
    """Thin wrapper over torch.distributed (or nothing, for one process)."""

    def __init__(self, dist=None, device=None):
        self.dist = dist
        self.device = device

    @property
    def world(self):
        return self.dist.get_world_size() if self.dist is not None else 1

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def _reduce(self, x, op):
        if self.dist is None:
            return float(x)
        import torch

        t = torch.tensor([float(x)], dtype=torch.float64, device=self.device or "cpu")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX if self.dist is not None else None)

    def sum(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM if self.dist is not None else None)

    def throughput(self, frames_this_rank, ms_this_rank):
        """Whole-job frames/s: frames of all ranks over the slowest rank's time."""
        return self.sum(frames_this_rank) / (self.max(ms_this_rank) * 1e-3)
