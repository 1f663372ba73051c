"""Data-parallel plumbing of the training step: one process per GPU, `torch.distributed` (NCCL over NVLink on the
GPU box, gloo in the CPU tests).  The batch is split over ranks; the only exchange steps are
the gradient all-reduce after each backward pass (one flat bucket: 8.5 MB critic / 7.2 MB generator, latency-bound on
NVLink 5) and the per-channel sums of synchronised BatchNorm."""
import torch


class Comm:
    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.enabled = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.enabled else 1
        self.rank = dist.get_rank(group) if self.enabled else 0
        self._events = []      # (start, stop) CUDA events around every collective since reset_timers()
        self.timing = False

    def reset_timers(self):
        """Start recording the device time of every collective (CUDA events on the launching stream)."""
        self._events, self.timing = [], True

    def collective_ms(self):
        """Device milliseconds spent in collectives since reset_timers() (synchronises)."""
        if not self._events:
            return 0.0
        torch.cuda.synchronize()
        return float(sum(a.elapsed_time(b) for a, b in self._events))

    def _all_reduce(self, t):
        if self.timing and t.is_cuda:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
            b.record()
            self._events.append((a, b))
        else:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)

    def allreduce_sum(self, t):
        """In-place sum over ranks of one tensor."""
        if self.world > 1:
            self._all_reduce(t)
        return t

    def allreduce_grads(self, grads, skip=()):
        """Sums the gradient tensors over ranks through ONE flat bucket (a single collective)."""
        if self.world == 1:
            return grads
        if getattr(grads, "flat", None) is not None and not skip:
            self._all_reduce(grads.flat)          # the model's gradients already live in one flat buffer (train/nets.py FlatVars)
            return grads
        names = [n for n in grads if n not in skip]
        if not names:
            return grads
        flat = torch.cat([grads[n].reshape(-1) for n in names])
        self._all_reduce(flat)
        off = 0
        for n in names:
            k = grads[n].numel()
            grads[n].copy_(flat[off:off + k].view_as(grads[n]))
            off += k
        return grads


def shard_batch(n_global, rank, world):
    """Contiguous split of a global batch: returns (start, stop) of this rank's samples."""
    base, rem = divmod(n_global, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)
