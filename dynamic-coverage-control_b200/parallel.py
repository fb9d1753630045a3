"""Multi-GPU plumbing of the learner: one process per GPU, env instances sharded across ranks with
no data-path collective during the rollout; parameters replicated; per PPO epoch ONE all-reduce (SUM) of the flat
actor+critic gradient buffer, plus one all-reduce of 4 float64 statistics per update and one of the per-epoch loss
sums at its end.  `torch.distributed` (NCCL on GPUs; gloo in the CPU tests) is the transport."""
import os

import torch


class Comm:
    """Process-group view used by MAPPOTrainer.  world == 1 makes every collective a no-op."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist = dist
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.world = dist.get_world_size(group)
            self.rank = dist.get_rank(group)
        else:
            self.world, self.rank = 1, 0
        self.calls = 0
        self.bytes = 0
        self.timing = False        # bench.py: bracket every collective with CUDA events on the launching stream
        self._events = []

    def all_reduce_sum_(self, tensor):
        """In-place SUM all-reduce; returns the tensor."""
        if self.world > 1:
            ev = None
            if self.timing and tensor.is_cuda:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            self._dist.all_reduce(tensor, op=self._dist.ReduceOp.SUM, group=self.group)
            if ev is not None:
                ev[1].record()
                self._events.append(ev)
            self.calls += 1
            self.bytes += tensor.numel() * tensor.element_size()
        return tensor

    def collective_ms(self):
        """Device time spent inside the timed collectives since the last call (synchronises); includes waiting for peers."""
        if not self._events:
            return 0.0
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in self._events)
        self._events = []
        return ms

    def broadcast_(self, tensor, src=0):
        if self.world > 1:
            self._dist.broadcast(tensor, src=src, group=self.group)
        return tensor

    def barrier(self):
        if self.world > 1:
            self._dist.barrier(group=self.group)


def shard_envs(n_envs_total, world, rank):
    """Contiguous partition of the env axis: rank r owns [lo, hi).  Sizes differ by at most one."""
    base, rem = divmod(int(n_envs_total), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*)."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend, **kwargs)
    return Comm()
