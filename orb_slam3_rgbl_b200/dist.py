"""Multi-GPU plumbing: sequences are independent, so ranks never exchange data on the hot path.
One process per GPU; `sequence s -> rank s mod world`; the only collective is a MAX-reduce of the elapsed time
(and a SUM of processed frames) for the throughput report.  Backend: NCCL on GPUs, gloo in CPU tests."""
from __future__ import annotations

import os


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def sequences_of_rank(n_sequences: int, rank: int, world: int):
    """Static partition of sequence ids over ranks (round robin)."""
    return [s for s in range(n_sequences) if s % world == rank]


class Reporter:
    def __init__(self, backend: str | None = None, device=None):
        self.rank, self.world, self.local_rank = env_rank()
        self.dist = None
        self.device = device
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29511")
            if not dist.is_initialized():
                kw = {}
                if backend == "nccl" and device is not None:
                    kw["device_id"] = device
                dist.init_process_group(backend or "gloo", **kw)
            self.dist = dist

    def barrier(self):
        if self.dist:
            self.dist.barrier()

    def _tensor(self, vals):
        import torch
        t = torch.tensor(vals, dtype=torch.float64)
        if self.device is not None and self.dist and self.dist.get_backend() == "nccl":
            t = t.to(self.device)
        return t

    def max_over_ranks(self, x: float) -> float:
        if not self.dist:
            return x
        t = self._tensor([x])
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x: float) -> float:
        if not self.dist:
            return x
        t = self._tensor([x])
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def gather_floats(self, x: float):
        """x of every rank, as a list (rank order)."""
        if not self.dist:
            return [x]
        import torch
        t = self._tensor([x])
        out = [torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    def throughput(self, frames_this_rank: int, elapsed_s_this_rank: float) -> float:
        """Whole-job frames/s = all frames of all ranks / the slowest rank's time."""
        return self.sum_over_ranks(float(frames_this_rank)) / self.max_over_ranks(elapsed_s_this_rank)

    def close(self):
        if self.dist and self.dist.is_initialized():
            self.dist.destroy_process_group()
