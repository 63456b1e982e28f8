"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch).

The reference is single-process; batch sharding is what this build adds.  To stay
results-equal to the reference's full-batch step, three things are made global:

  1. batch-norm statistics — per-column (sum z, sum z^2) all-reduced as fp64 (exact integers for
     +-1 x +-1 layers, hence order independent), mean/var divide by the global batch;
  2. the BN backward reductions (sum delta, sum delta*(z-mu)) and the loss gradient's 1/B;
  3. dW — all-reduced in fp32 before `W -= lr*dW`.

Fault-injection sweeps and inference shard by trial / by row with no communication at all; only the
final gather of per-trial accuracies uses the process group.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


class Comm:
    """Sum all-reduce over a torch.distributed process group (any backend: nccl on GPUs, gloo in
    the CPU tests of the host-side logic)."""

    def __init__(self, group=None):
        assert dist.is_initialized(), "init_process_group first (see init_from_env)"
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.collectives = 0
        self.global_batch = None  # rows of the current step over all ranks (None: local * world)

    def global_rows(self, local_rows: int) -> int:
        """Batch-norm statistics and the loss gradient divide by the GLOBAL batch (layers.py:81-82,
        loss.py:66).  Equal shards are assumed unless the driver states the global batch."""
        return self.global_batch if self.global_batch is not None else local_rows * self.world

    def allreduce_sum(self, t: torch.Tensor):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            self.collectives += 1
        return t

    def allreduce_sum_async(self, t: torch.Tensor):
        """Start the all-reduce on the communication stream and return a handle whose `.wait()` makes
        the CURRENT stream wait for it (no host block), so the next layers' kernels overlap it."""
        if self.world == 1:
            return None
        self.collectives += 1
        return dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def allgather(self, t: torch.Tensor) -> torch.Tensor:
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t, group=self.group)
        return torch.stack(out)

    def barrier(self):
        if self.world > 1:
            dist.barrier(group=self.group)


def init_from_env(backend: str = None) -> Comm | None:
    """Join the torchrun rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*) if there is one."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None
    if not dist.is_initialized():
        local = int(os.environ.get("LOCAL_RANK", "0"))
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return Comm()


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [start, stop) share of `n_items` independent units for `rank`: the first
    n % world ranks take one extra item."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)
