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


class PeerExchange:
    """Symmetric buffer (torch symmetric memory: the same allocation mapped into every rank over
    NVLink / NVSwitch) for the batch-norm exchanges that libbnn_b200 fuses into its finalize kernels
    (bnn_bn_stats_finalize_p2p, bnn_allreduce_f64_p2p).  Every call site (layer x direction) owns a
    flag row and a data slot; layout, identical on every rank:
        [0, 64 * n_sites)   flag rows  uint32[n_sites][16]   (row = call site, column = sending rank)
        then n_sites slots of `slot_doubles` fp64 each
    The epoch of every exchange is the training-step counter `comm.step_dev`, a device word that the
    kernels read themselves.  PyTorch only allocates and exchanges the mappings; the kernels do the
    signalling and the sums."""

    def __init__(self, comm: "Comm", max_cols: int, n_sites: int):
        self.comm = comm
        self.slot_doubles = 2 * max_cols
        self.max_cols = max_cols
        self.n_sites = n_sites
        self.flag_bytes = -(-(64 * n_sites) // 256) * 256
        nbytes = self.flag_bytes + n_sites * self.slot_doubles * 8
        self.buf, self.base, self.handle = _symmetric_buffer(comm, nbytes)
        self.last_step = [0] * n_sites
        torch.cuda.synchronize()
        comm.barrier()  # every flag row is zero before anyone signals

    def begin(self, site: int, n_doubles: int):
        """Returns (this rank's slot of call site `site` as an fp64 tensor, call descriptor)."""
        assert n_doubles <= self.slot_doubles and 0 <= site < self.n_sites
        # A call site is used once per step, and between two uses of one site every rank passes another site's
        # barrier — that is what makes refilling the slot safe (p2p.cu).  Code that drives a layer outside
        # Network.train_step may use the SAME site twice in a row (a stand-alone BatchNormLayer called repeatedly):
        # then nothing orders this rank's refill after the peers' reads of the previous contents, so the step is
        # advanced behind a full barrier (host-side, identically on every rank; never taken inside a captured step).
        if self.last_step[site] == self.comm.step:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("PeerExchange: call site used twice within one captured step")
            torch.cuda.current_stream().synchronize()
            self.comm.barrier()
            self.comm.begin_step()
        self.last_step[site] = self.comm.step
        off = self.flag_bytes + site * self.slot_doubles * 8
        slot = self.buf[off: off + n_doubles * 8].view(torch.float64)
        call = dict(data=[b + off for b in self.base], flags=[b + 64 * site for b in self.base],
                    rank=self.comm.rank, world=self.comm.world, step=self.comm.step_dev)
        self.comm.collectives += 1
        return slot, call


def _symmetric_buffer(comm: "Comm", nbytes: int):
    """(local uint8 tensor, [address of every rank's copy as mapped into this process])."""
    import torch.distributed._symmetric_memory as symm
    group = comm.group if comm.group is not None else dist.group.WORLD
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            symm.enable_symm_mem_for_group(group.group_name)
    except Exception:
        pass
    buf = symm.empty(nbytes, dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device()))
    buf.zero_()
    handle = symm.rendezvous(buf, group)
    base = [int(p) for p in handle.buffer_ptrs]
    assert len(base) == comm.world and base[comm.rank] == buf.data_ptr()
    return buf, base, handle


def dw_exchange_layout(shapes, world: int, align: int = 1024):
    """Byte layout of the DwExchange arena (identical on every rank): returns (flag bytes, per-layer
    dicts with K, N, rows_per_owner, w_off, stage_off, total bytes).  rows_per_owner is the number of
    128-row GEMM tiles of W divided evenly over the ranks, rounded up, times 128 — rank r owns rows
    [r*rows_per_owner, (r+1)*rows_per_owner) (the last owners of a short matrix own fewer, or none)."""
    def up(n):
        return -(-n // align) * align
    flag_bytes = up(2 * len(shapes) * 16 * 4)
    off = flag_bytes
    layout = []
    for (k, n) in shapes:
        tiles = -(-k // 128)
        rpo = -(-tiles // world) * 128
        w_off = off
        off += up(k * n * 4)
        s_off = off
        off += up(world * rpo * n * 4)
        layout.append(dict(K=int(k), N=int(n), rows_per_owner=rpo, w_off=w_off, stage_off=s_off))
    return flag_bytes, layout, off


class DwExchange:
    """Symmetric arena for the dW all-reduce that libbnn_b200 fuses into its own kernels
    (`dW` all-reduce-sum before `W -= lr*dW`, layers.py:268).

    Per layer the arena holds, at the same offset on every rank,
        W      fp32 [K, N]                      the layer's latent weights (re-homed here)
        stage  fp32 [world, rows_per_owner, N]  partial dW rows pushed by every rank's dW GEMM
    and a flag pad uint32[2 * n_layers][16] at offset 0 (slot 2l: "rank r has pushed its partial dW of
    layer l", slot 2l+1: "rank r's updated rows of W_l have landed").  Rank r owns rows
    [r*rows_per_owner, (r+1)*rows_per_owner) of every layer's W (rows_per_owner a multiple of the
    GEMM's 128-row tile): the dW GEMM's epilogue stores each row tile into its owner's stage over
    NVLink (reduce-scatter), the owner sums, updates and stores the new rows into every rank's W
    (all-gather).  PyTorch allocates and maps the arena; the kernels move and synchronise the data."""

    ALIGN = 1024

    def __init__(self, comm: "Comm", shapes):
        self.comm = comm
        self.world, self.rank = comm.world, comm.rank
        self.shapes = [(int(k), int(n)) for k, n in shapes]
        n_layers = len(self.shapes)
        self.flag_bytes, self.layout, self.nbytes = dw_exchange_layout(self.shapes, self.world, self.ALIGN)
        self.buf, self.base, self.handle = _symmetric_buffer(comm, self.nbytes)
        self.counters = torch.zeros(n_layers, dtype=torch.int32, device=self.buf.device)
        # all-gather half: "ce" = copy engines on a side stream (overlaps the backward pass),
        # "kernel" = peer stores from the owner-update kernel itself
        self.copy_engine_broadcast = os.environ.get("BNN_DW_BCAST", "ce").lower() != "kernel"
        self.side = torch.cuda.Stream()
        self.reduced = [torch.cuda.Event() for _ in range(n_layers)]
        self.copied = [torch.cuda.Event() for _ in range(n_layers)]
        self.deferred = []    # layers updated locally whose copy-engine broadcast has not started yet
        self.in_flight = []   # layers whose rows are on their way to the peers (this step)
        self.pending = []     # layers whose updated rows have not been awaited yet (this step)
        torch.cuda.synchronize()
        comm.barrier()  # every flag pad is zero before anyone signals

    def weight_view(self, layer: int) -> torch.Tensor:
        L = self.layout[layer]
        return self.buf[L["w_off"]: L["w_off"] + L["K"] * L["N"] * 4].view(torch.float32).view(L["K"], L["N"])

    def w_ptrs(self, layer: int):
        return [b + self.layout[layer]["w_off"] for b in self.base]

    def stage_ptrs(self, layer: int):
        return [b + self.layout[layer]["stage_off"] for b in self.base]

    def flag_ptrs(self):
        return list(self.base)

    def start_broadcasts(self) -> None:
        """Queue the copy-engine all-gather of every layer whose owner update has been queued: the copy stream
        waits for that update (event `reduced[li]`, recorded by the layer on whichever stream ran it), then
        copies this rank's rows (and the "landed" flag) into every peer."""
        if not self.deferred:
            return
        from . import kernels as K_
        for li in self.deferred:
            L = self.layout[li]
            with torch.cuda.stream(self.side):
                self.side.wait_event(self.reduced[li])
                K_.peer_broadcast_rows(self.w_ptrs(li), self.flag_ptrs(), 2 * li + 1, self.rank,
                                       self.world, L["K"], L["N"], L["rows_per_owner"],
                                       self.comm.step_dev)
                self.copied[li].record(self.side)
            self.in_flight.append(li)
        self.deferred.clear()

    def finish_step(self) -> None:
        """End of a training step: join the copy-engine stream, then ONE kernel waits until every
        rank's updated rows of every layer exchanged this step have landed in the local W."""
        self.start_broadcasts()
        main = torch.cuda.current_stream()
        for li in self.in_flight:
            main.wait_event(self.copied[li])
        self.in_flight.clear()
        if self.pending:
            from . import kernels as K_
            lo, n = min(self.pending), len(self.pending)
            assert sorted(self.pending) == list(range(lo, lo + n))
            K_.peer_wait(self.base[self.rank], 2 * lo + 1, 2, n, self.rank, self.world,
                         self.comm.step_dev)
            self.pending.clear()


class Comm:
    """Sum all-reduce over a torch.distributed process group (any backend: nccl on GPUs, gloo in
    the CPU tests of the host-side logic)."""

    def __init__(self, group=None):
        assert dist.is_initialized(), "init_process_group first (see init_from_env)"
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.collectives = 0
        self.nccl_calls = 0   # torch.distributed collectives issued through this object (bench.py reports them)
        self.global_batch = None  # rows of the current step over all ranks (None: local * world)
        self.peer_exchange = None  # PeerExchange once enable_peer_exchange() succeeded
        self.dw_exchange = None    # DwExchange once enable_dw_exchange() succeeded
        # training-step counter: the epoch of every peer-memory exchange.  The device word is what
        # the kernels read (so a captured step replays correctly); `step` mirrors it on the host.
        self.step = 0
        self.step_dev = None

    def begin_step(self) -> None:
        """Advance the step counter (host mirror + device word); called once per training step by
        Network.train_step_async on every rank."""
        if self.peer_exchange is None and self.dw_exchange is None:
            return
        if self.step_dev is None:
            self.step_dev = torch.zeros(1, dtype=torch.int32, device="cuda")
        from . import kernels as K_
        self.step += 1
        K_.step_tick(self.step_dev)

    def finish_step(self) -> None:
        if self.dw_exchange is not None:
            self.dw_exchange.finish_step()

    def enable_peer_exchange(self, max_cols: int, n_sites: int = 2) -> bool:
        """Route the small batch-norm exchanges through NVLink peer memory instead of NCCL.  Returns
        False (and keeps NCCL) where symmetric memory is unavailable; BNN_P2P=0 disables it."""
        if self.world == 1 or os.environ.get("BNN_P2P", "1") == "0" or not torch.cuda.is_available():
            return False
        if dist.get_backend(self.group) != "nccl":
            return False
        try:
            self.peer_exchange = PeerExchange(self, max_cols, n_sites)
        except Exception as exc:  # pragma: no cover - depends on the platform
            print(f"[mlpcode.parallel] peer exchange unavailable, staying on NCCL: {exc}")
            self.peer_exchange = None
        # all ranks must agree, or the kernels would wait for peers that never signal
        if not self._all_agree(self.peer_exchange is not None):
            self.peer_exchange = None
        if self.peer_exchange is not None and self.step_dev is None:
            self.step_dev = torch.zeros(1, dtype=torch.int32, device="cuda")
        return self.peer_exchange is not None

    def _all_agree(self, ok: bool) -> bool:
        t = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return int(t.item()) == 1

    def enable_dw_exchange(self, shapes) -> bool:
        """Route the dW all-reduce through the fused reduce-scatter epilogue + owner update over NVLink
        peer memory.  BNN_DW_P2P=0 keeps the NCCL all-reduce."""
        if (self.world == 1 or os.environ.get("BNN_DW_P2P", "1") == "0" or not torch.cuda.is_available()
                or dist.get_backend(self.group) != "nccl" or self.world > 16):
            return False
        try:
            self.dw_exchange = DwExchange(self, shapes)
        except Exception as exc:  # pragma: no cover - depends on the platform
            print(f"[mlpcode.parallel] dW peer exchange unavailable, staying on NCCL: {exc}")
            self.dw_exchange = None
        if not self._all_agree(self.dw_exchange is not None):
            self.dw_exchange = None
        if self.dw_exchange is not None and self.step_dev is None:
            self.step_dev = torch.zeros(1, dtype=torch.int32, device="cuda")
        return self.dw_exchange is not None

    def global_rows(self, local_rows: int) -> int:
        """Batch-norm statistics and the loss gradient divide by the GLOBAL batch (layers.py:81-82,
        loss.py:66).  Equal shards are assumed unless the driver states the global batch."""
        return self.global_batch if self.global_batch is not None else local_rows * self.world

    def allreduce_sum(self, t: torch.Tensor):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            self.collectives += 1
            self.nccl_calls += 1
        return t

    def allreduce_sum_async(self, t: torch.Tensor):
        """Start the all-reduce on the communication stream and return a handle whose `.wait()` makes
        the CURRENT stream wait for it (no host block), so the next layers' kernels overlap it."""
        if self.world == 1:
            return None
        self.collectives += 1
        self.nccl_calls += 1
        return dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def allgather(self, t: torch.Tensor) -> torch.Tensor:
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t, group=self.group)
        self.nccl_calls += 1
        return torch.stack(out)

    def barrier(self):
        if self.world > 1:
            dist.barrier(group=self.group)
            self.nccl_calls += 1


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
