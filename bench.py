#!/usr/bin/env python
"""bench.py — BNN MLP training throughput on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Own arm: K timed steps of the 784-4096-4096-4096-10 binarized MLP (BASELINE configs[1]), batch 8192
per GPU (weak scaling; N = 8 is configs[2]'s global batch 65536), through mlpcode.network.Network
on libbnn_b200 kernels.  `value` is timed with the inputs resident in HBM; `e2e` repeats the
measurement through the public per-batch call with pinned HOST buffers (H2D of X and y, D2H of the
loss inside the timed region).  `roofline` describes the dominant kernel, timed live with CUDA events
on the launching stream; `binary_gemm` is the +-1 x +-1 forward GEMM the metric's second half names;
`cpu_baseline` is the oracle port of the reference NumPy path timed on this box's host cores.

Reference arm (--impl reference): the reference's CPU path for the same workload — the NumPy oracle
port (the reference itself is Python + CuPy/h5py imports and cannot travel to the GPU box; the port is
pinned bit-exactly against it by oracle/gen_golden.py) with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO / "deep-neural-net_b200"))

UNITS = [784, 4096, 4096, 4096, 10]
BATCH = 8192
LR = 0.01
METRIC, UNIT = "BNN MLP train samples/s", "samples/s"
ROOFLINE_PROFILE = REPO / "profiles" / "r02_roofline_kernel.json"  # written by tools/ncu_table.py --roofline


def workload_config(n_gpus):
    return {
        "workload": "784-4096-4096-4096-10 binarized MLP training step, batch 8192 per GPU "
                    "(BASELINE configs[1]; at N=8 the global batch is configs[2]'s 65536)",
        "units": UNITS, "batch_per_gpu": BATCH, "global_batch": BATCH * n_gpus,
        "parallelism": f"dp{n_gpus}", "lr": LR,
        "l2": "per-step working set ~1.3 GB of activations/operands >> 126 MB L2 (no flush needed)",
    }


def synthetic_batch(seed, batch):
    import numpy as np
    rng = np.random.default_rng(seed)
    X = rng.random((batch, UNITS[0]), dtype=np.float32) * 2 - 1
    Y = np.eye(10, dtype=np.int8)[rng.integers(0, 10, batch)]
    return X, Y


def synthetic_params():
    import numpy as np
    rng = np.random.default_rng(0)
    return [(rng.standard_normal((k, n), dtype=np.float32) * np.float32(np.sqrt(1.0 / k)))
            for k, n in zip(UNITS[:-1], UNITS[1:])]


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        n = max((p.get("num_threads", 1) for p in threadpool_info()), default=1)
        return int(n)
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference NumPy path)
# ------------------------------------------------------------------------------------------------
def cpu_train_samples_per_s(steps, warmup, budget_s):
    """Times oracle.train_step at the workload's own batch (BATCH = 8192, BASELINE configs[1]); the time budget
    only ever cuts the NUMBER of steps, never the batch.  Returns (samples/s, ms/step measured, steps timed,
    warm-up steps done, sample text)."""
    sys.path.insert(0, str(REPO / "oracle"))
    import numpy as np
    import bnn_oracle as O
    W = synthetic_params()
    params = dict(W=[w.copy() for w in W],
                  bn=[dict(gamma=np.ones(n, np.float32), beta=np.zeros(n, np.float32),
                           mu=np.zeros(n, np.float32), sigma=np.ones(n, np.float32)) for n in UNITS[1:]])
    X, Y = synthetic_batch(2, BATCH)
    t_start = time.perf_counter()
    warm_done, last = 0, None
    for _ in range(max(warmup, 1)):
        # the first step pays page faults / thread-pool start-up; stop warming once half the budget would go
        if last is not None and (time.perf_counter() - t_start) + last > 0.25 * budget_s:
            break
        t0 = time.perf_counter()
        O.train_step(params, X, Y, LR)
        last = time.perf_counter() - t0
        warm_done += 1
    times = []
    for _ in range(steps):
        if times and (time.perf_counter() - t_start) + max(times) > budget_s:
            break
        t0 = time.perf_counter()
        O.train_step(params, X, Y, LR)
        times.append(time.perf_counter() - t0)
    dt = sum(times)
    sample = (f"{len(times)} train steps at the full batch {BATCH} after {warm_done} warm-up, NumPy "
              f"{np.__version__} oracle port of network.py:372-392"
              + ("" if len(times) == steps else f" ({steps} steps asked; cut by the {budget_s:.0f} s budget)"))
    return BATCH * len(times) / dt, 1e3 * dt / len(times), len(times), warm_done, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # torchrun pins every rank to OMP_NUM_THREADS=1; the reference arm uses all host threads
    ncpu = os.cpu_count() or 1
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        if os.environ.get(var, "") in ("", "1"):
            os.environ[var] = str(ncpu)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=ncpu)
    except Exception:
        pass
    val, ms_step, steps_done, warm_done, sample = cpu_train_samples_per_s(args.steps, args.warmup, budget_s=280.0)
    cores = host_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps_done, "warmup": warm_done, "ms_per_step": ms_step,
        "steps_requested": args.steps, "warmup_requested": args.warmup,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the CPU path times steps at batch_per_gpu = 8192 on rank 0's host cores whatever --gpus says (the reference "
                "has no multi-process CPU path); ms_per_step is the measured mean of the timed steps",
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.samples if t0 <= t <= t1] or [r for (_, r) in self.samples]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            parts = [p.strip() for p in r.split(",")]
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def measured_peaks():
    """(HBM GB/s, bf16 TFLOP/s burst, bf16 TFLOP/s sustained, source) from the driver-written file."""
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        burst = d.get("bf16_tflops", 1590.0)
        return d.get("hbm_gbs", 6650.0), burst, d.get("bf16_tflops_sustained", burst), "MEASURED_PEAKS.json"
    return 6650.0, 1590.0, 1590.0, "B200_PROFILING.md fallback"


def profiled_traffic():
    """dram__bytes_read + dram__bytes_write per launch of the roofline kernel, from the committed ncu --set full
    capture of this round (never measured in this run: ncu cannot run inside a benchmark)."""
    try:
        d = json.loads(ROOFLINE_PROFILE.read_text())
        return d["dram_bytes_per_launch"], f"profiles/{ROOFLINE_PROFILE.name}: {d['source']}"
    except Exception:
        return None, "no committed ncu capture of this kernel"


def measure_int8_peak(torch):
    """cuBLASLt int8 GEMM 8192^3 (library call, outside every timed region): the int8 tensor-pipe
    denominator SURVEY §8(d) asks the builder to measure."""
    try:
        n = 8192
        a = torch.randint(-1, 2, (n, n), device="cuda", dtype=torch.int8)
        b = torch.randint(-1, 2, (n, n), device="cuda", dtype=torch.int8)
        for _ in range(3):
            torch._int_mm(a, b)
        best = 1e9
        for _ in range(10):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            torch._int_mm(a, b)
            e.record()
            e.synchronize()
            best = min(best, s.elapsed_time(e))
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    except Exception:
        return None


def run_gpu(args):
    import numpy as np
    import torch
    from mlpcode import kernels as K
    from mlpcode import parallel
    from mlpcode.activation import ActivationFuncs as af
    from mlpcode.loss import LossFuncs as lf
    from mlpcode.network import Network

    comm = parallel.init_from_env("nccl")
    world = comm.world if comm else 1
    rank = comm.rank if comm else 0
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    nn = Network(UNITS, useGpu=True, binarized=True, useBatchNorm=True)
    nn.compile(lr=LR, hiddenAf=af.sign, outAf=af.softmax, lossF=lf.cross_entropy)
    nn.load_weights(synthetic_params())
    if comm:
        nn.set_data_parallel(comm)
    # every step is replayed as one CUDA graph (BNN_STEP_GRAPH=0: ~90 eager launches per step);
    # profiling runs stay eager so that ncu sees plain launches
    use_graph = os.environ.get("BNN_STEP_GRAPH", "1") != "0" and not args.profile
    nn.enable_step_graph(use_graph)
    Xh, Yh = synthetic_batch(100 + rank, BATCH)
    Xp, Yp = torch.from_numpy(Xh).pin_memory(), torch.from_numpy(Yh).pin_memory()
    Xd, Yd = Xp.cuda(), Yp.cuda()

    def fence():
        torch.cuda.synchronize()
        if comm:
            comm.barrier()
            torch.cuda.synchronize()

    host_ms = {}

    def timed(step_fn, steps, tag=None):
        fence()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0 = time.perf_counter()
        s.record()
        for _ in range(steps):
            step_fn()
        e.record()
        if tag:  # host time to ENQUEUE the steps (before waiting for the device)
            host_ms[tag] = 1e3 * (time.perf_counter() - h0) / steps
        fence()
        ms = torch.tensor([s.elapsed_time(e)], device="cuda", dtype=torch.float64)
        if comm:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms.item())

    # ---- warm-up, then the device-resident measurement with per-GEMM events ----------------
    for _ in range(max(args.warmup, 3) + (3 if use_graph else 0)):  # 2 eager + capture, then W replays
        nn.train_step_async(Xd, Yd, LR)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = K.launch_count
    t0 = time.perf_counter()
    ms_total = timed(lambda: nn.train_step_async(Xd, Yd, LR), args.steps, "device_resident")
    t1 = time.perf_counter()
    launches = K.launch_count - launches0
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    value = world * BATCH * args.steps / (ms_total * 1e-3)
    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": ms_total / args.steps,
                              "gpu_launches": launches}), flush=True)
        return 0

    # per-kernel accounting: the same K steps launched eagerly with a CUDA-event pair around every
    # GEMM (events cannot be timed inside a graph replay)
    K.gemm_events = []
    ms_eager = timed(lambda: nn.train_step_async(Xd, Yd, LR), args.steps, "eager")
    events, K.gemm_events = K.gemm_events, None
    # per call site (GEMM i of the step, averaged over the steps) on every rank: shows WHICH products stretch in
    # data-parallel runs and whether one GPU is slower than the others
    fence()
    per_step = len(events) // args.steps if args.steps else 0
    site_ms = [0.0] * per_step
    for i, (s_, e_, *_rest) in enumerate(events[: per_step * args.steps]):
        site_ms[i % per_step] += s_.elapsed_time(e_) / args.steps
    site_kinds = [ev[2] for ev in events[:per_step]]
    site_ranks = [site_ms]
    if comm:
        site_ranks = [None] * world
        torch.distributed.all_gather_object(site_ranks, site_ms)

    # ---- end-to-end through the public API with HOST buffers: every step copies its inputs in from
    # pinned memory and its loss out.  Two public calls are measured: the per-batch `train_step`
    # (copy, compute and the blocking loss read serialised, like the reference's __updateBatch) and the
    # batch-stream call `train_stream` (same copies and reads, overlapped with the previous /
    # next step's kernels); the headline e2e is the stream call.
    for _ in range(2):
        nn.train_step(Xp, Yp, LR)
    ms_serial = timed(lambda: nn.train_step(Xp, Yp, LR), args.steps)
    # the step's host->device copy alone, every rank copying at once (what the host side can deliver)
    h2d_tmp = torch.empty_like(Xd)
    h2d_tmp.copy_(Xp, non_blocking=True)
    ms_h2d = timed(lambda: h2d_tmp.copy_(Xp, non_blocking=True), 10) / 10
    del h2d_tmp
    nn.train_stream([(Xp, Yp)] * 3, LR)
    e2e_losses = []
    ms_e2e = timed(lambda: e2e_losses.extend(nn.train_stream([(Xp, Yp)] * args.steps, LR)), 1)
    assert len(e2e_losses) == args.steps and all(l == l for l in e2e_losses)
    e2e_value = world * BATCH * args.steps / (ms_e2e * 1e-3)

    # cross-rank exchanges of one step, counted on one eager step (2 per batch-norm layer + 1 dW
    # exchange per layer); all of them are peer-memory kernels / copy-engine copies, none is an NCCL call
    exchanges_per_step = None
    if comm:
        nn.enable_step_graph(False)
        c0, n0 = comm.collectives, comm.nccl_calls
        nn.train_step_async(Xd, Yd, LR)
        n_coll, n_nccl = comm.collectives - c0, comm.nccl_calls - n0   # before the fence's own barrier
        fence()
        exchanges_per_step = {"count": n_coll,
                              "nccl_calls": n_nccl,
                              "nccl_calls_how": "counted: every torch.distributed collective of the step goes "
                                                "through mlpcode.parallel.Comm, which counts the ones it hands to NCCL",
                              "transport": "NVLink peer memory (own kernels + copy engines)"
                              if comm.dw_exchange is not None else "NCCL"}
    # ---- the other BASELINE workloads (configs 0, 3, 4 and the image sweep), outside the headline's timed
    # region, every rank taking its share of the trials / rows; rank 0 adds the CPU legs at N = 1
    del Xd, Yd, nn
    torch.cuda.empty_cache()
    int8_peak = measure_int8_peak(torch) if rank == 0 else None
    extras = {}
    if not args.no_extras:
        sys.path.insert(0, str(REPO / "tools"))
        import bench_sweep as BS
        cpu = world == 1 and not args.no_cpu
        if rank == 0:
            extras["config1"] = BS.config1_section(steps=50, cpu_steps=20 if cpu else 0)
        extras["inference"] = BS.inference_section(comm, cpu_rows=2048 if cpu else 0, int8_peak_tops=int8_peak)
        extras["weight_sweep"] = BS.weight_sweep_section(comm, n_trials=args.sweep_trials, cpu_trials=1 if cpu else 0,
                                                         int8_peak_tops=int8_peak)
        extras["image_sweep"] = BS.image_sweep_section(comm, n_trials=args.image_trials, cpu_trials=1 if cpu else 0)
    if rank != 0:
        return 0

    # ---- per-kernel accounting from the events recorded inside the timed region ------------------
    groups = {}
    for (s, e, tag, flops, passes) in events:
        g = groups.setdefault(tag, {"ms": 0.0, "n": 0, "flops": 0.0, "exec": 0.0})
        dt = s.elapsed_time(e)
        g["ms"] += dt
        g["n"] += 1
        g["flops"] += flops
        g["exec"] += flops * passes
    hbm_peak, peak_burst, peak_sustained, peak_kind = measured_peaks()
    # which regime did the timed region run in?  The sustained figure belongs to a kernel timed inside a long,
    # power-limited run; a region of a few tens of milliseconds at full SM clocks is the burst regime.
    region_s = ms_total * 1e-3
    sm_frac = (clocks["sm_mhz"] / clocks["sm_max_mhz"]) if clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz") else None
    sustained = region_s >= 2.0 or (sm_frac is not None and sm_frac < 0.9)
    tensor_peak = peak_sustained if sustained else peak_burst
    traffic, traffic_source = profiled_traffic()
    split = groups.get("f16", {"ms": 0.0, "n": 0, "flops": 0.0, "exec": 0.0})
    i8 = groups.get("i8", {"ms": 0.0, "n": 0, "flops": 0.0, "exec": 0.0})
    i8dw = groups.get("i8dw", {"ms": 0.0, "n": 0, "flops": 0.0, "exec": 0.0})
    roofline = None
    if split["n"]:
        ach = split["flops"] / (split["ms"] * 1e-3) / 1e12
        roofline = {
            "kernel": "gemm_tc_kernel<f16 split> (dX of every layer, layer-0 forward and dW, 10-wide dW)",
            "bound": "tensor",
            "achieved": ach, "peak": tensor_peak, "unit": "TFLOP/s", "frac": ach / tensor_peak,
            "traffic": traffic, "traffic_source": traffic_source,
            "peak_source": f"bf16_tflops{'_sustained' if sustained else ' (burst)'} of {peak_kind}: timed region "
                           f"{region_s * 1e3:.0f} ms at {sm_frac if sm_frac is None else round(sm_frac, 3)} of the "
                           "maximum SM clock (nvidia-smi median under load)",
            "frac_vs_burst": ach / peak_burst, "frac_vs_sustained": ach / peak_sustained,
            "executed_tflops": split["exec"] / (split["ms"] * 1e-3) / 1e12,
            "executed_frac": split["exec"] / (split["ms"] * 1e-3) / 1e12 / tensor_peak,
            "executed_frac_vs_burst": split["exec"] / (split["ms"] * 1e-3) / 1e12 / peak_burst,
            "avg_launch_ms": split["ms"] / split["n"], "launches": split["n"],
            "share_of_step": split["ms"] / ms_total,
            "timed_in": f"eager pass of the same {args.steps} steps ({ms_eager / args.steps:.3f} ms/step, "
                        "host-paced when the host is busy) with a CUDA-event pair around every GEMM "
                        "launch; share_of_step = that GEMM device time / the timed region's step time "
                        "(single-GPU runs overlap the dW GEMMs with HBM-bound kernels)",
            "note": "achieved counts the ALGORITHMIC 2MNK of the fp32 product; the kernel executes "
                    "2-3 fp16 passes per product (executed_*), so frac <= 1/2..1/3 by construction",
        }
    binary = None
    f4 = groups.get("f4", {"ms": 0.0, "n": 0, "flops": 0.0, "exec": 0.0})
    bg, bkind = (f4, "f4") if f4["n"] else (i8, "i8")
    if bg["n"]:
        tops = bg["flops"] / (bg["ms"] * 1e-3) / 1e12
        binary = {"kernel": ("gemm_tc_kernel<mxf4> (+-1 x +-1 forward of layers 1..3 on E2M1 operands, unit scale "
                             "factors, fused column sums)" if bkind == "f4" else
                             "gemm_tc_kernel<i8> (+-1 x +-1 forward, layers 1..3)"),
                  "achieved": tops, "unit": "TOP/s", "peak": int8_peak,
                  "peak_source": "cuBLASLt int8 8192^3 on this box" +
                                 (" (no library FP4 GEMM takes +-1 operands with integer outputs; the kind's MMA-only "
                                  "rate is 2x kind::i8, tools/probe_mxf4.cu)" if bkind == "f4" else ""),
                  "frac": (tops / int8_peak) if int8_peak else None,
                  "avg_launch_ms": bg["ms"] / bg["n"], "launches": bg["n"],
                  "share_of_step": bg["ms"] / ms_total}

    dw_int8 = None
    if i8dw["n"]:
        tops = i8dw["exec"] / (i8dw["ms"] * 1e-3) / 1e12
        dw_int8 = {"kernel": "gemm_tc_kernel<i8, digit split> (dW = X^T.delta' as 3 exact int8 passes, "
                             "fused SGD / reduce-scatter epilogue)",
                   "executed": tops, "unit": "TOP/s", "peak": int8_peak,
                   "frac": (tops / int8_peak) if int8_peak else None,
                   "algorithmic_tflops": i8dw["flops"] / (i8dw["ms"] * 1e-3) / 1e12,
                   "avg_launch_ms": i8dw["ms"] / i8dw["n"], "launches": i8dw["n"],
                   "share_of_step": i8dw["ms"] / ms_total}
    # the CPU baseline is a rank-0, N=1 measurement (torchrun pins every rank to one OpenMP thread)
    cpu_baseline = None
    if world == 1 and not args.no_cpu:
        cpu_val, _, _, _, cpu_sample = cpu_train_samples_per_s(steps=3, warmup=1, budget_s=40.0)
        cpu_baseline = {"value": cpu_val, "unit": UNIT, "cores": host_threads(), "kind": "port",
                        "sample": cpu_sample}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": ("f4 (E2M1 +-1, exact)" if f4["n"] else "i8") +
                 " fwd + i8x3 fixed-point dW / f16x2-split f32-acc dX", "data": "synthetic",
        "config": workload_config(world),
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(Xh.nbytes + Yh.nbytes), "d2h_bytes_per_step": 8,
                "h2d_alone_ms": ms_h2d, "h2d_alone_gb_per_s_per_gpu": Xh.nbytes / (ms_h2d * 1e-3) / 1e9,
                "api": "Network.train_stream(batches) - pinned host batches in, losses out, copies "
                       "overlapped with compute",
                "serial_train_step": {"value": world * BATCH * args.steps / (ms_serial * 1e-3),
                                      "ms_per_step": ms_serial / args.steps,
                                      "api": "Network.train_step(X_host, y_host) -> float"}},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "binary_gemm": binary,
        "dw_int8_gemm": dw_int8,
        "gemm_sites": {"kinds": site_kinds, "ms_rank0": [round(x, 4) for x in site_ms],
                       "ms_max_over_ranks": [round(max(r[i] for r in site_ranks), 4) for i in range(per_step)],
                       "sum_ms_per_rank": [round(sum(r), 4) for r in site_ranks],
                       "note": "eager pass, CUDA-event pair around every GEMM launch on the main stream, in step order "
                               "(f16: layer-0 forward, 10-wide dW, dX x3, layer-0 dW; f4 / i8: +-1 forward; i8dw: dW); "
                               "side-stream GEMMs overlap the main stream's, so the sum exceeds the step"},
        "cpu_baseline": cpu_baseline,
        "step_launch": "one CUDA graph replay per step" if use_graph else "eager kernel launches",
        "ms_per_step_eager": ms_eager / args.steps,
        "host_enqueue_ms_per_step": host_ms.get("device_resident"),
        "host_enqueue_ms_per_step_eager": host_ms.get("eager"),
        "exchanges_per_step": exchanges_per_step,
    }
    line.update(extras)
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="headline workload only (skip configs 0, 3, 4)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU-baseline legs")
    ap.add_argument("--sweep-trials", type=int, default=10000, help="fault trials of the weight sweep (config 3)")
    ap.add_argument("--image-trials", type=int, default=800, help="trials of the image bit-flip sweep")
    ap.add_argument("--profile", action="store_true",
                    help="profiling runs (ncu): skip the e2e, int8-peak and CPU-baseline legs")
    args = ap.parse_args()
    rc = run_reference(args) if args.impl == "reference" else run_gpu(args)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass
    return rc


if __name__ == "__main__":
    sys.exit(main())
