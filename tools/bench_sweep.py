"""The workloads of BASELINE.json beside the headline training step: configs[0] (784-256-256-10 training, batch
256), configs[3] (weight bit-flip sweep, 10 000 trials x 1..1000 flipped bits on the 4096-wide BNN), the image
bit-flip sweep of imageBitflipTrainModelSel.py:19-32 and configs[4] (3072-8192-8192-10 inference, batch 65536).

Library: bench.py calls the `*_section` functions after its timed region and prints their results as extra keys
of its JSON line.  CLI (development runs, ncu captures):
    python tools/bench_sweep.py [--trials 10000] [--image-trials 800] [--cpu-trials 1] [--sections ...]
    torchrun --nproc-per-node N tools/bench_sweep.py ...       # trials / rows sharded over ranks, no data-path comm

Timing: CUDA events on the launching stream around the whole section, a barrier + synchronize on both sides,
MAX over ranks.  Trials are independent units: rank r takes the contiguous share shard_range(n, r, world) and the
only communication is the final gather of the per-trial accuracies (inside the timed region).  CPU legs: the
NumPy oracle port of the reference trial on this box's host cores, rank 0, on a bounded sample (stated)."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
for _p in (str(REPO / "deep-neural-net_b200"), str(REPO / "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

UNITS4 = [784, 4096, 4096, 4096, 10]
UNITS5 = [3072, 8192, 8192, 10]
ROWS4 = 10000


def _sum_kn(units):
    return sum(a * b for a, b in zip(units[:-1], units[1:]))


def build(units, seed=0, out_af=None, lr=0.01):
    import bnn_oracle as O
    from mlpcode.activation import ActivationFuncs as af
    from mlpcode.loss import LossFuncs as lf
    from mlpcode.network import Network
    params = O.new_params(units, np.random.default_rng(seed))
    nn = Network(units, useGpu=True, binarized=True, useBatchNorm=True)
    nn.compile(lr=lr, hiddenAf=af.sign, outAf=out_af or af.softmax, lossF=lf.cross_entropy)
    nn.load_weights([w.copy() for w in params["W"]])
    return nn, params


def mnist_like(rows=ROWS4, seed=1):
    """uint8 images, mostly background, as the MNIST test set is; labels uniform."""
    rng = np.random.default_rng(seed)
    imgs = rng.integers(0, 256, (rows, 784), dtype=np.uint8)
    imgs[rng.random((rows, 784)) < 0.8] = 0
    labels = rng.integers(0, 10, rows).astype(np.uint8)
    return imgs, labels


class DeviceTimer:
    """CUDA-event time of fn() on the current stream, barrier + synchronize on both sides, max over ranks (ms)."""

    def __init__(self, comm):
        self.comm = comm

    def fence(self):
        import torch
        torch.cuda.synchronize()
        if self.comm:
            self.comm.barrier()
            torch.cuda.synchronize()

    def __call__(self, fn):
        import torch
        self.fence()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = fn()
        e.record()
        self.fence()
        ms = torch.tensor([s.elapsed_time(e)], device="cuda", dtype=torch.float64)
        if self.comm:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms.item()), out


def call_timeline(fn, period):
    """CUDA-event time of every C-ABI call issued by fn(), summed by position within a launch group of
    `period` calls: {"<pos>:<symbol>": ms}."""
    import torch
    import mlpcode.kernels as K_
    rec, orig = [], K_._call

    def timed(name, *a):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        orig(name, *a)
        e.record()
        rec.append((name, s, e))
    K_._call = timed
    try:
        fn()
        torch.cuda.synchronize()
    finally:
        K_._call = orig
    agg = {}
    period = period or len(rec)
    for i, (name, s, e) in enumerate(rec):
        key = f"{i % period:02d}:{name}"
        agg[key] = agg.get(key, 0.0) + s.elapsed_time(e)
    return agg


# ------------------------------------------------------------------------------------------------ config 1
def config1_section(steps=50, cpu_steps=20):
    """BASELINE configs[0]: 784-256-256-10, batch 256 — the reference's own CPU-runnable case.  One GPU (a batch
    of 256 does not shard); device-resident CUDA-graph steps and, beside them, the public per-batch call with host
    arrays.  CPU: the oracle port on the host cores."""
    import torch
    import bnn_oracle as O
    import mlpcode.kernels as K_
    units, B, lr = [784, 256, 256, 10], 256, 0.01
    nn, params = build(units, seed=3, lr=lr)
    nn.enable_step_graph(True)
    rng = np.random.default_rng(4)
    X = rng.random((B, units[0]), dtype=np.float32) * 2 - 1
    Y = np.eye(10, dtype=np.int8)[rng.integers(0, 10, B)]
    Xp, Yp = torch.from_numpy(X).pin_memory(), torch.from_numpy(Y).pin_memory()
    Xd, Yd = Xp.cuda(), Yp.cuda()
    for _ in range(6):
        nn.train_step_async(Xd, Yd, lr)
    timer = DeviceTimer(None)
    l0 = K_.launch_count
    ms, _ = timer(lambda: [nn.train_step_async(Xd, Yd, lr) for _ in range(steps)])
    launches = K_.launch_count - l0
    for _ in range(3):
        nn.train_step(Xp, Yp, lr)
    ms_host, _ = timer(lambda: [nn.train_step(Xp, Yp, lr) for _ in range(steps)])
    out = {"workload": "784-256-256-10 binarized MLP training step, batch 256 (BASELINE configs[0]), one GPU",
           "train_samples_per_s": B * steps / (ms * 1e-3), "ms_per_step": ms / steps, "steps": steps,
           "gpu_launches": launches,
           "e2e": {"train_samples_per_s": B * steps / (ms_host * 1e-3), "ms_per_step": ms_host / steps,
                   "api": "Network.train_step(X_host, y_host) -> float (H2D of the batch, D2H of the loss per step)",
                   "h2d_bytes_per_step": int(X.nbytes + Y.nbytes), "d2h_bytes_per_step": 4},
           "note": "latency-bound: ~70 kernels of a few microseconds each; one CUDA-graph replay per step"}
    if cpu_steps > 0:
        ref = O.copy_params(params)
        O.train_step(ref, X.copy(), Y, lr)
        t0 = time.perf_counter()
        for _ in range(cpu_steps):
            O.train_step(ref, X.copy(), Y, lr)
        dt = (time.perf_counter() - t0) / cpu_steps
        out["cpu_baseline"] = {"value": B / dt, "unit": "samples/s", "ms_per_step": 1e3 * dt,
                               "cores": os.cpu_count(), "kind": "port",
                               "sample": f"{cpu_steps} oracle train steps at batch 256 after 1 warm-up"}
    return out


# ------------------------------------------------------------------------------------------------ config 4
def weight_sweep_section(comm, n_trials=10000, per_launch=8, cpu_trials=1, int8_peak_tops=None, timeline=False):
    """BASELINE configs[3]: 10 000 fault trials on the 784-4096^3-10 BNN, 10 000 evaluation rows fed as
    normalize(uint8 images) (weightsErrorInjectionv2.py:15), trials sharded over the ranks.
      bernoulli  every weight of every layer flips with p = 0.1 (callbacks.py:79-83 semantics)
      exact_k    trial t flips exactly 1 + (t mod 1000) distinct weights of the whole network."""
    import torch
    from mlpcode.sweep import WeightFaultSweep
    from mlpcode.utils import normalize
    import mlpcode.kernels as K_
    rank = comm.rank if comm else 0
    world = comm.world if comm else 1
    timer = DeviceTimer(comm)
    nn, params = build(UNITS4)
    imgs, labels = mnist_like()
    sweep = WeightFaultSweep(nn, normalize(imgs, newMin=-1., newMax=1.), labels, trials_per_launch=per_launch)
    ops_per_trial = 2.0 * ROWS4 * _sum_kn(UNITS4)
    out = {"workload": f"weight bit-flip sweep, {n_trials} trials x 10 000 evaluation rows on 784-4096-4096-4096-10 "
                       "(BASELINE configs[3]), trials sharded over the GPUs, no data-path communication",
           "n_gpus": world, "trials": n_trials, "trials_per_launch": per_launch,
           "layer0": "u8 x s8 on the raw bytes" if sweep.u8 is not None else "fp16 split",
           "clean_accuracy": sweep.clean_accuracy(), "ops_per_trial": ops_per_trial}
    warm = per_launch * 2 * world
    for name, fn in (("bernoulli_p0.1", lambda n: sweep.run_bernoulli(0.1, n, 7, comm)),
                     ("exact_k_1..1000", lambda n: sweep.run_exactk(lambda t: 1 + (t % 1000), n, 7, comm))):
        fn(warm)
        l0 = K_.launch_count
        ms, acc = timer(lambda: fn(n_trials))
        tps = n_trials / (ms * 1e-3)
        tops = tps * ops_per_trial / 1e12
        out[name] = {"trials_per_s": tps, "seconds": ms * 1e-3, "mean_accuracy": float(np.mean(acc)),
                     "top_per_s": tops, "gpu_launches": K_.launch_count - l0,
                     "frac_of_int8_peak": (tops / (world * int8_peak_tops)) if int8_peak_tops else None,
                     "int8_peak_source": "cuBLASLt int8 8192^3 on this box, per GPU; the +-1 layers run on kind::mxf4, "
                                         "whose MMA-only rate is 2x kind::i8 (tools/probe_mxf4.cu: 8.7 vs 4.2 POP/s)"}
    out["pipes"] = ["kind::i8 (u8 x s8)" if sweep.u8 is not None else "kind::f16"] + \
                   ["kind::mxf4 (E2M1)" if f else "kind::i8" for f in sweep.f4[1:]]
    if timeline and world == 1:   # one launch group, every C-ABI call between its own pair of CUDA events
        tl = call_timeline(lambda: sweep.run_bernoulli(0.1, per_launch, 7), None)
        out["bernoulli_ms_per_trial_by_call"] = {k: v / per_launch for k, v in tl.items()}
    del sweep
    torch.cuda.empty_cache()
    if rank == 0 and cpu_trials > 0:
        import bnn_oracle as O
        X = O.normalize(imgs, -1, 1)
        t0 = time.perf_counter()
        for t in range(cpu_trials):
            np.random.seed(t)
            hooks = [(lambda W: O.flip_bnn_mask(W, np.random.uniform(size=W.shape) < 0.1))] * 4
            O.accuracy(params, X.copy(), labels, weight_hooks=hooks)
        dt = (time.perf_counter() - t0) / cpu_trials
        out["cpu_baseline"] = {"value": 1.0 / dt, "unit": "trials/s", "seconds_per_trial": dt,
                               "cores": os.cpu_count(), "kind": "port",
                               "sample": f"{cpu_trials} Bernoulli(0.1) trial(s) of 10 000 rows (ErrorCallback + "
                                         f"get_accuracy port); {n_trials} trials would take "
                                         f"{dt * n_trials / 3600:.1f} h (linear extrapolation, not run)"}
    return out


def image_sweep_section(comm, n_trials=800, per_launch=8, cpu_trials=1, timeline=False):
    """imageBitflipTrainModelSel.py:19-32: p = 0.14 (878 of the 6272 bits of every image), 8 models x 100 trials
    in the reference = 800 trials here, on the same 4096-wide network, sharded over the ranks."""
    import torch
    from mlpcode.activation import ActivationFuncs as af
    from mlpcode.sweep import ImageFaultSweep
    import mlpcode.kernels as K_
    rank = comm.rank if comm else 0
    world = comm.world if comm else 1
    timer = DeviceTimer(comm)
    nn, params = build(UNITS4, out_af=af.identity)
    imgs, labels = mnist_like()
    isw = ImageFaultSweep(nn, imgs, labels, trials_per_launch=per_launch, batched=True)
    isw.run(0.14, per_launch * 2 * world, 11, comm=comm)
    l0 = K_.launch_count
    ms, acc = timer(lambda: isw.run(0.14, n_trials, 11, comm=comm))
    tps = n_trials / (ms * 1e-3)
    out = {"workload": f"image bit-flip sweep, p = 0.14, {n_trials} trials x 10 000 uint8 images on "
                       "784-4096-4096-4096-10 (imageBitflipTrainModelSel.py:19-32), trials sharded over the GPUs",
           "n_gpus": world, "trials": n_trials, "trials_per_s": tps, "seconds": ms * 1e-3,
           "mean_accuracy": float(np.mean(acc)), "gpu_launches": K_.launch_count - l0,
           "top_per_s": tps * 2.0 * ROWS4 * _sum_kn(UNITS4) / 1e12,
           "flip_bytes_per_trial": 2 * ROWS4 * 784}
    if timeline and world == 1:
        tl = call_timeline(lambda: isw.run(0.14, per_launch, 11), None)
        out["ms_per_trial_by_call"] = {k: v / per_launch for k, v in tl.items()}
    del isw
    torch.cuda.empty_cache()
    if rank == 0 and cpu_trials > 0:
        import bnn_oracle as O
        t0 = time.perf_counter()
        for t in range(cpu_trials):
            r2 = np.random.default_rng(t)  # callbacks.py:125-172 (row loop, choice without replacement)
            bad = np.stack([O.flip_image_row(row, r2.choice(784 * 8, 878, replace=False)) for row in imgs])
            O.accuracy(params, O.normalize(bad, -1, 1), labels, out_af="identity")
        dt = (time.perf_counter() - t0) / cpu_trials
        out["cpu_baseline"] = {"value": 1.0 / dt, "unit": "trials/s", "seconds_per_trial": dt,
                               "cores": os.cpu_count(), "kind": "port",
                               "sample": f"{cpu_trials} image trial(s), 10 000 rows (ImageErrorCallback + normalize + "
                                         "get_accuracy port)"}
    return out


# ------------------------------------------------------------------------------------------------ config 5
def inference_section(comm, batch=65536, reps=10, cpu_rows=2048, int8_peak_tops=None):
    """BASELINE configs[4]: 3072-8192-8192-10 on CIFAR-10-shaped input, batch 65536 per GPU (rows sharded, weights
    replicated, no communication).  Device-resident predict() and, beside it, predict() on pinned HOST rows."""
    import torch
    import mlpcode.kernels as K_
    rank = comm.rank if comm else 0
    world = comm.world if comm else 1
    timer = DeviceTimer(comm)
    nn, params = build(UNITS5, seed=2)
    g = torch.Generator(device="cuda").manual_seed(5 + rank)
    X = torch.rand((batch, UNITS5[0]), device="cuda", generator=g) * 2 - 1
    for _ in range(3):
        nn.predict(X)
    l0 = K_.launch_count
    ms, _ = timer(lambda: [nn.predict(X) for _ in range(reps)])
    launches = K_.launch_count - l0
    sps = world * batch * reps / (ms * 1e-3)
    ops = 2.0 * _sum_kn(UNITS5)                    # per row
    pm1_ops = 2.0 * _sum_kn(UNITS5[1:])            # of which +-1 x +-1
    out = {"workload": f"3072-8192-8192-10 BNN inference, batch {batch} per GPU (BASELINE configs[4]), rows sharded",
           "n_gpus": world, "samples_per_s": sps, "ms_per_batch": ms / reps, "reps": reps, "gpu_launches": launches,
           "flop_equiv_per_s": sps * ops, "binary_top_per_s": sps * pm1_ops / 1e12,
           "frac_of_int8_peak_binary_layers_only": None}
    # the same rows as uint8 pixels (what CIFAR-10 / MNIST files hold): normalize() keeps them as bytes and layer 0
    # becomes one exact u8 x s8 product instead of two fp16 passes (mlpcode.utils.NormalizedU8)
    from mlpcode.utils import normalize
    X8 = torch.randint(0, 256, (batch, UNITS5[0]), device="cuda", generator=g, dtype=torch.uint8)
    Xn = normalize(X8, newMin=-1., newMax=1.)
    for _ in range(2):
        nn.predict(Xn)
    ms8, _ = timer(lambda: [nn.predict(Xn) for _ in range(reps)])
    out["uint8_input"] = {"samples_per_s": world * batch * reps / (ms8 * 1e-3), "ms_per_batch": ms8 / reps,
                          "note": "same network and batch, rows given as normalize(uint8 pixels, -1, 1): layer 0 on "
                                  "kind::i8 (u8 x s8), exact"}
    del X8, Xn
    # host-fed: rows arrive from pinned host memory in chunks, class scores return to the host
    chunk = 16384
    Xh = torch.empty((batch, UNITS5[0]), dtype=torch.float32).pin_memory()
    Xh.copy_(X)
    out_h = torch.empty((batch, UNITS5[-1]), dtype=torch.float32).pin_memory()

    def host_fed():
        for k in range(0, batch, chunk):
            xb = Xh[k:k + chunk].cuda(non_blocking=True)
            out_h[k:k + chunk].copy_(nn.predict(xb).t, non_blocking=True)
    host_fed()
    ms_h, _ = timer(lambda: [host_fed() for _ in range(2)])
    out["e2e"] = {"samples_per_s": world * batch * 2 / (ms_h * 1e-3), "ms_per_batch": ms_h / 2,
                  "h2d_bytes_per_batch": int(Xh.numel() * 4), "d2h_bytes_per_batch": int(out_h.numel() * 4),
                  "api": f"Network.predict on pinned host rows in chunks of {chunk}"}
    del X, Xh, nn
    torch.cuda.empty_cache()
    if rank == 0 and cpu_rows > 0:
        import bnn_oracle as O
        Xc = np.random.default_rng(6).random((cpu_rows, UNITS5[0]), dtype=np.float32) * 2 - 1
        O.predict(params, Xc.copy())
        t0 = time.perf_counter()
        O.predict(params, Xc.copy())
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": cpu_rows / dt, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": f"one predict() of {cpu_rows} rows (of {batch}) after 1 warm-up"}
    return out


SECTIONS = ("config1", "weights", "image", "inference")


def main():
    import torch
    from mlpcode import parallel
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=10000)
    ap.add_argument("--image-trials", type=int, default=800)
    ap.add_argument("--cpu-trials", type=int, default=1)
    ap.add_argument("--per-launch", type=int, default=8)
    ap.add_argument("--sections", default="weights,image,inference",
                    help="comma-separated subset of: " + ", ".join(SECTIONS))
    args = ap.parse_args()
    sections = set(args.sections.split(","))
    comm = parallel.init_from_env("nccl")
    rank = comm.rank if comm else 0
    out = {"n_gpus": comm.world if comm else 1}
    if "config1" in sections and rank == 0:
        out["config1"] = config1_section(cpu_steps=20 if args.cpu_trials else 0)
    if "weights" in sections:
        out["weight_sweep"] = weight_sweep_section(comm, args.trials, args.per_launch, args.cpu_trials, timeline=True)
    if "image" in sections:
        out["image_sweep"] = image_sweep_section(comm, args.image_trials, args.per_launch, args.cpu_trials,
                                                 timeline=True)
    if "inference" in sections:
        out["inference"] = inference_section(comm, cpu_rows=2048 if args.cpu_trials else 0)
    if rank == 0:
        print(json.dumps(out))
    if comm:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
