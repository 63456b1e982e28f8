"""Fault-sweep and inference throughput (BASELINE configs[3] and [4]) on this rank's GPU(s).
    python tools/bench_sweep.py [--trials 64] [--cpu-trials 1]
    torchrun --nproc-per-node N tools/bench_sweep.py ...       # trials sharded over ranks, no comm
Prints one JSON line on rank 0: weight-fault trials/s (Bernoulli p=0.1 and exact-k 1..1000 on the
784-4096^3-10 BNN, 10 000 evaluation rows), inference samples/s (3072-8192-8192-10, batch 65536), and the
NumPy oracle port of one reference trial timed on the host cores."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "deep-neural-net_b200"), str(REPO / "oracle")]


def build(units, seed=0):
    import bnn_oracle as O
    from mlpcode.activation import ActivationFuncs as af
    from mlpcode.network import Network
    params = O.new_params(units, np.random.default_rng(seed))
    nn = Network(units, useGpu=True, binarized=True, useBatchNorm=True)
    nn.compile(lr=0.01, hiddenAf=af.sign, outAf=af.softmax)
    nn.load_weights(params["W"])
    return nn, params


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
    for i, (name, s, e) in enumerate(rec):
        key = f"{i % period}:{name}"
        agg[key] = agg.get(key, 0.0) + s.elapsed_time(e)
    return agg


def main():
    import torch
    from mlpcode import parallel
    from mlpcode.sweep import WeightFaultSweep
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=64)
    ap.add_argument("--cpu-trials", type=int, default=1)
    ap.add_argument("--per-launch", type=int, default=8)
    ap.add_argument("--image-trials", type=int, default=32)
    ap.add_argument("--sections", default="weights,image,inference",
                    help="comma-separated subset of: weights, image, inference")
    args = ap.parse_args()
    sections = set(args.sections.split(","))
    comm = parallel.init_from_env("nccl")
    rank = comm.rank if comm else 0
    world = comm.world if comm else 1

    def sync():
        torch.cuda.synchronize()
        if comm:
            comm.barrier()
            torch.cuda.synchronize()

    out = {"n_gpus": world}
    # ---- config 4: weight-fault sweep -----------------------------------------------------------
    units = [784, 4096, 4096, 4096, 10]
    nn, params = build(units)
    rng = np.random.default_rng(1)
    X = rng.random((10000, 784), dtype=np.float32) * 2 - 1
    labels = rng.integers(0, 10, 10000).astype(np.uint8)
    imgs = rng.integers(0, 256, (10000, 784), dtype=np.uint8)
    imgs[rng.random((10000, 784)) < 0.8] = 0                      # MNIST-like: mostly background
    sweep = WeightFaultSweep(nn, X, labels, trials_per_launch=args.per_launch)
    out["clean_accuracy"] = sweep.clean_accuracy()
    n_trials = args.trials * world
    for name, fn in (("bernoulli_p0.1", lambda: sweep.run_bernoulli(0.1, n_trials, 7, comm)),
                     ("exactk_1..1000", lambda: sweep.run_exactk(lambda t: 1 + (t % 1000), n_trials, 7, comm))):
        if "weights" not in sections:
            break
        fn()                                                  # warm-up (allocations, first launches)
        sync()
        t0 = time.perf_counter()
        acc = fn()
        sync()
        dt = time.perf_counter() - t0
        out[name] = {"trials": n_trials, "seconds": dt, "trials_per_s": n_trials / dt,
                     "mean_accuracy": float(np.mean(acc)),
                     "flop_equiv_per_trial": 2.0 * 10000 * sum(a * b for a, b in zip(units[:-1], units[1:]))}
    if "weights" in sections:
        # the same sweep fed `normalize(uint8 test set)` as weightsErrorInjectionv2.py:15 does: layer 0 on bytes
        from mlpcode.utils import normalize
        sweep8 = WeightFaultSweep(nn, normalize(imgs, newMin=-1., newMax=1.), labels,
                                  trials_per_launch=args.per_launch)
        sweep8.run_bernoulli(0.1, n_trials, 7, comm)
        sync()
        t0 = time.perf_counter()
        acc = sweep8.run_bernoulli(0.1, n_trials, 7, comm)
        sync()
        dt = time.perf_counter() - t0
        out["bernoulli_p0.1_uint8_images"] = {"trials": n_trials, "seconds": dt, "trials_per_s": n_trials / dt,
                                              "mean_accuracy": float(np.mean(acc)), "layer0": "u8 x s8" if sweep8.u8 is not None else "fp16 split"}
        del sweep8
    # ---- image bit-flip sweep (imageBitflipTrainModelSel.py:19-32) on the same network ----------
    if args.image_trials > 0 and "image" in sections:
        from mlpcode.sweep import ImageFaultSweep
        n_img = args.image_trials * world
        for name, batched in (("image_p0.14_u8_batched", True), ("image_p0.14_per_trial_fp32", False)):
            isw = ImageFaultSweep(nn, imgs, labels, trials_per_launch=args.per_launch, batched=batched)
            isw.run(0.14, n_img, 11, comm=comm)
            sync()
            t0 = time.perf_counter()
            acc = isw.run(0.14, n_img, 11, comm=comm)
            sync()
            dt = time.perf_counter() - t0
            if batched and world == 1:           # a collective-free run only (the sweep gathers at its end)
                tl = call_timeline(lambda: isw.run(0.14, n_img, 11, comm=comm), 7)
                out["image_u8_batched_ms_per_trial_by_call"] = {k: v / n_img for k, v in tl.items()}
            out[name] = {"trials": n_img, "seconds": dt, "trials_per_s": n_img / dt,
                         "mean_accuracy": float(np.mean(acc)),
                         # flips: 7.84 MB read + written per trial; layer 0 reads the bytes once more
                         "image_bytes_per_trial": 3 * 10000 * 784}
            del isw
        if rank == 0 and args.cpu_trials > 0:
            import bnn_oracle as O
            t0 = time.perf_counter()
            for t in range(args.cpu_trials):
                # callbacks.py:125-172 (row loop, choice without replacement) + normalize + get_accuracy
                r2 = np.random.default_rng(t)
                bad = np.stack([O.flip_image_row(row, r2.choice(784 * 8, 878, replace=False)) for row in imgs])
                O.accuracy(params, O.normalize(bad, -1, 1), labels, out_af="identity")
            dt = (time.perf_counter() - t0) / args.cpu_trials
            out["cpu_reference_port_image"] = {"seconds_per_trial": dt, "trials_per_s": 1.0 / dt,
                                               "sample": f"{args.cpu_trials} image trial(s), p = 0.14, 10 000 rows"}
    # ---- config 5: large-batch inference --------------------------------------------------------
    if "inference" in sections:
        del sweep, nn
        torch.cuda.empty_cache()
        units5 = [3072, 8192, 8192, 10]
        nn5, _ = build(units5, seed=2)
        B5 = 65536
        X5 = torch.rand((B5, 3072), device="cuda") * 2 - 1
        for _ in range(2):
            nn5.predict(X5)
        sync()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            nn5.predict(X5)
        sync()
        dt = time.perf_counter() - t0
        out["inference_3072-8192-8192-10"] = {"batch_per_gpu": B5, "samples_per_s": world * B5 * reps / dt,
                                              "ms_per_batch": 1e3 * dt / reps}
    # ---- CPU baseline: the oracle port of one reference trial (callbacks.py:79-83 + get_accuracy) --
    if rank == 0 and args.cpu_trials > 0 and "weights" in sections:
        import bnn_oracle as O
        t0 = time.perf_counter()
        for t in range(args.cpu_trials):
            np.random.seed(t)
            hooks = [(lambda W: O.flip_bnn_mask(W, np.random.uniform(size=W.shape) < 0.1))] * 4
            O.accuracy(params, X.copy(), labels, weight_hooks=hooks)
        dt = (time.perf_counter() - t0) / args.cpu_trials
        out["cpu_reference_port"] = {"seconds_per_trial": dt, "trials_per_s": 1.0 / dt,
                                     "sample": f"{args.cpu_trials} Bernoulli(0.1) trial(s), 10 000 rows"}
    if rank == 0:
        print(json.dumps(out))
    if comm:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
