#!/usr/bin/env python
"""Kernel timeline of one data-parallel training step (config 2 shapes, 8192 rows per GPU).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dp_timeline.py

Runs a few steps under torch.profiler (CUPTI) on every rank; rank 0 writes the kernels of the last
step (name, stream, start, duration) to gpurun_out/dp_timeline_n<N>.json and prints a summary: time
per kernel class, how long communication kernels ran and how much of that overlapped compute, and
the idle gaps on the compute stream.  A diagnostic, not a bench: numbers under CUPTI are inflated.
"""
import json
import os
import re
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "deep-neural-net_b200"))

import torch  # noqa: E402

import bench  # noqa: E402


def short(name):
    m = re.search(r"(\w+_kernel(<[^>]*>)?)", name)
    if m:
        return m.group(1)
    return name.split("(")[0][-60:]


def main():
    from mlpcode import parallel
    from mlpcode.activation import ActivationFuncs as af
    from mlpcode.loss import LossFuncs as lf
    from mlpcode.network import Network

    comm = parallel.init_from_env("nccl")
    world = comm.world if comm else 1
    rank = comm.rank if comm else 0
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    nn = Network(bench.UNITS, useGpu=True, binarized=True, useBatchNorm=True)
    nn.compile(lr=bench.LR, hiddenAf=af.sign, outAf=af.softmax, lossF=lf.cross_entropy)
    nn.load_weights(bench.synthetic_params())
    if comm:
        nn.set_data_parallel(comm)
    Xh, Yh = bench.synthetic_batch(100 + rank, bench.BATCH)
    Xd, Yd = torch.from_numpy(Xh).cuda(), torch.from_numpy(Yh).cuda()
    for _ in range(5):
        nn.train_step_async(Xd, Yd, bench.LR)
    torch.cuda.synchronize()
    if comm:
        comm.barrier()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            nn.train_step_async(Xd, Yd, bench.LR)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ks = sorted(((e.time_range.start, e.time_range.end - e.time_range.start, e.name,
                  getattr(e, "stream", None)) for e in evs), key=lambda k: k[0])
    # last step: from the last layer-0 binarize (first kernel of a step) onwards
    starts = [i for i, k in enumerate(ks) if "absmax" in k[2]]
    ks = ks[starts[-1] - 1:] if starts else ks
    t0 = ks[0][0]
    # every rank writes its own time line; abs_us is CUPTI's host-clock time stamp, comparable across the ranks of
    # one node (to within a few microseconds), so waits can be attributed to the rank that was late
    out = [{"name": k[2][:90], "start_us": k[0] - t0, "dur_us": k[1], "abs_us": k[0]} for k in ks]
    Path(REPO / "gpurun_out").mkdir(exist_ok=True)
    suffix = "" if rank == 0 else f"_r{rank}"
    (REPO / "gpurun_out" / f"dp_timeline_n{world}{suffix}.json").write_text(json.dumps(out, indent=0))
    if rank != 0:
        return
    comm_k = [k for k in ks if "nccl" in k[2].lower()]
    comp_k = [k for k in ks if "nccl" not in k[2].lower()]
    span = max(k[0] + k[1] for k in ks) - t0
    busy = sum(k[1] for k in comp_k)
    print(f"world {world}: step span {span:.0f} us, compute kernels {busy:.0f} us in {len(comp_k)} launches, "
          f"comm kernels {sum(k[1] for k in comm_k):.0f} us in {len(comm_k)} launches")
    agg = {}
    for k in ks:
        name = short(k[2])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += k[1]
    for name, (n, us) in sorted(agg.items(), key=lambda x: -x[1][1])[:24]:
        print(f"  {name:60s} {n:3d} {us:9.1f}")
    for k in ks:
        if any(t in k[2].lower() for t in ("nccl", "gemm_tc", "p2p", "sgd", "peer_", "memcpy")):
            print(f"    {k[0] - t0:9.1f} +{k[1]:8.1f}  {short(k[2])}")


if __name__ == "__main__":
    main()
