#!/usr/bin/env python
"""Times the tcgen05 GEMM at the config-2 shapes with and without the fused column reductions
(BNN_STATS_*), L2 flushed between launches.  Prints one line per case: avg ms, TFLOP/s executed."""
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "deep-neural-net_b200"))

import torch  # noqa: E402

from mlpcode import _native as nat  # noqa: E402
from mlpcode import kernels as K  # noqa: E402


def timeit(fn, reps=10):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    return sum(ts) / len(ts), min(ts)


def main():
    dev = "cuda"
    M = 8192
    one = torch.ones(1, device=dev)
    for (N, Kd, passes, label) in ((4096, 4096, ((0, 0), (0, 1), (1, 0)), "dX 8192x4096x4096 3p"),
                                   (4096, 16, ((0, 0), (0, 1), (1, 0)), "dX(L3) 8192x4096x10 3p"),
                                   (4096, 784, ((0, 0), (1, 0)), "fwd0 8192x4096x784 2p")):
        A = [torch.randn(M, Kd, device=dev).half() for _ in range(2)]
        B = [torch.randn(N, Kd, device=dev).half() for _ in range(2)]
        D = torch.empty(M, N, device=dev)
        Z = torch.randint(-4096, 4096, (M, N), device=dev, dtype=torch.int16)
        mu = torch.zeros(N, device=dev)
        red = torch.zeros(2, N, device=dev, dtype=torch.float64)
        rmax = torch.zeros(2, N, device=dev, dtype=torch.int32)
        cases = [("plain", None), ("colsum", dict(mode=nat.STATS_COLSUM, sums=red)),
                 ("bn_bwd", dict(mode=nat.STATS_BN_BWD, sums=red, max=rmax, Z=Z, mu=mu))]
        for name, st in cases:
            def run():
                K.gemm(M=M, N=N, K=Kd, A=A, B=B, lda=Kd, ldb=Kd, D=D, ldd=N, in_dtype=nat.DT_F16,
                       epilogue=nat.EPI_STORE_F32, passes=passes, sa=one, sb=one, stats=st)
            avg, best = timeit(run)
            fl = 2.0 * M * N * Kd * len(passes)
            print(f"{label:28s} {name:8s} avg {avg*1e3:8.1f} us  best {best*1e3:8.1f} us  {fl/avg/1e9:8.1f} TFLOP/s exec")
    N = Kd = 4096
    A = torch.randint(0, 2, (M, Kd), device=dev, dtype=torch.int8) * 2 - 1
    B = torch.randint(0, 2, (N, Kd), device=dev, dtype=torch.int8) * 2 - 1
    D = torch.empty(M, N, device=dev, dtype=torch.int16)
    red = torch.zeros(2, N, device=dev, dtype=torch.float64)
    for name, st in (("plain", None), ("colsum", dict(mode=nat.STATS_COLSUM, sums=red))):
        def run():
            K.gemm(M=M, N=N, K=Kd, A=(A,), B=(B,), lda=Kd, ldb=Kd, D=D, ldd=N, in_dtype=nat.DT_I8,
                   epilogue=nat.EPI_STORE_S16, stats=st)
        avg, best = timeit(run)
        print(f"{'i8 fwd 8192x4096x4096':28s} {name:8s} avg {avg*1e3:8.1f} us  best {best*1e3:8.1f} us  {2.0*M*N*Kd/avg/1e9:8.1f} TOP/s")


if __name__ == "__main__":
    main()
