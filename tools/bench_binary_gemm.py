"""Times the two variants of the +-1 x +-1 GEMM (north star: pick the winner from ncu counters) at the
config-2 layer shape 8192 x 4096 x 4096, plus cuBLASLt int8 as the int8 tensor-pipe yardstick.
    python tools/bench_binary_gemm.py [--reps 20]       # prints one JSON line
Run it under `ncu --set full -k regex:gemm_(tc|popc)_kernel` for the pipe-utilisation counters."""
import argparse
import json
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "deep-neural-net_b200"))


def main():
    import torch
    from mlpcode import _native as nat, kernels as K
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    M, N, Kd = 8192, 4096, 4096
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randint(0, 2, (M, Kd), generator=g, device="cuda", dtype=torch.int8) * 2 - 1
    B = torch.randint(0, 2, (N, Kd), generator=g, device="cuda", dtype=torch.int8) * 2 - 1
    ab = torch.zeros((M, Kd // 32), dtype=torch.int32, device="cuda")
    bb = torch.zeros((N, Kd // 32), dtype=torch.int32, device="cuda")
    K.pack_i8_rows(A, Kd, ab)
    K.pack_i8_rows(B, Kd, bb)
    Z1 = torch.zeros((M, N), dtype=torch.int16, device="cuda")
    Z2 = torch.zeros((M, N), dtype=torch.int16, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed(fn):
        best, tot = 1e9, 0.0
        for i in range(args.reps + 3):
            flush.zero_()                      # evict L2 between timed launches
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            e.synchronize()
            if i >= 3:
                ms = s.elapsed_time(e)
                best, tot = min(best, ms), tot + ms
        return best, tot / args.reps

    def run_i8():
        K.gemm(M=M, N=N, K=Kd, A=(A,), B=(B,), lda=Kd, ldb=Kd, D=Z1, ldd=N, in_dtype=nat.DT_I8,
               epilogue=nat.EPI_STORE_S16, impl=nat.GEMM_TCGEN05)

    def run_popc():
        K.gemm_popc(ab, bb, M, N, Kd, Z2, out_s16=True)

    ops = 2.0 * M * N * Kd
    out = {"shape": [M, N, Kd], "ops": ops}
    for name, fn in (("tcgen05_i8", run_i8), ("xor_popc", run_popc),
                     ("cublaslt_int8", lambda: torch._int_mm(A, B.t()))):
        best, avg = timed(fn)
        out[name] = {"best_ms": best, "avg_ms": avg, "tops_best": ops / best / 1e9,
                     "tops_avg": ops / avg / 1e9}
    out["bit_identical"] = bool(torch.equal(Z1, Z2))
    out["winner"] = "tcgen05_i8" if out["tcgen05_i8"]["avg_ms"] < out["xor_popc"]["avg_ms"] else "xor_popc"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
