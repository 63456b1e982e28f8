"""Single-CTA vs CTA-pair form of the tcgen05 GEMM on isolated launches at the config-2 shapes (L2 flushed between
launches).  python tools/bench_gemm_forms.py   -> one JSON line.  BNN_GEMM_CTA2 is read per call."""
import json
import os
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO / "deep-neural-net_b200"))


def main():
    import torch
    from mlpcode import _native as nat, kernels as K
    g = torch.Generator(device="cuda").manual_seed(0)
    M, N, Kd = 8192, 4096, 4096
    A8 = torch.randint(0, 2, (M, Kd), generator=g, device="cuda", dtype=torch.int8) * 2 - 1
    B8 = torch.randint(0, 2, (N, Kd), generator=g, device="cuda", dtype=torch.int8) * 2 - 1
    Ah, Al = (torch.randn((M, Kd), generator=g, device="cuda").half() for _ in range(2))
    Bh, Bl = (torch.randn((N, Kd), generator=g, device="cuda").half() for _ in range(2))
    Z16 = torch.zeros((M, N), dtype=torch.int16, device="cuda")
    S8 = torch.zeros((M, N), dtype=torch.int8, device="cuda")
    F = torch.zeros((M, N), dtype=torch.float32, device="cuda")
    thr = torch.zeros(N, device="cuda")
    sgn = torch.ones(N, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    def f4(x):   # +-1 int8 [R, K] -> E2M1 nibbles [R, K / 2]
        nib = torch.where(x > 0, 0x2, 0xA).to(torch.uint8)
        return (nib[:, 0::2] | (nib[:, 1::2] << 4)).contiguous()
    A4, B4 = f4(A8), f4(B8)
    S4 = torch.zeros((M, N // 2), dtype=torch.uint8, device="cuda")
    cases = {
        "f4_sign_f4": lambda: K.gemm(M=M, N=N, K=Kd, A=(A4,), B=(B4,), lda=Kd, ldb=Kd, D=S4, ldd=N,
                                     in_dtype=nat.DT_F4, epilogue=nat.EPI_SIGN_F4, epi_thr=thr, epi_sgn=sgn),
        "f4_store_s16": lambda: K.gemm(M=M, N=N, K=Kd, A=(A4,), B=(B4,), lda=Kd, ldb=Kd, D=Z16, ldd=N,
                                       in_dtype=nat.DT_F4, epilogue=nat.EPI_STORE_S16),
        "i8_store_s16": lambda: K.gemm(M=M, N=N, K=Kd, A=(A8,), B=(B8,), lda=Kd, ldb=Kd, D=Z16, ldd=N,
                                       in_dtype=nat.DT_I8, epilogue=nat.EPI_STORE_S16),
        "i8_sign_i8": lambda: K.gemm(M=M, N=N, K=Kd, A=(A8,), B=(B8,), lda=Kd, ldb=Kd, D=S8, ldd=N,
                                     in_dtype=nat.DT_I8, epilogue=nat.EPI_SIGN_I8, epi_thr=thr, epi_sgn=sgn),
        "f16_1pass_f32": lambda: K.gemm(M=M, N=N, K=Kd, A=(Ah,), B=(Bh,), lda=Kd, ldb=Kd, D=F, ldd=N,
                                        in_dtype=nat.DT_F16, epilogue=nat.EPI_STORE_F32),
        "f16_3pass_f32": lambda: K.gemm(M=M, N=N, K=Kd, A=(Ah, Al), B=(Bh, Bl), lda=Kd, ldb=Kd, D=F, ldd=N,
                                        in_dtype=nat.DT_F16, epilogue=nat.EPI_STORE_F32,
                                        passes=((0, 0), (0, 1), (1, 0))),
    }
    out = {}
    for name, fn in cases.items():
        for form in (("0",) if (name.startswith("f4") or os.environ.get("FORMS_SINGLE_ONLY")) else ("0", "1")):
            os.environ["BNN_GEMM_CTA2"] = form
            times = []
            for i in range(13):
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                fn()
                e.record()
                e.synchronize()
                if i >= 3:
                    times.append(s.elapsed_time(e))
            # back-to-back (sustained): 20 launches, one event pair
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(20):
                fn()
            e.record()
            e.synchronize()
            passes = 3 if "3pass" in name else 1
            ops = 2.0 * M * N * Kd * passes
            out[f"{name}/{'pair' if form == '1' else 'single'}"] = {
                "best_ms": min(times), "median_ms": sorted(times)[len(times) // 2],
                "b2b_ms": s.elapsed_time(e) / 20, "tops_best": ops / min(times) / 1e9,
                "tops_b2b": ops / (s.elapsed_time(e) / 20) / 1e9}
    os.environ["BNN_GEMM_CTA2"] = "0"
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
