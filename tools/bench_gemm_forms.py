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
    # one K step per tile: the time per tile is the epilogue's (plus the fixed per-tile costs)
    Ks = 128
    cases.update({
        "epi_only_i8_store_s16": lambda: K.gemm(M=M, N=N, K=Ks, A=(A8,), B=(B8,), lda=Kd, ldb=Kd, D=Z16, ldd=N,
                                                in_dtype=nat.DT_I8, epilogue=nat.EPI_STORE_S16),
        "epi_only_i8_sign_i8": lambda: K.gemm(M=M, N=N, K=Ks, A=(A8,), B=(B8,), lda=Kd, ldb=Kd, D=S8, ldd=N,
                                              in_dtype=nat.DT_I8, epilogue=nat.EPI_SIGN_I8, epi_thr=thr, epi_sgn=sgn),
        "epi_only_i8_sign_f4": lambda: K.gemm(M=M, N=N, K=Ks, A=(A8,), B=(B8,), lda=Kd, ldb=Kd, D=S4, ldd=N,
                                              in_dtype=nat.DT_I8, epilogue=nat.EPI_SIGN_F4, epi_thr=thr, epi_sgn=sgn),
        "epi_only_f4_sign_f4": lambda: K.gemm(M=M, N=N, K=2 * Ks, A=(A4,), B=(B4,), lda=Kd, ldb=Kd, D=S4, ldd=N,
                                              in_dtype=nat.DT_F4, epilogue=nat.EPI_SIGN_F4, epi_thr=thr, epi_sgn=sgn),
        "epi_only_f4_store_s16": lambda: K.gemm(M=M, N=N, K=2 * Ks, A=(A4,), B=(B4,), lda=Kd, ldb=Kd, D=Z16, ldd=N,
                                                in_dtype=nat.DT_F4, epilogue=nat.EPI_STORE_S16),
        "epi_only_f16_store_f32": lambda: K.gemm(M=M, N=N, K=64, A=(Ah,), B=(Bh,), lda=Kd, ldb=Kd, D=F, ldd=N,
                                                 in_dtype=nat.DT_F16, epilogue=nat.EPI_STORE_F32),
        "k784_i8_sign_f4": lambda: K.gemm(M=M, N=N, K=784, A=(A8,), B=(B8,), lda=Kd, ldb=Kd, D=S4, ldd=N,
                                          in_dtype=nat.DT_I8, epilogue=nat.EPI_SIGN_F4, epi_thr=thr, epi_sgn=sgn),
        "k768_i8_sign_f4": lambda: K.gemm(M=M, N=N, K=768, A=(A8,), B=(B8,), lda=Kd, ldb=Kd, D=S4, ldd=N,
                                          in_dtype=nat.DT_I8, epilogue=nat.EPI_SIGN_F4, epi_thr=thr, epi_sgn=sgn),
    })
    # many tiles, one K step each: per-tile cost of the epilogue with the launch overhead amortised (55 tile waves)
    Mb = 65536
    Ab = torch.randint(0, 2, (Mb, 128), generator=g, device="cuda", dtype=torch.int8) * 2 - 1
    Zb = torch.zeros((Mb, N), dtype=torch.int16, device="cuda")
    Sb = torch.zeros((Mb, N), dtype=torch.int8, device="cuda")
    Fb = torch.zeros((Mb // 4, N), dtype=torch.float32, device="cuda")
    cases.update({
        "epibig_i8_store_s16": lambda: K.gemm(M=Mb, N=N, K=128, A=(Ab,), B=(B8,), lda=128, ldb=Kd, D=Zb, ldd=N,
                                              in_dtype=nat.DT_I8, epilogue=nat.EPI_STORE_S16),
        "epibig_i8_sign_i8": lambda: K.gemm(M=Mb, N=N, K=128, A=(Ab,), B=(B8,), lda=128, ldb=Kd, D=Sb, ldd=N,
                                            in_dtype=nat.DT_I8, epilogue=nat.EPI_SIGN_I8, epi_thr=thr, epi_sgn=sgn),
        "epibig_i8_sign_f4": lambda: K.gemm(M=Mb, N=N, K=128, A=(Ab,), B=(B8,), lda=128, ldb=Kd, D=Sb, ldd=2 * N,
                                            in_dtype=nat.DT_I8, epilogue=nat.EPI_SIGN_F4, epi_thr=thr, epi_sgn=sgn),
        "epibig_f16_store_f32": lambda: K.gemm(M=Mb // 4, N=N, K=64, A=(Ah,), B=(Bh,), lda=Kd, ldb=Kd, D=Fb, ldd=N,
                                               in_dtype=nat.DT_F16, epilogue=nat.EPI_STORE_F32),
    })
    out = {}
    only = os.environ.get("FORMS_CASES")
    if only:
        cases = {k: v for k, v in cases.items() if k in only.split(",")}
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
            kk = Kd
            if name.startswith("epi_only"):
                kk = 64 if "f16" in name else (2 * Ks if "_f4_s" in name and name.startswith("epi_only_f4") else Ks)
            elif name.startswith("k7"):
                kk = int(name[1:4])
            ops = 2.0 * M * N * kk * passes
            Mt = Mb if name.startswith("epibig_i8") else (Mb // 4 if name.startswith("epibig") else M)
            tiles_per_cta = -(-((Mt // 128) * (-(-N // (224 if name.split("_")[0] == "f4" or "only_f4" in name else 256)))) // 148)
            out[f"{name}/{'pair' if form == '1' else 'single'}"] = {
                "best_ms": min(times), "median_ms": sorted(times)[len(times) // 2],
                "b2b_ms": s.elapsed_time(e) / 20, "tops_best": ops / min(times) / 1e9,
                "us_per_tile_wave": 1e3 * min(times) / tiles_per_cta,
                "tops_b2b": ops / (s.elapsed_time(e) / 20) / 1e9}
    os.environ["BNN_GEMM_CTA2"] = "0"
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
