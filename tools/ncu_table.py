"""Condense an `ncu --page raw --csv` export into one row per launch with the metrics the profiles/
summaries quote.  usage: python tools/ncu_table.py gpurun_out/prof_gemm_raw.csv [--json out.json]"""
import csv
import json
import sys

COLS = [
    ("us", "gpu__time_duration.sum", 1e-3),
    ("dram_rd_MB", "dram__bytes_read.sum", None),
    ("dram_wr_MB", "dram__bytes_write.sum", None),
    ("sm_GHz", "sm__cycles_elapsed.avg.per_second", None),
    ("tensor_pct", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1),
    ("tc_pipe_pct", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", 1),
    ("dram_pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("lts_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("l2_hit", "lts__t_sector_hit_rate.pct", 1),
    ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("issue_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
    ("occ_pct", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
    ("regs", "launch__registers_per_thread", 1),
    ("grid", "launch__grid_size", 1),
]
UNIT_SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "hz": 1e-9, "Khz": 1e-6, "Mhz": 1e-3,
              "Ghz": 1.0, "ns": 1.0, "us": 1e3, "ms": 1e6}


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def table(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in data:
        rec = {"kernel": r[idx["Kernel Name"]]}
        for short, col, scale in COLS:
            if col not in idx:
                continue
            v = num(r[idx[col]])
            if v is None:
                continue
            u = units[idx[col]]
            if scale is None:
                v *= UNIT_SCALE.get(u.split("/")[0], 1.0)
            elif short == "us":
                v *= UNIT_SCALE.get(u, 1.0) * 1e-3
            rec[short] = round(v, 3)
        out.append(rec)
    return out


if __name__ == "__main__":
    t = table(sys.argv[1])
    keys = ["us", "dram_rd_MB", "dram_wr_MB", "sm_GHz", "tensor_pct", "tc_pipe_pct", "dram_pct", "lts_pct",
            "l2_hit", "issue_pct", "occ_pct", "regs", "grid"]
    print("kernel".ljust(58) + "".join(k.rjust(12) for k in keys))
    for rec in t:
        print(rec["kernel"][:57].ljust(58) + "".join(str(rec.get(k, "")).rjust(12) for k in keys))
    if "--json" in sys.argv:
        json.dump(t, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
    if "--roofline" in sys.argv:
        # --roofline <kernel-name substring> <out.json>: dram bytes per launch of the kernel bench.py's roofline
        # object describes, averaged over its launches in this capture (bench.py reads the file)
        i = sys.argv.index("--roofline")
        pat, dst = sys.argv[i + 1], sys.argv[i + 2]
        rows = [r for r in t if pat in r["kernel"]]
        per = [(r["dram_rd_MB"] + r["dram_wr_MB"]) * 1e6 for r in rows]
        json.dump({"kernel": pat, "launches": len(rows), "dram_bytes_per_launch": sum(per) / len(per),
                   "dram_bytes_each": per, "us_each": [r["us"] for r in rows],
                   "source": f"ncu --set full --clock-control none, {sys.argv[1].split('/')[-1]}, "
                             f"dram__bytes_read.sum + dram__bytes_write.sum averaged over {len(rows)} launches "
                             "of one eager training step"}, open(dst, "w"), indent=1)
