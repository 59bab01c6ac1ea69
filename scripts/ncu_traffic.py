"""profiles/r02_traffic.json from an `ncu --set full` capture of the fused step (raw-page CSV):

    ncu -i gpurun_out/r02_gemm_full.ncu-rep --page raw --csv > gpurun_out/r02_gemm_full_raw.csv
    python scripts/ncu_traffic.py gpurun_out/r02_gemm_full_raw.csv "capture description" [H V]

Per GEMM kind (forward / d_hidden / d_weight) the MEDIAN over the full-chunk launches of dram__bytes_read.sum,
dram__bytes_write.sum, gpu__time_duration.sum, sm__pipe_tensor_cycles_active (% of peak), the L2 hit rate and the
registers; a compact per-launch table goes to profiles/<stem>_metrics.csv.  The JSON carries the sha256 of the kernel
sources (bench.kernel_source_hash): bench.py refuses the table when the sources changed after the capture.
"""
import csv
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

WANT = {
    "dram_read_bytes": "dram__bytes_read.sum",
    "dram_write_bytes": "dram__bytes_write.sum",
    "duration_ns": "gpu__time_duration.sum",
    "tensor_active_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "sm_clock_hz": "sm__cycles_elapsed.avg.per_second",
    "registers": "launch__registers_per_thread",
    "lts_throughput_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
}
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1.0, "us": 1e3, "ms": 1e6,
              "s": 1e9, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9, "hz": 1.0, "Khz": 1e3,
              "Mhz": 1e6, "Ghz": 1e9, "cycle/nsecond": 1e9, "cycle/usecond": 1e6, "cycle/second": 1.0}


def kind_of(name):
    if "gemm_sm100_kernel" not in name:
        return None
    if "EpiLseStoreRel" in name:
        return "forward"
    if "EpiStoreBF16" in name:
        return "d_hidden"
    if "EpiTmaF32" in name or "EpiStoreF32" in name:
        return "d_weight"
    return None


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def main():
    path, desc = sys.argv[1], sys.argv[2]
    H, V = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (3584, 152064)
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[hdr_i], rows[hdr_i + 1]
    col = {h: i for i, h in enumerate(hdr)}
    per_kind = {}
    table = []
    for r in rows[hdr_i + 2:]:
        if len(r) < len(hdr):
            continue
        k = kind_of(r[col["Kernel Name"]])
        if k is None:
            continue
        rec = {}
        for key, metric in WANT.items():
            if metric not in col:
                continue
            v = num(r[col[metric]])
            if v is not None:
                rec[key] = v * UNIT_SCALE.get(units[col[metric]], 1.0)
        per_kind.setdefault(k, []).append(rec)
        table.append((k, rec))
    out = {"capture": desc, "H": H, "V": V, "kernels": {}}
    for k, recs in per_kind.items():
        # full-chunk launches only: a partial last chunk moves fewer bytes - keep launches within 10 % of the longest
        tmax = max(r["duration_ns"] for r in recs)
        full = [r for r in recs if r["duration_ns"] > 0.9 * tmax]
        out["kernels"][k] = {key: statistics.median(r[key] for r in full if key in r) for key in WANT
                             if any(key in r for r in full)}
        out["kernels"][k]["launches_in_capture"] = len(full)
    import bench

    out["source_hash"] = bench.kernel_source_hash()
    stem = os.path.splitext(os.path.basename(path))[0].replace("_raw", "")
    with open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    with open(os.path.join(ROOT, "profiles", stem + "_metrics.csv"), "w") as f:
        keys = list(WANT)
        f.write("kind," + ",".join(keys) + "\n")
        for k, rec in table:
            f.write(k + "," + ",".join(("%g" % rec[x]) if x in rec else "" for x in keys) + "\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
