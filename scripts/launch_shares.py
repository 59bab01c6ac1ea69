"""Per-kernel shares of ONE fused step from an ncu launch list (`--metrics gpu__time_duration.sum --csv`):

    python scripts/launch_shares.py gpurun_out/r02h_launches_c2.csv profiles/r02h_launches_c2_step.csv

A step = the launches from a `loss_rowstats_kernel` (the mask pre-pass that opens skyb200_policy_step) to the next
`loss_finalize_kernel`; the LAST complete step of the capture is cut out, written as a compact CSV (kernel, grid,
block, ns) and summarised as a markdown table (launches, serialised time, share, average).
"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("skyb200::", "")
    m = re.match(r"(void )?(gemm_sm100_kernel)<.*?(Epi\w+)>", name)
    if m:
        return f"gemm_sm100_kernel<{m.group(3)}>"
    return re.sub(r"^void ", "", re.sub(r"<.*$", "", name))


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    col = {h: i for i, h in enumerate(rows[hi])}
    data = [(short(r[col["Kernel Name"]]), r[col["Grid Size"]], r[col["Block Size"]], float(r[col["Metric Value"]]))
            for r in rows[hi + 1:] if len(r) == len(rows[hi]) and r[col["Metric Name"]] == "gpu__time_duration.sum"]
    ends = [i for i, d in enumerate(data) if d[0] == "loss_finalize_kernel"]
    starts = [i for i, d in enumerate(data) if d[0] == "loss_rowstats_kernel"]
    end = ends[-1]
    prev_end = ends[-2] if len(ends) > 1 else -1
    start = min(i for i in starts if prev_end < i < end)
    step = data[start:end + 1]
    with open(dst, "w") as f:
        f.write("kernel,grid,block,ns\n")
        for k, g, b, ns in step:
            f.write(f"{k},\"{g}\",\"{b}\",{ns:.0f}\n")
    agg = collections.OrderedDict()
    for k, _, _, ns in step:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    print(f"{len(step)} launches, {total / 1e6:.1f} ms serialised")
    print("| kernel | launches | serialised time | share | avg |")
    print("|---|---|---|---|---|")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {ns / 1e6:.2f} ms | {100 * ns / total:.1f} % | {ns / n / 1e3:.1f} us |")


if __name__ == "__main__":
    main()
