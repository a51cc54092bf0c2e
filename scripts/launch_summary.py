#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one step's kernels in launch order + totals.

    python scripts/launch_summary.py gpurun_out/launches.csv [first_kernel_regex]

A step is delimited by the first kernel of the step (default: the first cast_split launch after a fused backward)."""
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for x in csv.DictReader(lines):
        if x.get("Metric Name") == "gpu__time_duration.sum":
            k = re.sub(r"^void ", "", x["Kernel Name"])
            k = re.sub(r"\(.*", "", k.replace("(anonymous namespace)::", ""))
            rows.append((int(x["ID"]), k, float(x["Metric Value"].replace(",", "")) / 1000.0, x["Grid Size"], x["Block Size"]))
    return rows


def main():
    rows = load(sys.argv[1])
    # the last complete step: from the launch after the second-to-last fused backward's tail to the end of the last one
    idx = [i for i, r in enumerate(rows) if "infonce_bwd_fused_kernel" in r[1]]
    if len(idx) < 2:
        print("fewer than two steps in the list")
        return
    # a step starts at the first cast_split after the previous fused backward's trailing kernels
    starts = [i for i, r in enumerate(rows) if "cast_split" in r[1] and (i == 0 or "cast_split" not in rows[i - 1][1])]
    step_starts = []
    for b in idx:
        s = [i for i in starts if i < b]
        step_starts.append(s[-4] if len(s) >= 4 else s[0])
    a = step_starts[-2]
    b = step_starts[-1]
    step = rows[a:b]
    tot = sum(r[2] for r in step)
    print(f"| # | kernel | us | grid | block |\n|---|---|---|---|---|")
    for r in step:
        print(f"| {r[0]} | `{r[1][:80]}` | {r[2]:.1f} | {r[3]} | {r[4]} |")
    ours = sum(r[2] for r in step if "mmg::" in r[1])
    print(f"\n{len(step)} launches, {tot:.1f} us of kernel time in the step ({sum(1 for r in step if 'mmg::' in r[1])} launches / {ours:.1f} us from libmmgclip_b200.so)")
    agg = {}
    for r in step:
        agg.setdefault(r[1], [0, 0.0])
        agg[r[1]][0] += 1
        agg[r[1]][1] += r[2]
    print("\n| kernel | launches | us | share |\n|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:80]}` | {n} | {t:.1f} | {100 * t / tot:.1f} % |")


if __name__ == "__main__":
    main()
