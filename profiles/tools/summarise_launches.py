"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list (markdown table on stdout).
usage: python profiles/tools/summarise_launches.py launches.csv [steps]"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*\)$", "", name)
    name = name.replace("gaussian_process_liouville_equation_b200::", "").replace("gple::", "")
    return name


def main():
    path, steps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rows = []
    with open(path) as f:
        lines = [line for line in f if line.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((short(r["Kernel Name"]), float(r["Metric Value"].replace(",", "")) / 1e3))  # ns -> us
    agg = OrderedDict()
    for k, us in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    print(f"Total profiled GPU time: {total / 1e3:.1f} ms over {len(rows)} launches ({total / 1e3 / steps:.1f} ms per step over {steps} steps)\n")
    print("| share | total ms | launches | avg us | kernel |\n|---:|---:|---:|---:|---|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {100 * us / total:.2f}% | {us / 1e3:.3f} | {n} | {us / n:.1f} | `{k}` |")


if __name__ == "__main__":
    main()
