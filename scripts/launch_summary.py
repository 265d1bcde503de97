"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: python scripts/launch_summary.py launches.csv "header line" > summary.txt"""
import csv
import sys
from collections import defaultdict


def main():
    path, head = sys.argv[1], sys.argv[2]
    tot, cnt = defaultdict(float), defaultdict(int)
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        for cut in ("(", "<7", "<3"):
            name = name.split(cut)[0] if cut == "(" else name
        name = name.replace("void ", "")
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v_us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        tot[name] += v_us
        cnt[name] += 1
    s = sum(tot.values())
    print(head)
    print("(cold-cache, serialised launch times: compare shares, not absolutes)\n")
    for k in sorted(tot, key=lambda k: -tot[k]):
        print(f"{k[:58]:58s} launches {cnt[k]:5d}  total {tot[k]:10.1f} us  share {100 * tot[k] / s:5.1f}%")


if __name__ == "__main__":
    main()
