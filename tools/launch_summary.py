"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import sys


def main(path, title=""):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        tot[k] += v
        cnt[k] += 1
    T = sum(tot.values())
    if title:
        print(title + "\n")
    print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"| {k} | {cnt[k]} | {v:.1f} | {v / cnt[k]:.1f} | {100 * v / T:.1f}% |")
    print(f"\ntotal {T:.1f} us over {sum(cnt.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
