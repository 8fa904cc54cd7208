"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean, share."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = row["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace("void ", "")[:48]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        agg.setdefault(k, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':50s} {'n':>5s} {'mean us':>9s} {'total us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:50s} {len(v):5d} {sum(v) / len(v):9.2f} {sum(v):10.1f} {100 * sum(v) / tot:6.1f}%")
    print(f"{'TOTAL':50s} {sum(len(v) for v in agg.values()):5d} {'':9s} {tot:10.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
