"""Top SASS instructions of a kernel by warp-stall samples, from `ncu -i X.ncu-rep --page source --csv`.
usage: python tools/ncu_src.py report.ncu-rep [topN] [launch_index]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    # the CSV holds one table per launch, each introduced by a "Kernel Name" line
    tables, cur = [], None
    for line in txt.splitlines():
        if line.startswith('"Kernel Name"'):
            cur = []
            tables.append(cur)
            continue
        if cur is not None:
            cur.append(line)
    rows = list(csv.DictReader(io.StringIO("\n".join(tables[which]))))
    stall_cols = [c for c in rows[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r["# Samples"] or 0) for r in rows)
    agg = {c: sum(int(r[c] or 0) for r in rows) for c in stall_cols}
    print(f"total samples {tot}; instructions {len(rows)}")
    print("stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
    ranked = sorted(enumerate(rows), key=lambda ir: -int(ir[1]["# Samples"] or 0))[:top]
    for i, r in sorted(ranked, key=lambda ir: ir[0]):
        s = int(r["# Samples"] or 0)
        why = sorted(((int(r[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
        print(f"{i:5d} {s:6d} {100 * s / max(tot, 1):5.1f}%  {r['Source'].strip()[:70]:70s} {why[0][1]}={why[0][0]} {why[1][1]}={why[1][0]}")


if __name__ == "__main__":
    main()
