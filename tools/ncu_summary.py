"""Key metrics of `ncu --set full` captures as JSON (one object per profiled launch).
usage: python tools/ncu_summary.py out.json a.ncu-rep [b.ncu-rep ...]   (reads with `ncu -i X --page raw --csv`)"""
import csv
import io
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def main():
    out, reps = sys.argv[1], sys.argv[2:]
    rows = []
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        lines = [l for l in txt.splitlines() if not l.startswith("==")]
        rd = list(csv.reader(io.StringIO("\n".join(lines))))
        if len(rd) < 3:
            continue
        hdr, units = rd[0], rd[1]
        for vals in rd[2:]:
            d = dict(zip(hdr, vals))
            u = dict(zip(hdr, units))
            o = {"report": rep.split("/")[-1], "Kernel Name": d.get("Kernel Name", "")}
            for k in KEEP:
                if k in d:
                    o[k] = d[k]
                    o.setdefault("units", {})[k] = u.get(k, "")
            rows.append(o)
    json.dump(rows, open(out, "w"), indent=1)
    for r in rows:
        print(r["Kernel Name"][:60], r.get("gpu__time_duration.sum"), r.get("units", {}).get("gpu__time_duration.sum"),
              "dram rd/wr", r.get("dram__bytes_read.sum"), r.get("dram__bytes_write.sum"), "warps active %", r.get("sm__warps_active.avg.pct_of_peak_sustained_active"))


if __name__ == "__main__":
    main()
