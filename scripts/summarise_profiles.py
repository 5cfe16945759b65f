#!/usr/bin/env python
"""Turn the scratch ncu outputs in gpurun_out/ into the committed summaries under profiles/.

    python scripts/summarise_profiles.py TAG launches.csv [full.ncu-rep]

writes profiles/TAG_launches.csv (the launch list, ncu banner lines dropped), profiles/TAG_launch_summary.csv
(per-kernel totals and shares) and, when a --set full report is given, profiles/TAG_ncu_full_selected.csv.
"""
import csv, io, os, re, subprocess, sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SELECT = ["launch__cluster_size", "launch__registers_per_thread", "gpu__time_duration.sum", "sm__cycles_elapsed.max",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
          "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct",
          "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
          "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum",
          # L2 / L1TEX / DRAM bytes per second (north_star: achieved L2 and HBM GB/s) and issue efficiency
          "lts__t_sectors.sum.per_second", "lts__t_sectors.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum", "dram__bytes.sum.per_second",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active",
          "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("void ", "").replace("unnamed>::", "")
    return name.strip()


def main():
    tag, launches = sys.argv[1], sys.argv[2]
    note = os.environ.get("PROFILE_NOTE", "")
    rows = [l for l in open(launches) if l.startswith('"')]
    rd = list(csv.DictReader(io.StringIO("".join(rows))))
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none launch list; {note}\n")
        f.write("id,kernel,block,grid,ns\n")
        for r in rd:
            f.write(f"{r['ID']},\"{short(r['Kernel Name'])}\",\"{r['Block Size']}\",\"{r['Grid Size']}\",{r['Metric Value']}\n")
    agg = OrderedDict()
    for r in rd:
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0])
        a[0] += 1
        a[1] += int(r["Metric Value"].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(ROOT, "profiles", f"{tag}_launch_summary.csv"), "w") as f:
        f.write(f"# per-kernel totals of profiles/{tag}_launches.csv (serialised, cold-cache replays: compare SHARES); {note}\n")
        f.write("kernel,launches,total_ns,share\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{a[0]},{a[1]},{a[1] / tot:.4f}\n")
    if len(sys.argv) > 3:
        # a .ncu-rep, or its `ncu -i REP --page raw --csv` export made on the GPU box (reports of --set full are too large to bring back)
        raw = open(sys.argv[3]).read() if sys.argv[3].endswith(".csv") else subprocess.run(["ncu", "-i", sys.argv[3], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rd = list(csv.reader(io.StringIO(raw)))
        hdr, units, body = rd[0], rd[1], rd[2:]
        idx = [hdr.index(m) for m in SELECT if m in hdr]
        base = [hdr.index(c) for c in ("Kernel Name", "Grid Size", "Block Size")]
        with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_selected.csv"), "w") as f:
            f.write(f"# ncu --set full --clock-control none --import-source on; {note}\n")
            f.write("# cold-cache, serialised replays: use for ratios/shares and DRAM traffic, not for absolute times\n")
            w = csv.writer(f)
            w.writerow([hdr[i] for i in base + idx])
            w.writerow([units[i] for i in base + idx])
            for b in body:
                w.writerow([short(b[i]) if i == base[0] else b[i] for i in base + idx])
        # DRAM bytes per launch of the dominant kernel, for bench.py's roofline.traffic (profiles/ncu_traffic.json)
        wl = os.environ.get("PROFILE_WORKLOAD")
        if wl and "dram__bytes_read.sum" in hdr and "dram__bytes_write.sum" in hdr:
            import json
            ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
            unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            vals = [(float(b[ir].replace(",", "")) * unit.get(units[ir], 1.0) + float(b[iw].replace(",", "")) * unit.get(units[iw], 1.0)) for b in body if "inner_bnb" in b[ik]]
            names = sorted({short(b[ik]) for b in body if "inner_bnb" in b[ik]})
            if vals:
                path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
                t = json.load(open(path)) if os.path.exists(path) else {}
                t[wl] = {"kernel": names[0], "dram_bytes_per_launch": sum(vals) / len(vals), "launches": len(vals),
                         "source": f"profiles/{tag}_ncu_full_selected.csv: mean dram__bytes_read.sum + dram__bytes_write.sum over the {len(vals)} {names[0]} launches of one registration (cold cache; the gathers themselves are served by L2)"}
                json.dump(t, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
