#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into the CSV/JSON files kept under profiles/.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1 [launches.csv]"""
import csv, io, json, subprocess, sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__occupancy_limit_registers", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
           "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size",
           "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "lts__t_sectors_op_red.sum",
           "lts__t_sectors_op_atom.sum"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index("Kernel Name")] + [hdr.index(m) for m in METRICS if m in hdr]
    with open(out + "_ncu_kernels.csv", "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow([hdr[i] + (f" [{units[i]}]" if units[i] else "") for i in cols])
        for r in rows[2:]:
            w.writerow([r[i] for i in cols])
    # per-kernel means
    agg = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        d = agg.setdefault(name, {"n": 0})
        d["n"] += 1
        for m in METRICS:
            if m in hdr:
                try:
                    d[m] = d.get(m, 0.0) + float(r[hdr.index(m)].replace(",", ""))
                except ValueError:
                    pass
    summ = {}
    for k, d in agg.items():
        n = d.pop("n")
        e = {m: v / n for m, v in d.items()}
        e["launches_captured"] = n
        un = {m: units[hdr.index(m)] for m in e if m in hdr}
        if "dram__bytes_read.sum" in e:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            e["dram_bytes_per_launch"] = (e["dram__bytes_read.sum"] * scale.get(un["dram__bytes_read.sum"], 1.0)
                                          + e["dram__bytes_write.sum"] * scale.get(un["dram__bytes_write.sum"], 1.0))
        e["units"] = un
        summ[k] = e
    json.dump(summ, open(out + "_ncu_summary.json", "w"), indent=1)
    if len(sys.argv) > 3:
        # launch list: share of the step per kernel
        tot = {}
        txt = open(sys.argv[3]).read()
        start = txt.index('"ID"')
        rr = list(csv.reader(io.StringIO(txt[start:])))
        h = rr[0]
        for r in rr[1:]:
            if len(r) < len(h) or not r[h.index("Metric Value")]:
                continue
            name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
            val = float(r[h.index("Metric Value")].replace(",", ""))
            unit = r[h.index("Metric Unit")]
            val *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
            t = tot.setdefault(name, [0, 0.0]); t[0] += 1; t[1] += val
        total = sum(v[1] for v in tot.values())
        with open(out + "_launch_shares.csv", "w", newline="") as fh:
            w = csv.writer(fh)
            w.writerow(["kernel", "launches", "total_us", "avg_us", "share_of_captured_time"])
            for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
                w.writerow([k, n, f"{t:.1f}", f"{t / n:.2f}", f"{t / total:.4f}"])


main()
