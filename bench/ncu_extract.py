#!/usr/bin/env python3
"""Condense an ncu report for profiles/: the raw page as CSV (one launch) plus a short key-metric text.

    python bench/ncu_extract.py gpurun_out/x.ncu-rep profiles/r2_<name> [launch]   -> <name>_full_raw.csv, <name>_key_metrics.txt,
                                                                                     <name>_top_source_lines.txt
`launch`: index of the launch inside the report (default 0)."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__sass_inst_executed_op_tma_ld.sum", "smsp__sass_inst_executed_op_global_ld.sum",
        "smsp__sass_inst_executed_op_global_st.sum", "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_local_ld.sum"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2 + which]
    with open(out + "_full_raw.csv", "w") as f:      # header, units and the selected launch only
        w = csv.writer(f)
        w.writerows([hdr, units, vals])
    with open(out + "_key_metrics.txt", "w") as f:
        f.write("kernel: %s\n" % vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write("%-70s %s %s\n" % (k, vals[i], units[i]))
        f.write("\nwarp stall reasons (cycles per issued instruction, > 0.2):\n")
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and "per_issue_active" in h:
                try:
                    v = float(vals[i])
                except ValueError:
                    continue
                if v > 0.2:
                    f.write("  %-28s %.2f\n" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
    launch_id = vals[hdr.index("ID")] if "ID" in hdr else str(which)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + (["--launch-skip", str(which), "--launch-count", "1"] if which else []),
                         capture_output=True, text=True).stdout
    tmp = out + "_src.tmp.csv"
    open(tmp, "w").write(src)
    top = subprocess.run([sys.executable, __file__.replace("ncu_extract.py", "ncu_top_lines.py"), tmp, "40"], capture_output=True, text=True).stdout
    open(out + "_top_source_lines.txt", "w").write(top)
    import os
    os.unlink(tmp)
    print(open(out + "_key_metrics.txt").read())


if __name__ == "__main__":
    main()
