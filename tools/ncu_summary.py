#!/usr/bin/env python
"""Summarises an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the few numbers the
design notes quote: duration, DRAM bytes, issue rate, top stall reasons per kernel."""
import csv
import subprocess
import sys


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
            "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
            "sm__cycles_elapsed.avg", "sm__cycles_active.avg"]
    for r in rows[2:]:
        print("== %s" % r[ix["Kernel Name"]])
        for k in keys:
            if k in ix:
                print("   %-70s %s %s" % (k, r[ix[k]], units[ix[k]]))
        st = [(h, float(r[i] or 0)) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("per_issue_active.ratio")]
        st.sort(key=lambda x: -x[1])
        print("   stalls (warps per issue-active cycle): " + ", ".join("%s=%.2f" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v) for h, v in st[:6]))


if __name__ == "__main__":
    main(sys.argv[1])
