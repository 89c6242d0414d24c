"""Prints the handful of ncu metrics the roofline discussion needs from a .ncu-rep (run where ncu is installed)."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__inst_executed.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("== %s" % r[idx["Kernel Name"]][:90])
        for k in KEYS:
            if k in idx:
                print("   %-82s %s %s" % (k, r[idx[k]], units[idx[k]]))
        st = [(h, float(r[idx[h]])) for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and r[idx[h]]]
        for h, v in sorted(st, key=lambda x: -x[1])[:5]:
            print("   stall %-60s %.3f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))


if __name__ == "__main__":
    main(sys.argv[1])
