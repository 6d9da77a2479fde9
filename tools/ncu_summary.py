"""Summarise a .ncu-rep (raw page) into a small CSV: one row per captured launch with the
metrics the roofline discussion uses.   python tools/ncu_summary.py REPORT.ncu-rep OUT.csv"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio",
        "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg", "smsp__cycles_active.avg"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    cols = [(w, hdr.index(w)) for w in WANT if w in hdr]
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow([c for c, _ in cols])
        w.writerow([units[i] for _, i in cols])
        for r in body:
            w.writerow([r[i].split("(")[0] if c == "Kernel Name" else r[i] for c, i in cols])
    print("wrote", out, len(body), "launches")


if __name__ == "__main__":
    main()
