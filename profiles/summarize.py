#!/usr/bin/env python3
"""Turn an .ncu-rep (captured on the B200 box with `ncu --set full --clock-control none
--import-source on`) into the small per-launch metric table kept under profiles/.

    python profiles/summarize.py gpurun_out/prof.ncu-rep > profiles/rNN_<kernel>.csv
"""
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    w = csv.writer(sys.stdout)
    w.writerow(["launch", "kernel", "metric", "unit", "value"])
    for li, r in enumerate(rows[2:]):
        kname = r[name_i].split("(")[0]
        for h, u, v in zip(hdr, units, r):
            if h in KEEP:
                w.writerow([li, kname, h, u, v])


if __name__ == "__main__":
    main(sys.argv[1])
