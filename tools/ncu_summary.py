"""Compact summary of an .ncu-rep (read here on the CPU box): one line per metric we track.
    python tools/ncu_summary.py gpurun_out/r01_prof_dec2.ncu-rep [more.ncu-rep ...]"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.avg.per_second",
    "launch__registers_per_thread",
    "launch__grid_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
]


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


for path in sys.argv[1:]:
    hdr, units, data = rows_of(path)
    idx = {h: i for i, h in enumerate(hdr)}
    for r in data:
        print(f"## {path}: {r[idx['Kernel Name']]}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
        for w in WANT:
            if w in idx:
                print(f"  {w:86s} {r[idx[w]]:>18s} {units[idx[w]]}")
