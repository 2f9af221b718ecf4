"""Summarise `ncu --set full` reports (one row per kernel launch) into a small markdown table for profiles/.
    python scripts/ncu_summary.py gpurun_out/prof_volume.ncu-rep [more.ncu-rep ...] > profiles/rNN_xxx.md
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX % of peak"),
    ("l1tex__t_sector_hit_rate.pct", "L1 sector hit rate"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data-pipe (LSU) wavefronts % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts.sum", "L1 data-pipe (LSU) wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "  of which shared-memory wavefronts"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests (L1 tag stage)"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("smsp__inst_executed_pipe_lsu.sum", "LSU instructions"),
    ("sm__inst_executed_pipe_uniform.sum", "uniform-pipe instructions"),
    ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "stall: long scoreboard (global/L2 latency) %"),
    ("smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "stall: short scoreboard (smem/MUFU) %"),
    ("smsp__warp_issue_stalled_barrier_per_warp_active.pct", "stall: barrier %"),
    ("smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "stall: LSU queue throttle %"),
    ("smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "stall: MIO throttle %"),
    ("smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "stall: math pipe throttle %"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor (hmma) pipe active %"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active (realtime) %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
]


def main():
    for path in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        print(f"## `{path.split('/')[-1]}`\n")
        for r in rows[2:]:
            name = r[col["Kernel Name"]].split("(")[0]
            print(f"### `{name}`\n")
            print("| metric | value |")
            print("|---|---|")
            for key, label in METRICS:
                hit = [h for h in hdr if h == key or h.endswith("." + key)]
                if not hit or r[col[hit[0]]] == "":
                    continue
                i = col[hit[0]]
                print(f"| {label} (`{key}`) | {r[i]} {units[i]} |")
            print()


if __name__ == "__main__":
    main()
