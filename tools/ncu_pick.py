#!/usr/bin/env python
"""Pick the metrics that matter out of an `ncu --page raw --csv` dump and print them as a markdown table.

    python tools/ncu_pick.py raw.csv [title] > profiles/<name>.md"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.sum",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
    "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_sectors_op_red.sum",
    "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed_op_tma_ld.sum", "smsp__sass_inst_executed_op_tmem_ldt.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    print(f"# {sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]}\n")
    print("Source: `ncu --set full --clock-control none --import-source on` on one B200 (numbers under a profiler are not bench values).\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"## `{d['Kernel Name'][:110]}`  grid {d['Grid Size']} block {d['Block Size']}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in d and d[k] not in ("", None):
                print(f"| `{k}` | {d[k]} | {units[hdr.index(k)]} |")
        print()


if __name__ == "__main__":
    main()
