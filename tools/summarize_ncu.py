#!/usr/bin/env python
"""Turn gpurun_out/<tag>_prof_{f32,bf16}.ncu-rep + <tag>_launches.csv into the tracked summaries under profiles/:
profiles/<tag>_ncu_summary.md, profiles/<tag>_launches.csv (our kernels only) and profiles/traffic.json (dram bytes per
launch of the dominant kernel, read by bench.py).  Usage: python tools/summarize_ncu.py <tag>"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %peak"),
    ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "L1 LSU wavefronts %peak (1 wavefront/clk/SM)"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "  of which shared memory"),
    ("l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed", "L1->XBAR request port busy %"),
    ("l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum", "RED sectors L1->L2"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %peak"),
    ("lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed", "L2 atomic unit busy %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue slots busy %"),
    ("sm__inst_executed.sum", "warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "CTAs/SM (register limit)"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected / issue"),
]


def raw(rep):
    """Rows of a report: from the .ncu-rep if it came back, else from the `--page raw --csv` export made on the GPU box."""
    if os.path.exists(rep):
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        out = open(rep.replace(".ncu-rep", "_raw.csv")).read()
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def to_bytes(val, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(val) * mult[unit]


def main():
    tag = sys.argv[1]
    out_md = [f"# ncu summary `{tag}` (B200, `ncu --set full --clock-control none`, BASELINE config 2: N=64, 64x64x256, G=8, dist T)\n",
              "Raw reports: `gpurun_out/%s_prof_{f32,bf16}.ncu-rep` (scratch, not tracked).  Numbers under a profiler are "
              "never bench values; kernel times here are cold-cache and serialised.\n" % tag]
    traffic = {}
    for dt in ("f32", "bf16"):
        rep = os.path.join(ROOT, "gpurun_out", f"{tag}_prof_{dt}.ncu-rep")
        if not os.path.exists(rep) and not os.path.exists(rep.replace(".ncu-rep", "_raw.csv")):
            continue
        rows, units = raw(rep)
        out_md.append(f"\n## {dt}\n")
        names = [r["Kernel Name"].split("(")[0].replace("void gp::", "") for r in rows]
        out_md.append("| metric | " + " | ".join(names) + " |")
        out_md.append("|---|" + "---|" * len(rows))
        for key, label in KEYS:
            if key not in rows[0]:
                continue
            vals = []
            for r in rows:
                v = r[key]
                try:
                    v = f"{float(v):,.4g}" if abs(float(v)) < 1e6 else f"{float(v):,.0f}"
                except ValueError:
                    pass
                vals.append(f"{v} {units[key]}".strip())
            out_md.append(f"| {label} (`{key}`) | " + " | ".join(vals) + " |")
        for r in rows:
            kind = "fwd" if "fwd" in r["Kernel Name"] else "bwd" if "bwd" in r["Kernel Name"] else None
            if kind:
                traffic[f"dcnv3_{kind}_{dt}_dram_bytes"] = int(to_bytes(r["dram__bytes_read.sum"], units["dram__bytes_read.sum"]) +
                                                              to_bytes(r["dram__bytes_write.sum"], units["dram__bytes_write.sum"]))
    # other captures of the round: the tcgen05 dense layer (tensor-pipe evidence) and the kernels added this round
    TC_KEYS = [("gpu__time_duration.sum", "duration"),
               ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (of active cycles)"),
               ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (of elapsed cycles)"),
               ("sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.sum", "tcgen05 (UTCHMMA) bf16->fp32 ops = 2*M*N*K"),
               ("sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.sum", "legacy mma.sync (HMMA) ops"),
               ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor memory (TMEM) active %"),
               ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
               ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %peak"),
               ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %peak"),
               ("launch__registers_per_thread", "registers/thread"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %")]
    for name, title, keys in (("tcl_mem", "tcgen05 dense layer, M=262144 N=256 K=256 (HBM-bound projection shape)", TC_KEYS),
                              ("tcl_fc1", "tcgen05 dense layer, M=4096 N=2048 K=8192 (fc1||fc1_z, tensor-bound)", TC_KEYS),
                              ("new", "kernels added in this round (tools/profile_target_r06.py)", KEYS)):
        rep = os.path.join(ROOT, "gpurun_out", f"{tag}_prof_{name}.ncu-rep")
        if not os.path.exists(rep) and not os.path.exists(rep.replace(".ncu-rep", "_raw.csv")):
            continue
        rows, units = raw(rep)
        # keep the LAST launch of every distinct kernel (the first pays cold caches)
        last = {}
        for r in rows:
            last[r["Kernel Name"].split("(")[0]] = r
        rows = list(last.values())
        out_md.append(f"\n## {title}\n")
        names = [r["Kernel Name"].split("(")[0].replace("void gp::", "") for r in rows]
        out_md.append("| metric | " + " | ".join(names) + " |")
        out_md.append("|---|" + "---|" * len(rows))
        tensor_keys = [k for k in rows[0] if "tensor" in k and ("pct_of_peak_sustained_active" in k or k.endswith(".sum"))][:0]
        for key, label in keys:
            if key not in rows[0]:
                continue
            vals = []
            for r in rows:
                v = r[key]
                try:
                    v = f"{float(v.replace(',', '')):,.4g}" if abs(float(v.replace(',', ''))) < 1e6 else f"{float(v.replace(',', '')):,.0f}"
                except ValueError:
                    pass
                vals.append(f"{v} {units[key]}".strip())
            out_md.append(f"| {label} (`{key}`) | " + " | ".join(vals) + " |")
    # launch list of the bench command
    lp = os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv")
    if os.path.exists(lp):
        lines = [l for l in open(lp) if l.startswith('"')]
        rows = list(csv.DictReader(io.StringIO("".join(lines))))
        ours = [r for r in rows if "gp::" in r["Kernel Name"]]
        with open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv"), "w") as f:
            f.write("id,kernel,grid,block,duration_ns\n")
            for r in ours:
                f.write(f'{r["ID"]},"{r["Kernel Name"].split("(")[0]}","{r["Grid Size"]}","{r["Block Size"]}",{r["Metric Value"]}\n')
        agg = {}
        for r in ours:
            k = r["Kernel Name"].split("<")[0].replace("void gp::", "")
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += float(r["Metric Value"].replace(",", ""))
        tot = sum(a[1] for a in agg.values())
        out_md.append(f"\n## launch list of `bench.py --steps 2 --warmup 3` ({len(ours)} launches of our kernels; profiles/{tag}_launches.csv)\n")
        out_md.append("| kernel | launches | avg duration (us) | share of our kernel time |")
        out_md.append("|---|---|---|---|")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out_md.append(f"| {k} | {n} | {t / n / 1e3:.1f} | {100 * t / tot:.1f} % |")
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.md"), "w").write("\n".join(out_md) + "\n")
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    old = json.load(open(tp)) if os.path.exists(tp) else {}
    old.update(traffic)
    old["source"] = f"profiles/{tag}_ncu_summary.md (dram__bytes_read.sum + dram__bytes_write.sum per launch)"
    json.dump(old, open(tp, "w"), indent=1)
    print("\n".join(out_md[-12:]))
    print(traffic)


if __name__ == "__main__":
    main()
