"""Summarises ncu artefacts into small tracked text files under profiles/.
  python scripts/ncu_summary.py launches gpurun_out/launches.csv profiles/r1_launches.md "<cmd>"
  python scripts/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/r1_numeric_rows.md "<note>"
"""
import collections, csv, re, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__cycles_elapsed.avg", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def launches(src, dst, cmd=""):
    rows = list(csv.reader(open(src)))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    agg = collections.OrderedDict()
    for d in data:
        n = re.sub(r"\(.*", "", re.sub(r"<.*", "", d["Kernel Name"]))[:70]
        agg.setdefault(n, []).append(float(d["Metric Value"].replace(",", "")) / 1e3)
    other = lambda n: ("k_probe" in n) or ("at::" in n)   # roofline probes and torch's L2-flush fill: not part of a step
    tot = sum(sum(v) for n, v in agg.items() if not other(n))
    with open(dst, "w") as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n`{cmd}`\n\n")
        f.write("Share = share of the assembly kernels of one step (roofline probe kernels and torch's L2-flush fill are listed but excluded).\n\n")
        f.write("| kernel | launches | mean us | total us | share of step kernels |\n|---|---|---|---|---|\n")
        for n, v in agg.items():
            sh = "-" if other(n) else f"{100*sum(v)/tot:.1f}%"
            f.write(f"| {n} | {len(v)} | {sum(v)/len(v):.1f} | {sum(v):.1f} | {sh} |\n")
    print(open(dst).read())


def full(src, dst, note=""):
    # src: a .ncu-rep, or the CSV of `ncu -i rep --page raw --csv` exported on the GPU box (reports over 64 MB do not travel)
    out = open(src).read() if src.endswith(".csv") else subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    u = dict(zip(hdr, units))
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of {src.split('/')[-1]}\n\n{note}\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"\n## {d['Kernel Name'][:110]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k]} | {u.get(k,'')} |\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:])
