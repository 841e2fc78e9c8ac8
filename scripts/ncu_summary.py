"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.
usage: python scripts/ncu_summary.py <report.ncu-rep> <out.txt> [title]"""
import collections, csv, io, re, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg", "smsp__warps_eligible.avg.per_cycle_active",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
with open(out, "w") as f:
    f.write(f"# {title}\n# source: ncu --set full --clock-control none --import-source on (report read with ncu -i ... --page raw)\n")
    for r in rows[2:]:
        f.write(f"\n## launch {r[hdr.index('ID')]}: {r[hdr.index('Kernel Name')]}\n")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"{k:72s} {r[i]:>16s} {units[i]}\n")
        st = {h: float(r[i]) for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued")}
        tot = sum(st.values()) or 1
        f.write("warp stall samples: " + ", ".join(f"{k.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * v / tot:.1f}%" for k, v in sorted(st.items(), key=lambda x: -x[1])[:10]) + "\n")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    if len(srows) > 2 and "Source" in srows[1]:
        h = srows[1]; iS, iE = h.index("Source"), h.index("Instructions Executed")
        op = collections.Counter()
        for d in srows[2:]:
            if len(d) <= iE or not d[iE].isdigit():
                continue
            m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", d[iS]); op[m.group(1) if m else "?"] += int(d[iE])
        tot = sum(op.values()) or 1
        f.write("\nexecuted warp instructions by opcode (first kernel in the report): " + ", ".join(f"{o} {100 * v / tot:.1f}%" for o, v in op.most_common(16)) + f"; total {tot}\n")
print("wrote", out)
