"""Summarise an .ncu-rep: key metrics + stall breakdown of the hot loop.  usage: ncu_summary.py rep [min_exec_frac]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__cycles_active.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__cycles_elapsed.avg.per_second",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subunit_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
for h, u, v in zip(hdr, units, vals):
    if h in keys:
        print(f"{h:75s} {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
from collections import Counter
cnt = Counter(int(r[ix['Instructions Executed']]) for r in data if int(r[ix['Instructions Executed']]) > 1000)
mode = max(cnt.items(), key=lambda kv: kv[1] * kv[0])[0]          # the step loop: many instrs x many executions
loop = [r for r in data if 0.9 * mode <= int(r[ix['Instructions Executed']]) <= 1.1 * mode]
tot = sum(int(r[ix['# Samples']]) for r in data)
ls = sum(int(r[ix['# Samples']]) for r in loop)
print(f"samples total {tot}, in hot loop ({len(loop)} instrs) {ls} = {100*ls/tot:.1f}%")
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(int(r[ix[s]]) for r in loop) for s in stalls}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]:
    print(f"   {k:28s} {v:7d} {100*v/ls:5.1f}%")
if len(sys.argv) > 2:
    for r in loop:
        st = {s: int(r[ix[s]]) for s in stalls}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print(r[ix['# Samples']].rjust(6), " ".join(f"{k[6:]}={v}" for k, v in top if v).ljust(34), r[ix['Source']][:90])
