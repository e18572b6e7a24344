"""Summarise an .ncu-rep: headline metrics per launch + executed-instruction mix by opcode."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__occupancy_limit_registers",
        "sm__cycles_elapsed.avg.per_second", "l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("== launch")
    for k in keys:
        if k in d:
            print(f"  {k} = {d[k]} {units[hdr.index(k)]}")
    stalls = sorted(((float(v), k) for k, v in d.items() if "issue_stalled" in k and k.endswith("per_warp_active.pct") and v), reverse=True)[:6]
    for v, k in stalls:
        print(f"  stall {k.split('issue_stalled_')[1].split('_per_warp')[0]} = {v:.1f}%")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr, data = None, []
for r in rows:
    if r and r[0] == "Kernel Name":
        if data:
            break
    elif r and r[0] == "Address":
        hdr = r
    elif hdr:
        data.append(r)
iS, iE, iSmp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
ops, smp, tot, totS = collections.Counter(), collections.Counter(), 0, 0
for r in data:
    parts = r[iS].split()
    op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
    e, s = int(r[iE] or 0), int(r[iSmp] or 0)
    ops[op] += e
    smp[op] += s
    tot += e
    totS += s
print(f"== first launch: {tot} warp instructions executed, {len(data)} static SASS instructions")
for op, c in ops.most_common(22):
    print(f"  {op:10s} {c / tot * 100:6.2f}% exec   {smp[op] / max(1, totS) * 100:6.2f}% stall samples")
