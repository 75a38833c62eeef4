"""Read an `ncu --set full --import-source on` report here (no GPU needed): pipe utilisation, stall reasons and the SASS opcodes
the warps of the kernel sit on.   python tools/ncu_stalls.py gpurun_out/<name>.ncu-rep [n_top]"""
import csv, io, subprocess, sys
from collections import Counter

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
m = dict(zip(hdr, vals))
print(m.get("Kernel Name", "")[:150])
for k in ["gpu__time_duration.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "sm__cycles_elapsed.avg.per_second",
          "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum"]:
    if k in m:
        print(f"  {k:75s} {m[k]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]]) for r in data)
idle = sum(int(r[ix["# Samples"]]) for r in data if r[ix["Source"]].strip().endswith("EXIT ;") or " EXIT" in r[ix["Source"]])
print(f"samples {tot}, of which parked at EXIT (idle warps) {idle}")
agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
print("stall reasons (all warps):", ", ".join(f"{s[6:]} {100 * v / tot:.1f}%" for s, v in sorted(agg.items(), key=lambda x: -x[1])[:10]))
c, e = Counter(), Counter()
for r in data:
    s_ = r[ix["Source"]].strip()
    op = s_.split()[0] if not s_.startswith("@") else s_.split()[1]
    c[op] += int(r[ix["# Samples"]]); e[op] += int(r[ix["Instructions Executed"]])
for op, v in c.most_common(ntop):
    print(f"  {op:42s} samples {v:7d} ({100 * v / max(tot - idle, 1):5.1f}% of active)  executed {e[op]:>12d}")
