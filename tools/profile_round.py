"""One place for the per-round profiling recipe (B200_PROFILING.md) and its post-processing.

On the GPU box (one call, ONE GPU; numbers printed under ncu are never bench values):
    python tools/profile_round.py capture r02          # writes gpurun_out/r02_launches.csv, gpurun_out/r02_attn.ncu-rep (+ hd64)
Here, after the call merged gpurun_out/ back:
    python tools/profile_round.py summarise r02        # writes profiles/r02_launches_summary.csv, profiles/r02_ncu_full_summary.txt,
                                                       # profiles/r02_attn_traffic.json (dram bytes per launch, read by bench.py)
"""
import csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
           "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
           "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]


def capture(tag):
    os.makedirs(OUT, exist_ok=True)
    py = sys.executable
    cmds = [
        # launch list of one bench invocation: kernel SHARES of the step (cold-cache, serialised)
        ["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "-c", "400", "--csv", "--log-file",
         os.path.join(OUT, f"{tag}_launches.csv"), py, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3", "--no-cpu-baseline", "--no-sweep"],
        # full capture of the dominant kernel at the headline shape (third launch), source pages included (-lineinfo build)
        ["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", "regex:sage_attn", "-s", "2", "-c", "1", "-f",
         "-o", os.path.join(OUT, f"{tag}_attn"), py, os.path.join(ROOT, "tools", "run_attn_once.py"), "4", "32", "8192", "128", "0", "per_thread", "3"],
        ["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", "regex:sage_attn", "-s", "2", "-c", "1", "-f",
         "-o", os.path.join(OUT, f"{tag}_attn_hd64"), py, os.path.join(ROOT, "tools", "run_attn_once.py"), "1", "32", "8192", "64", "0", "per_thread", "3"],
    ]
    for c in cmds:
        print("+", " ".join(c), flush=True)
        r = subprocess.run(c, capture_output=True, text=True, timeout=900)
        print((r.stdout + r.stderr)[-600:], flush=True)


def _rows(path):
    with open(path, newline="") as fh:
        lines = [l for l in fh if not l.startswith("==")]
    return list(csv.DictReader(lines))


def summarise(tag):
    os.makedirs(PROF, exist_ok=True)
    # ---- launch list -> shares
    launches = os.path.join(OUT, f"{tag}_launches.csv")
    if os.path.exists(launches):
        agg = {}
        for r in _rows(launches):
            if r.get("Metric Name") != "gpu__time_duration.sum":
                continue
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "us")
            us = v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(unit, 1.0)
            a = agg.setdefault(r["Kernel Name"], [0, 0.0])
            a[0] += 1; a[1] += us
        total = sum(a[1] for a in agg.values()) or 1.0
        with open(os.path.join(PROF, f"{tag}_launches_summary.csv"), "w") as fh:
            fh.write(f"# {tag} — every launch of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` under `ncu --metrics gpu__time_duration.sum "
                     "--clock-control none`\n# cold-cache, serialised: compare SHARES, not absolutes.\nkernel,launches,total_us,share,avg_us\n")
            for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                fh.write(f"\"{k}\",{n},{us:.1f},{us / total:.3f},{us / n:.1f}\n")
        print("wrote", f"profiles/{tag}_launches_summary.csv")
    # ---- full captures -> metric summary + traffic
    lines = [f"# {tag} — `ncu --set full --clock-control none --import-source on` captures on B200 (tools/profile_round.py capture {tag});\n"
             "# units as ncu prints them; numbers under ncu are not bench values.\n"]
    for rep in (f"{tag}_attn", f"{tag}_attn_hd64"):
        path = os.path.join(OUT, rep + ".ncu-rep")
        if not os.path.exists(path):
            continue
        r = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader([l for l in r.stdout.splitlines() if l and not l.startswith("==")]))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        for row in rows[2:]:
            d = dict(zip(hdr, row))
            lines.append(d.get("Kernel Name", rep) + "\n")
            vals = {}
            for m in METRICS:
                if m in d:
                    vals[m] = d[m]
                    lines.append(f"    {m} = {d[m]} {units[hdr.index(m)]}\n")
            if rep == f"{tag}_attn" and "dram__bytes_read.sum" in vals:
                u = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                rd = float(vals["dram__bytes_read.sum"].replace(",", "")) * u.get(units[hdr.index("dram__bytes_read.sum")], 1.0)
                wr = float(vals["dram__bytes_write.sum"].replace(",", "")) * u.get(units[hdr.index("dram__bytes_write.sum")], 1.0)
                json.dump({"kernel": d.get("Kernel Name", ""), "dram_bytes_per_launch": rd + wr, "read": rd, "write": wr,
                           "source": f"gpurun_out/{rep}.ncu-rep (ncu --set full, one launch, B=4 H=32 S=8192 D=128)"},
                          open(os.path.join(PROF, f"{tag}_attn_traffic.json"), "w"), indent=1)
    with open(os.path.join(PROF, f"{tag}_ncu_full_summary.txt"), "w") as fh:
        fh.writelines(lines)
    print("wrote", f"profiles/{tag}_ncu_full_summary.txt")


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "capture":
        capture(sys.argv[2])
    elif len(sys.argv) == 3 and sys.argv[1] == "summarise":
        summarise(sys.argv[2])
    else:
        print(__doc__)
