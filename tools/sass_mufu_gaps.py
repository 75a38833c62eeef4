"""Static check of the exponential loops in the SASS of a kernel: how evenly ptxas spaced the MUFU.EX2 instructions.
Back-to-back MUFUs of one warp issue ~8 cycles apart on B200 (measured: tools/microbench), interleaved ones ~2-4, so long runs of
consecutive MUFUs are what makes a softmax loop slow.   python tools/sass_mufu_gaps.py <lib.so> <kernel-name-substring>"""
import subprocess, sys, re
from collections import Counter
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, funcs = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        funcs[cur].append(line.split("*/", 1)[1].strip())
for name, ins in funcs.items():
    if pat not in name:
        continue
    pos = [i for i, s in enumerate(ins) if "MUFU.EX2" in s]
    if not pos:
        continue
    # split into loops (gap > 60 instructions)
    loops, start = [], 0
    for k in range(1, len(pos) + 1):
        if k == len(pos) or pos[k] - pos[k - 1] > 60:
            loops.append(pos[start:k]); start = k
    print(name[:110])
    for lp in loops:
        gaps = [b - a for a, b in zip(lp, lp[1:])]
        runs, r = [], 1
        for g in gaps:
            if g == 1: r += 1
            else: runs.append(r); r = 1
        runs.append(r)
        c = Counter(gaps)
        print(f"  loop at {lp[0]:5d}: {len(lp):3d} MUFU in {lp[-1] - lp[0] + 1:4d} instrs; back-to-back pairs {c.get(1, 0):3d}; longest run {max(runs):2d}; gap histogram {dict(sorted(c.items()))}")
    break
