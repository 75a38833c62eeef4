"""Evidence that the kernels are tcgen05 / TMEM / TMA code: SASS opcode counts per kernel of the in-tree library (no GPU needed).
    python tools/sass_opcodes.py [round-tag]        -> profiles/<tag>_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = os.path.join(ROOT, "sageattention_b200", "lib", "libsageattn_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
OPS = ["UTCIMMA", "UTCQMMA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "MUFU.EX2", "VIMNMX3", "FFMA2", "FADD2", "F2FP", "I2FP", "SYNCS", "USETMAXREG", "HMMA", "IMMA"]
cur, cnt = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); cnt[cur] = collections.Counter(); continue
    if cur:
        for o in OPS:
            if re.search(r"\b" + re.escape(o), line):
                cnt[cur][o] += 1
tot = collections.Counter()
lines = [f"# {tag} — SASS opcode counts per kernel of sageattention_b200/lib/libsageattn_b200.so (cuobjdump -sass, sm_100a), tools/sass_opcodes.py\n",
         "# UTC*MMA = tcgen05.mma (I: kind::i8, Q: kind::f8f6f4, H: kind::f16); LDTM / STTM = tcgen05.ld / st; UTMALDG = TMA tensor load; HMMA / IMMA (legacy mma.sync) = 0\n"]
for f, c in cnt.items():
    if not c:
        continue
    tot.update(c)
    d = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip()
    d = re.sub(r"\(CUtensorMap_st.*", "", d)
    lines.append(f"{d[:150]:150s} " + " ".join(f"{o}={c[o]}" for o in OPS if c[o]) + "\n")
lines.append("TOTAL " + " ".join(f"{o}={tot[o]}" for o in OPS) + "\n")
open(os.path.join(ROOT, "profiles", f"{tag}_sass_opcodes.txt"), "w").writelines(lines)
print(lines[-1])
