"""cuobjdump -sass digest of every kernel in librdvio_fe.so's objects: instruction count and the mnemonics that show what
the kernel is built from (TMA / mbarrier / REDUX / IDP / packed ops ...), plus registers and shared memory from the
ptxas logs.   python scripts/sass_summary.py > profiles/r2_sass_summary.md"""
import collections, glob, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "rd_vio_b200", "csrc", "_obj")
KEY = ["UTMALDG", "UBLKCP", "SYNCS", "REDUX", "IDP", "IMAD", "FFMA", "FMUL", "FADD", "DADD", "DMUL", "F2F", "F2I", "I2F", "I2FP",
       "LDS", "STS", "LDG", "STG", "ATOMS", "ATOMG", "RED", "SHFL", "PRMT", "LOP3", "SHF", "VIMNMX", "FMNMX", "BAR", "HMMA", "UTCMMA"]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
    except Exception:
        return name


regs = {}
for log in glob.glob(os.path.join(OBJ, "*.ptxas.log")):
    cur = None
    for line in open(log):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = m.group(1)
        m = re.search(r"Used (\d+) registers.*?(?:(\d+) bytes smem)?$", line.strip())
        if m and cur:
            sm = re.search(r"(\d+) bytes smem", line)
            regs[cur] = (int(m.group(1)), int(sm.group(1)) if sm else 0)
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores", line)
        if m and cur:
            regs[cur + "#spill"] = int(m.group(2))

print("# SASS digest of the sm_100a kernels (round 2)\n")
print("`cuobjdump -sass rd_vio_b200/csrc/_obj/*.o` (nvcc 12.9, `-gencode arch=compute_100a,code=sm_100a`), made by `scripts/sass_summary.py`.")
print("Columns: SASS instructions, registers / static shared memory / spill bytes from ptxas, then the count of selected mnemonics.\n")
print("| object | kernel | instr | regs | smem B | spill B | mnemonics |\n|---|---|---|---|---|---|---|")
for obj in sorted(glob.glob(os.path.join(OBJ, "*.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    fn, counts, n = None, None, 0
    rows = []
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if fn:
                rows.append((fn, n, counts))
            fn, counts, n = m.group(1), collections.Counter(), 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and fn:
            n += 1
            op = m.group(1)
            for k in KEY:
                if op == k or op.startswith(k):
                    counts[k] += 1
                    break
    if fn:
        rows.append((fn, n, counts))
    for fn, n, counts in rows:
        r = regs.get(fn, ("?", "?"))
        sp = regs.get(fn + "#spill", 0)
        mn = ", ".join(f"{k} {v}" for k, v in sorted(counts.items(), key=lambda kv: -kv[1]) if v)
        print(f"| {os.path.basename(obj)} | `{demangle(fn)}` | {n} | {r[0]} | {r[1]} | {sp} | {mn} |")
