"""List the backward-branch loops of one kernel in an object file with their instruction counts and
opcode histograms: python scripts/sass_loops.py <obj> <mangled-substring>"""
import collections, re, subprocess, sys

obj, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, funcs = None, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2)))
for name, ins in funcs.items():
    if pat not in name:
        continue
    print(name, len(ins), "instructions")
    for addr, text in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?`?\(?0x([0-9a-f]+)", text)
        if m and int(m.group(1), 16) < addr:
            tgt = int(m.group(1), 16)
            body = [t for a, t in ins if tgt <= a <= addr]
            ops = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", t).split()[0].split(".")[0] for t in body)
            print(f"  loop {tgt:#x}..{addr:#x}: {len(body)} instrs", dict(ops.most_common(14)))
