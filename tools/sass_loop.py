"""Static opcode count of the innermost hot loop of a kernel in a cubin / .o / .so (no GPU needed):
    python tools/sass_loop.py <file> <function-substring> [marker-opcode, default LDTM] [min markers, default 6]
The loop is the smallest backward-branch range holding at least <min> marker instructions.  Prints opcodes by pipe class
(alu: VIMNMX*, VIADD, IADD3, LOP3, ISETP, SHF, SEL, PLOP3, LEA, PRMT, FMNMX, FSEL, FSETP; fma: FFMA, FMUL, FADD, IMAD*...)."""
import collections, re, subprocess, sys
path, func = sys.argv[1], sys.argv[2]
marker = sys.argv[3] if len(sys.argv) > 3 else "LDTM"
need = int(sys.argv[4]) if len(sys.argv) > 4 else 6
out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
ins, on = [], False
for l in out.splitlines():
    if "Function :" in l:
        on = func in l
        continue
    if not on:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
best = None
for i, (a, s) in enumerate(ins):
    m = re.search(r"\bBRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?`?\(?(0x[0-9a-f]+)", s)
    if not m:
        continue
    t = int(m.group(1), 16)
    if t in addr and addr[t] <= i:
        j = addr[t]
        nm = sum(1 for _, q in ins[j:i + 1] if marker in q)
        if nm >= need and (best is None or i - j < best[1] - best[0]):
            best = (j, i)
if best is None:
    sys.exit("no loop found")
j, i = best
ALU = ("VIMNMX", "VIADD", "IADD3", "LOP3", "ISETP", "SHF", "SEL", "PLOP3", "LEA", "PRMT", "FMNMX", "FSEL", "FSETP", "IABS", "POPC", "FLO", "MOV", "VABSDIFF")
FMA = ("FFMA", "FMUL", "FADD", "IMAD", "HFMA2", "FFMA2", "FADD2", "FMUL2")
cnt, cls = collections.Counter(), collections.Counter()
for _, s in ins[j:i + 1]:
    op = re.sub(r"^@!?U?P\d+\s+", "", s).split()[0]
    base = op.split(".")[0]
    cnt[op] += 1
    cls["alu" if base.startswith(ALU) else "fma" if base.startswith(FMA) else "other"] += 1
print(f"loop {ins[j][0]:#x}..{ins[i][0]:#x}: {i - j + 1} instructions, {dict(cls)}")
for op, c in cnt.most_common():
    print(f"  {op:32s} {c}")
