"""Executed warp instructions per SASS opcode (and their share of the stall samples) of one kernel in an ncu report:
    python tools/ncu_opcodes.py report.ncu-rep kernel_regex scores_per_launch/32
"""
import collections, csv, io, re, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
unit = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
iE, iS = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
data = [r for r in rows[2:] if len(r) > 5 and r[0].startswith("0x")]
a0 = data[0][0]
for i in range(1, len(data)):
    if data[i][0] == a0:
        data = data[:i]
        break
agg, st, tot = collections.Counter(), collections.Counter(), 0
for r in data:
    s = re.sub(r"^@!?U?P\d+\s+", "", r[1].strip())
    op = s.split()[0]
    e = int(r[iE])
    agg[op] += e
    st[op] += int(r[iS])
    tot += e
stot = max(sum(st.values()), 1)
print(f"{len(data)} SASS instructions, {tot} executed warp instructions" + (f", {tot / unit:.3f} per unit" if unit else ""))
for op, e in agg.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 30):
    print(f"{op:34s} {e / tot * 100:5.1f}%" + (f"  {e / unit:6.3f}/unit" if unit else "") + f"  stall {st[op] / stot * 100:5.1f}%")
