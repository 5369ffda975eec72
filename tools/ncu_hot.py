"""Top stall sites of one kernel in an ncu report (SASS view, warp-sampling counts), with the opcode's dominant
stall reason.  Usage: python tools/ncu_hot.py report.ncu-rep kernel_regex [top_n]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[1]
ns, si = h.index("# Samples"), h.index("Source")
stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
data, seen = [], set()
for idx, r in enumerate(rows[2:]):
    if len(r) <= ns or not r[0].startswith("0x"):
        continue
    if r[0] in seen:
        break
    seen.add(r[0])
    top = sorted(((int(r[i] or 0), h[i][6:]) for i in stall), reverse=True)[:2]
    data.append((int(r[ns]), idx, r[si].strip()[:64], top))
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for d in sorted(data, reverse=True)[:top_n]:
    print("%6d %5.1f%% #%4d %-64s %s" % (d[0], 100 * d[0] / tot, d[1], d[2], d[3]))

# per-region view: buckets of 60 instructions with the memory / sync opcodes they contain
print()
for b in range(0, len(data), 60):
    chunk = sorted(d for d in data if b <= d[1] < b + 60)
    s_ = sum(c[0] for c in chunk)
    if s_ < tot / 400:
        continue
    ops = set()
    for c in chunk:
        t = c[2].split()
        op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        if op.startswith(("LDTM", "STTM", "UTC", "LDS", "STS", "STG", "ATOMS", "REDUX", "CREDUX", "BAR", "SYNCS", "LDGSTS", "UBLKCP", "LDG", "NANOSLEEP", "VOTE")):
            ops.add(op.split(".")[0])
    print("%5d-%5d %6d samples %5.1f%%  %s" % (b, b + 59, s_, 100 * s_ / tot, ",".join(sorted(ops))))
