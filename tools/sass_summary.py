"""Opcode histogram per kernel of the shipped library (cuobjdump -sass), the evidence that the tensor-core kernels
really are tcgen05 / TMEM / bulk-TMA code:  python tools/sass_summary.py > profiles/sass_summary.txt
UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / tcgen05.st (tensor memory), UTCBAR = tcgen05.commit,
UBLKCP = cp.async.bulk (1-D bulk TMA), SYNCS = mbarrier, UCGABAR = cluster barrier, REDUX = redux.sync."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "speech-imagery-eeg_b200", "lib", "libign_b200.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "UTMALDG", "SYNCS", "UCGABAR", "REDUX", "LDGSTS",
       "FFMA", "FADD", "FSETP", "FMNMX", "MUFU", "DADD", "DFMA", "LDS", "STS", "LDG", "STG", "HMMA", "IMMA")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = collections.Counter()
            kernels[m.group(1)] = cur
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            if m.group(1) in ("LDTM", "STTM", "UBLKCP", "MUFU") and m.group(2):
                cur[m.group(1) + m.group(2)] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# cuobjdump -sass of %s: instruction counts per kernel (selected opcodes; total = all SASS instructions)" % os.path.relpath(LIB, ROOT))
    for (name, cnt), dm in zip(kernels.items(), demangle):
        short = dm.replace("ign::(anonymous namespace)::", "").replace("(anonymous namespace)::", "")
        short = re.sub(r"^void ", "", short)
        short = re.sub(r"\((?:const |ign::|float|int|unsigned|void|double).*$", "", short)       # drop the argument list, keep <template args>
        sel = ["%s=%d" % (k, v) for k, v in sorted(cnt.items()) if any(k == p or k.startswith(p + ".") for p in KEY)]
        print("%-70s total=%-6d %s" % (short[:70], sum(v for k, v in cnt.items() if "." not in k), " ".join(sel)))


if __name__ == "__main__":
    sys.exit(main())
