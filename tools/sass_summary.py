"""Per-kernel SASS evidence for profiles/: python tools/sass_summary.py > profiles/r2/sass_summary.txt

Counts, for every kernel in libm2h_b200.so, the mnemonics that prove the Blackwell-native paths (B200_PROFILING.md):
UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG (TMA), UTCBAR, FADD2/FMUL2/FFMA2 (packed fp32),
the scalar FP32 ops, and local-memory traffic (LDL/STL = spills)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "move2hear_active_av_separation_b200", "libm2h_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "FADD2", "FMUL2", "FFMA2", "FADD", "FMUL", "FFMA", "HMMA", "LDL", "STL", "SHFL", "LDS", "STS", "LDG", "STG", "BAR"]
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["total"] += 1
        if op in KEYS:
            counts[cur][op] += 1
        elif op.startswith("UTC") and op.endswith("MMA"):
            counts[cur]["UTCHMMA"] += 1
print("SASS mnemonic counts per kernel of", os.path.basename(lib), "(static instruction counts; cuobjdump -sass)")
print("%-64s %6s  %s" % ("kernel", "instr", "non-zero counts"))
for k, c in counts.items():
    print("%-64s %6d  %s" % (k[:64], c["total"], "  ".join("%s %d" % (n, c[n]) for n in KEYS if c[n])))
