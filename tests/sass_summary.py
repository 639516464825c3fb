"""Per-kernel SASS evidence of the Blackwell-native path: counts of the tcgen05 / TMA mnemonics in the shipped library
(B200_PROFILING.md: UTCHMMA = tcgen05.mma, .2CTA = cta_group::2, LDTM = tcgen05.ld, UTMALDG = TMA tensor load,
UBLKCP = bulk copy, UTCBAR = tcgen05.commit).  Usage: python tests/sass_summary.py > profiles/rN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "video_gcp_b200", "libgcpb200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
MN = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "FFMA"]
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        for k in MN:
            if op == k or op.startswith(k + "."):
                if k == "UTCHMMA" and op.startswith("UTCHMMA.2CTA"):
                    continue
                counts[cur][k] += 1
                break
print("%s: %d kernels, arch %s" % (os.path.basename(lib), len(counts),
      ",".join(sorted(set(re.findall(r"arch = (sm_\w+)", sass))))))
print("%-72s %s" % ("kernel", " ".join("%12s" % k for k in MN)))
for k, c in counts.items():
    print("%-72s %s" % (k[:72], " ".join("%12d" % c[m] for m in MN)))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("%-72s %s" % ("TOTAL", " ".join("%12d" % tot[m] for m in MN)))
