"""Per-kernel SASS mnemonic counts of the shipped library, so the TMA / FP64 claims can be checked
without the binary:   python tools/sass_digest.py [lib.so] > profiles/r02_sass_digest.txt
UTMALDG / UTMASTG = tensor-map TMA (cp.async.bulk.tensor), UBLKCP = 1-D bulk TMA, DFMA / DMMA = FP64
vector / tensor pipe, SYNCS = mbarrier operations; UTC*MMA / LDTM (tcgen05) are expected to be absent:
the path is complex FP64 and tcgen05 has no FP64 kind."""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "blockcg_b200", "libblockcg_b200.so")
WANT = ["UTMALDG", "UTMASTG", "UBLKCP", "DFMA", "DMMA", "DADD", "DMUL", "LDS", "STS", "SYNCS", "UTCHMMA", "UTCQMMA", "LDTM", "HMMA"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = {}
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for w in WANT:
            if op == w or op.startswith(w + "."):
                counts[cur][w] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
print("# %s\n# %-88s %7s %s" % (os.path.basename(lib), "kernel", "instr", " ".join("%7s" % w for w in WANT)))
for (k, c), nm in zip(counts.items(), names):
    nm = re.sub(r"\(.*", "", nm).replace("bcg::", "")
    for w in WANT:
        tot[w] += c[w]
    if c["_total"] < 64 and not any(c[w] for w in ("UTMALDG", "UTMASTG", "UBLKCP", "DMMA")):
        continue
    print("%-90s %7d %s" % (nm[:90], c["_total"], " ".join("%7d" % c[w] for w in WANT)))
print("%-90s %7s %s" % ("TOTAL (all kernels)", "", " ".join("%7d" % tot[w] for w in WANT)))
