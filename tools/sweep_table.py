"""Format tools/microbench.py --sweep output (jsonl) as the table kept under profiles/:
    python tools/sweep_table.py sweep.jsonl > profiles/r02_microbench_sweep.txt"""
import json
import sys

PEAK = 6535.7
try:
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass
cols = ["dirac", "dirac_gram", "gram", "axpy_gram", "shift_update_S1"]
print("# Block Dirac apply / Gram / fused-update micro-benchmark sweep (BASELINE configs[4]): tools/microbench.py --sweep")
print("# B200, CUDA events, reps after 3 warm-ups; GB/s = algorithmic bytes (SURVEY 8d) / time; %% = of measured HBM copy peak %.1f GB/s." % PEAK)
print("# Fields below ~40 MB sit in the 126 MB L2.  N = 4, 8, 12, 16: pipeline kernels (parity-chain stencil, pipelined Q update,")
print("# tensor-instruction update and Gram); N = 1, 2, 3, 6, 32: first-generation kernels.")
print("%10s %3s | " % ("V", "N") + " | ".join("%-22s" % c for c in cols))
for line in open(sys.argv[1]):
    d = json.loads(line)
    if "error" in d:
        print("%10d %3d | ERROR %s" % (d["V"], d["N"], d["error"]))
        continue
    print("%10d %3d | " % (d["V"], d["N"]) + " | ".join("%8.1f us %5d (%3.0f%%)" % (d[c]["us"], d[c]["GBps"], 100.0 * d[c]["GBps"] / PEAK) for c in cols))
