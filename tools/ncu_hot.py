"""Top stalled SASS instructions per kernel from an .ncu-rep: python tools/ncu_hot.py report.ncu-rep [topN] [kernel-substr]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
sub = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kern, hdr, rows, seen = None, None, [], set()
def flush():
    if not kern or kern in seen or sub not in kern or not rows:
        return
    seen.add(kern)
    ia, isrc, ist, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    tot = sum(int(r[ist] or 0) for r in rows)
    print("\n== %s\n   total samples %d, instructions %d" % (kern[:110], tot, len(rows)))
    order = sorted(range(len(rows)), key=lambda i: -int(rows[i][ist] or 0))[:top]
    for i in sorted(order):
        r = rows[i]
        print("  %5d %5.1f%%  [%4d] %s" % (int(r[ist] or 0), 100.0 * int(r[ist] or 0) / max(tot, 1), i, r[isrc].strip()[:90]))
for rec in csv.reader(io.StringIO(raw)):
    if rec and rec[0] == "Kernel Name":
        flush()
        kern, hdr, rows = rec[1], None, []
    elif rec and rec[0] == "Address":
        hdr = rec
    elif hdr and rec:
        rows.append(rec)
flush()
