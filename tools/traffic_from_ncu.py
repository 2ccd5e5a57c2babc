"""DRAM bytes per launch from ncu reports -> profiles/r02_traffic.json (read by bench.py for roofline.traffic):
    python tools/traffic_from_ncu.py WORKLOAD report1.ncu-rep [report2.ncu-rep ...]
Every kernel instance found is listed; per bench.py kernel name the MEAN over its instances is stored
(shift_pair: mean of odd and even launches = the average launch, as bench.py times it)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = [("shift_stag_kernel", "shift_stag4"), ("shift_pair_kernel", "shift_pair"), ("shift_dmma_kernel", "shift_dmma_pair"), ("shift_pipe_kernel", "shift_update"),
         ("axpy_pipe_kernel", "axpy_gram"), ("dirac_chain_kernel", "dirac_gram")]
wname, reps = sys.argv[1], sys.argv[2:]
inst = {}
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, data = rows[0], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in data:
        k = r[ix["Kernel Name"]]
        for sub, nm in NAMES:
            if sub in k:
                rd = float(r[ix["dram__bytes_read.sum"]].replace(",", ""))
                wr = float(r[ix["dram__bytes_write.sum"]].replace(",", ""))
                un = rows[1][ix["dram__bytes_read.sum"]].lower()
                scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(un, 1)
                dur = r[ix["gpu__time_duration.sum"]]
                inst.setdefault(nm, []).append({"read": rd * scale, "write": wr * scale, "duration": dur,
                                                "duration_unit": rows[1][ix["gpu__time_duration.sum"]], "report": os.path.basename(rep)})
                break
path = os.path.join(ROOT, "profiles", "r02_traffic.json")
try:
    out = json.load(open(path))
except Exception:
    out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none; "
                       "written by tools/traffic_from_ncu.py from the reports named in _instances"}
out.setdefault(wname, {})
out.setdefault("_instances", {}).setdefault(wname, {})
for nm, lst in inst.items():
    out[wname][nm] = sum(i["read"] + i["write"] for i in lst) / len(lst)
    out["_instances"][wname][nm] = lst
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out[wname], indent=1))
