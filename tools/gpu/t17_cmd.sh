mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/t17_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/t17_pytest.log
timeout 400 python tools/ab_lib.py tools/gpu/old/libblockcg_b200_old.so > gpurun_out/t17_old.jsonl 2> gpurun_out/t17_old.err; echo old rc=$?
timeout 400 python tools/ab_lib.py - > gpurun_out/t17_new.jsonl 2> gpurun_out/t17_new.err; echo new rc=$?
BCG_FOLD_A=0 timeout 400 python tools/ab_lib.py - 41472,12,600 331776,12,600 > gpurun_out/t17_new_nofold.jsonl 2> gpurun_out/t17_new_nofold.err; echo nofold rc=$?
python - <<'PY'
import json
o=[json.loads(l) for l in open("gpurun_out/t17_old.jsonl")]
n=[json.loads(l) for l in open("gpurun_out/t17_new.jsonl")]
for a,b in zip(o,n):
    same = a["sbcgrq_sha256"]==b["sbcgrq_sha256"] and a.get("bcg_sha256")==b.get("bcg_sha256") and a["iterations"]==b["iterations"]
    print(a["V"],a["N"],"identical" if same else "DIFFERENT", a["iterations"], b["iterations"], "ms/it old %.4f new %.4f"%(a["ms_per_iteration"],b["ms_per_iteration"]))
    if "profile_ms" in a: print("   old",a["profile_ms"]); print("   new",b["profile_ms"])
for l in open("gpurun_out/t17_new_nofold.jsonl"):
    d=json.loads(l); print("nofold", d["V"], d["sbcgrq_sha256"], "%.4f"%d["ms_per_iteration"], d.get("profile_ms"))
PY
tail -3 gpurun_out/t17_new.err
