mkdir -p gpurun_out
for f in 1 0; do BCG_FOLD_A=$f timeout 400 python tools/ab_lib.py - 1000,12,1000000 41472,12,600 331776,12,600 > gpurun_out/t18_fold$f.jsonl 2> gpurun_out/t18_fold$f.err; echo fold$f rc=$?; done
python - <<'PY'
import json
for f in (1,0):
  for l in open("gpurun_out/t18_fold%d.jsonl"%f):
    d=json.loads(l); print("fold",f, d["V"], d["sbcgrq_sha256"], "%.4f"%d["ms_per_iteration"], d.get("profile_ms"))
PY
