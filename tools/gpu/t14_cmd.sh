mkdir -p gpurun_out
for l in libblockcg_b200_old lib_new lib_newDBCG_EXP_CHOL_BRANCHY lib_newDBCG_EXP_MM_LOOP; do
  timeout 300 python tools/ab_lib.py tools/gpu/old/$l.so 1000,12,1000000 41472,12,600 >> gpurun_out/t14_ab.jsonl 2>> gpurun_out/t14.err; echo $l rc=$?
done
python - <<'PY'
import json
for l in open("gpurun_out/t14_ab.jsonl"):
    d=json.loads(l); print(d["lib"], d["V"], d["iterations"], d["sbcgrq_sha256"], "%.4f"%d["ms_per_iteration"], d.get("profile_ms"))
PY
