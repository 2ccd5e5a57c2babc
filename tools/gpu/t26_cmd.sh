mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/t26_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/t26_pytest.log
timeout 200 python tools/ab_lib.py - 1000,12,1000000 41472,12,600 > gpurun_out/t26_new.jsonl 2> gpurun_out/t26_new.err; echo new rc=$?
python - <<'PY'
import json
want={1000:"e030e071b0697fef",41472:"f2b45b66747f6688",331776:"722080df4e5eb215"}
for l in open("gpurun_out/t26_new.jsonl"):
    d=json.loads(l); print(d["V"], d["sbcgrq_sha256"], "identical" if want[d["V"]]==d["sbcgrq_sha256"] else "DIFFERENT", "%.4f"%d["ms_per_iteration"])
PY
