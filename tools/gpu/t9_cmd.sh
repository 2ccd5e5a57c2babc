mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/t9_pytest.log 2>&1; echo pytest rc=$?; tail -6 gpurun_out/t9_pytest.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/t9_smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/t9_smoke.log
timeout 700 python bench.py --record-iterations > gpurun_out/t9_bench_default.json 2> gpurun_out/t9_bench_default.err; echo bench rc=$?
python - <<'PY'
import json
for f in ("gpurun_out/t9_bench_default.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["iterations"], d["roofline"]["frac"], d["roofline"]["kernel"][:40], d["loop"].get("predicted_over_measured"), d["parity"].get("true_residual"), d["parity"]["lockstep"]["pass"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/t9_bench_default.err
