mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/t15_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/t15_pytest.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/t15_smoke.log 2>&1; echo smoke rc=$?; tail -1 gpurun_out/t15_smoke.log
timeout 700 python bench.py --record-iterations > gpurun_out/t15_bench_default.json 2> gpurun_out/t15_bench_default.err; echo bench rc=$?
timeout 300 python bench.py --impl reference > gpurun_out/t15_bench_reference.json 2> gpurun_out/t15_bench_reference.err; echo ref rc=$?
python - <<'PY'
import json
for f in ("gpurun_out/t15_bench_default.json","gpurun_out/t15_bench_reference.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d.get("iterations"), d.get("roofline",{}).get("frac"), d.get("roofline",{}).get("traffic"), d.get("loop",{}).get("predicted_over_measured"), d.get("loop",{}).get("in_loop_profile"))
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/t15_bench_default.err
