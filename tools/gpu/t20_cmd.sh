mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/t20_pytest.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/t20_pytest.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/t20_smoke.log 2>&1; echo smoke rc=$?; tail -1 gpurun_out/t20_smoke.log
timeout 700 python bench.py --record-iterations > gpurun_out/t20_bench_default.json 2> gpurun_out/t20_bench_default.err; echo bench rc=$?
B="python bench.py --steps 1 --warmup 1 --max-it 40 --no-cpu-baseline --profile-iters 0 --weak-iters 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/t20_launches.csv $B > gpurun_out/t20_ncu1.log 2>&1; echo launches rc=$?
python - <<'PY'
import json
for f in ("gpurun_out/t20_bench_default.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d.get("iterations"), d["roofline"]["frac"], d["roofline"]["traffic"], d["loop"]["predicted_over_measured"], d["loop"]["in_loop_profile"]["ms"], d["clocks"], d["parity"]["lockstep"]["pass"], max(d["parity"]["true_residual"]))
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/t20_bench_default.err
