mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "4d or bit_identical or solve_statistics or headline" > gpurun_out/t3_pytest.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/t3_pytest.log
for m in 1 0; do BCG_DIRAC4_TILE=$m timeout 200 python tools/bench4d.py 24 24 24 24 12 60 >> gpurun_out/t3_bench4d.jsonl 2>> gpurun_out/t3_bench4d.err; echo bench4d tile=$m rc=$?; done
BCG_DIRAC4_TILE=1 timeout 200 python tools/bench4d.py 32 32 32 32 12 20 >> gpurun_out/t3_bench4d.jsonl 2>> gpurun_out/t3_bench4d.err
tail -3 gpurun_out/t3_bench4d.err
timeout 600 python bench.py > gpurun_out/t3_bench_default.json 2> gpurun_out/t3_bench_default.err; echo bench rc=$?
BCG_PAIR=2 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/t3_bench_pair2.json 2> gpurun_out/t3_bench_pair2.err; echo bench2 rc=$?
python - <<'PY'
import json
for f in ("gpurun_out/t3_bench_default.json","gpurun_out/t3_bench_pair2.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["iterations"], d["roofline"]["frac"], d["roofline"]["kernel"][:30], d["loop"].get("predicted_over_measured"), d["parity"].get("true_residual"))
    except Exception as e: print(f, "ERR", e)
PY
cat gpurun_out/t3_bench4d.jsonl
