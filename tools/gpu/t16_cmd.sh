mkdir -p gpurun_out
timeout 120 tools/micro/p2p_latency.bin 20000 > gpurun_out/t16_p2p_latency.jsonl 2> gpurun_out/t16_p2p.err; echo p2p rc=$?; cat gpurun_out/t16_p2p_latency.jsonl
( time timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q ) > gpurun_out/t16_pytest_mg.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/t16_pytest_mg.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621"
timeout 600 $T bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/t16_2gpu.json 2> gpurun_out/t16_2gpu.err; echo bench rc=$?
python - <<'PY'
import json
for f in ("gpurun_out/t16_2gpu.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["iterations"], d["parity"]["multi_gpu"]["pass"], d.get("weak_scaling",{}).get("ms_per_iteration"), d["loop"]["in_loop_profile"]["ms"], d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
