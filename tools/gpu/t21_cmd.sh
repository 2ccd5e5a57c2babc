mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bit_identical" > gpurun_out/t21_pytest.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/t21_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29641"
timeout 400 $T bench.py --gpus 4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/t21_4gpu.json 2> gpurun_out/t21_4gpu.err; echo bench rc=$?
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/t21_4gpu.json").read().strip().splitlines()[-1])
    print(d["value"], d["iterations"], d["parity"]["multi_gpu"]["pass"], d["parity"]["multi_gpu"]["converged_x_rel_vs_single_domain"][:3], d.get("weak_scaling",{}).get("ms_per_iteration"), d["loop"]["in_loop_profile"]["ms"], d["clocks"]["sm_mhz"])
except Exception as e: print("ERR", e)
PY
tail -3 gpurun_out/t21_4gpu.err
