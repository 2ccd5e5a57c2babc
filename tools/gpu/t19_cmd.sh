mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q ) > gpurun_out/t19_pytest_mg.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/t19_pytest_mg.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for f in 1 0; do
BCG_FOLD_A=$f timeout 500 $T --master-port 2963$f bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/t19_2gpu_fold$f.json 2> gpurun_out/t19_2gpu_fold$f.err; echo fold$f rc=$?
done
python - <<'PY'
import json
for f in (1,0):
    try:
        d=json.loads(open("gpurun_out/t19_2gpu_fold%d.json"%f).read().strip().splitlines()[-1])
        print("fold",f, d["value"], d["iterations"], d["parity"]["multi_gpu"]["pass"], d["parity"]["multi_gpu"]["converged_x_rel_vs_single_domain"][:3], d.get("weak_scaling",{}).get("ms_per_iteration"), d["loop"]["in_loop_profile"]["ms"], d["clocks"]["sm_mhz"])
    except Exception as e: print(f, "ERR", e)
PY
