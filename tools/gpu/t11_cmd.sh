mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
# schedule 3 (deferral depth 4) forced on the slab decomposition: parity.multi_gpu + the weak-scaling probe (which picks it by default)
BCG_PAIR=3 timeout 500 $T bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/t11_2gpu_pair3.json 2> gpurun_out/t11_2gpu_pair3.err; echo pair3 rc=$?
tail -3 gpurun_out/t11_2gpu_pair3.err
BCG_PAIR=2 timeout 500 $T bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/t11_2gpu_pair2.json 2> gpurun_out/t11_2gpu_pair2.err; echo pair2 rc=$?
python - <<'PY'
import json
for f in ("gpurun_out/t11_2gpu_pair3.json","gpurun_out/t11_2gpu_pair2.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["iterations"], json.dumps(d["parity"].get("multi_gpu"))[:600], d.get("weak_scaling",{}).get("ms_per_iteration"), d["loop"]["in_loop_profile"]["ms"])
    except Exception as e: print(f, "ERR", e)
PY
