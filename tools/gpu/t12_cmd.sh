mkdir -p gpurun_out
timeout 200 python tools/solve_few.py 41472 12 40 > gpurun_out/t12_plain.log 2>&1; echo plain rc=$?; tail -1 gpurun_out/t12_plain.log
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"rq_step|dirac_chain_kernel|axpy_pipe_kernel|shift_dmma_kernel" -s 100 -c 10 -o gpurun_out/t12_small -f python tools/solve_few.py 41472 12 40 > gpurun_out/t12_ncu.log 2>&1; echo ncu rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 260 --csv --log-file gpurun_out/t12_launches.csv python tools/solve_few.py 41472 12 40 > gpurun_out/t12_ncu1.log 2>&1; echo launches rc=$?
for v in 82944 165888; do timeout 300 python tools/ab_shift.py $v 12 600 "BCG_PAIR=2" "BCG_PAIR=3,BCG_DEPTH=4" "BCG_PAIR=3,BCG_DEPTH=3" >> gpurun_out/t12_ab_depth.jsonl 2>> gpurun_out/t12_ab.err; done
python - <<'PY'
import json
for l in open("gpurun_out/t12_ab_depth.jsonl"):
    try:
        d=json.loads(l); print(d["variant"], d["V"], d["ms_per_iteration"], d.get("x_rel_vs_first"))
    except Exception as e: print("ERR", e, l[:100])
PY
