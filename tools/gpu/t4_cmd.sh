mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "4d" > gpurun_out/t4_pytest.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/t4_pytest.log
for w in 8 9 12 6; do echo "warps $w" >> gpurun_out/t4_bench4d.jsonl; BCG_DIRAC4_WARPS=$w timeout 200 python tools/bench4d.py 24 24 24 24 12 60 >> gpurun_out/t4_bench4d.jsonl 2>> gpurun_out/t4_bench4d.err; done
for w in 8 12; do echo "warps $w 32^4" >> gpurun_out/t4_bench4d.jsonl; BCG_DIRAC4_WARPS=$w timeout 200 python tools/bench4d.py 32 32 32 32 12 20 >> gpurun_out/t4_bench4d.jsonl 2>> gpurun_out/t4_bench4d.err; done
tail -3 gpurun_out/t4_bench4d.err
cat gpurun_out/t4_bench4d.jsonl
