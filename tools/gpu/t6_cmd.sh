mkdir -p gpurun_out
BCG_DIRAC4_WARPS=26 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "4d_tiled" > gpurun_out/t6_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/t6_pytest.log
for w in 0 26; do echo "warps $w" >> gpurun_out/t6_bench4d.jsonl; BCG_DIRAC4_WARPS=$w timeout 200 python tools/bench4d.py 24 24 24 24 12 60 >> gpurun_out/t6_bench4d.jsonl 2>> gpurun_out/t6_bench4d.err; done
echo "warps 0 32^4" >> gpurun_out/t6_bench4d.jsonl; timeout 200 python tools/bench4d.py 32 32 32 32 12 20 >> gpurun_out/t6_bench4d.jsonl 2>> gpurun_out/t6_bench4d.err
tail -3 gpurun_out/t6_bench4d.err
cat gpurun_out/t6_bench4d.jsonl
