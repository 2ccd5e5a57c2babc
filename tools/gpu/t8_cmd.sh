mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "4d" > gpurun_out/t8_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/t8_pytest.log
for d in "16 16 16 16 12 60" "24 24 24 24 12 60" "32 32 32 32 12 20" "24 24 24 24 4 60" "24 24 24 24 8 60" "16 16 16 16 16 40"; do for t in 1 0; do BCG_DIRAC4_TILE=$t timeout 200 python tools/bench4d.py $d 2>> gpurun_out/t8_bench4d.err | sed "s/^{/{\"tile\": $t, /" >> gpurun_out/t8_bench4d.jsonl; done; done
tail -3 gpurun_out/t8_bench4d.err
cat gpurun_out/t8_bench4d.jsonl
timeout 200 python tools/bench4d.py 24 24 24 24 12 4 > gpurun_out/t8_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dirac4_tile -s 6 -c 3 -o gpurun_out/t8_d4tile -f python tools/bench4d.py 24 24 24 24 12 4 > gpurun_out/t8_ncu.log 2>&1
echo ncu rc=$?
