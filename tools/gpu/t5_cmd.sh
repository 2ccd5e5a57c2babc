mkdir -p gpurun_out
BCG_DIRAC4_WARPS=6 timeout 200 python tools/bench4d.py 24 24 24 24 12 4 > gpurun_out/t5_plain.log 2>&1 && \
BCG_DIRAC4_WARPS=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dirac4_tile -s 6 -c 3 -o gpurun_out/t5_d4tile -f python tools/bench4d.py 24 24 24 24 12 4 > gpurun_out/t5_ncu.log 2>&1
echo rc=$?; tail -3 gpurun_out/t5_ncu.log
