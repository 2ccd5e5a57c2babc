mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --max-it 40 --no-cpu-baseline --profile-iters 0 --weak-iters 0"
timeout 300 $B > gpurun_out/t10_plain.json 2> gpurun_out/t10_plain.err; echo plain rc=$?; tail -2 gpurun_out/t10_plain.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/t10_launches.csv $B > gpurun_out/t10_ncu1.log 2>&1; echo launches rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"shift_stag_kernel|dirac_chain_kernel|axpy_pipe_kernel" -s 36 -c 24 -o gpurun_out/t10_top3 -f $B > gpurun_out/t10_ncu2.log 2>&1; echo full rc=$?
ls -la gpurun_out/
