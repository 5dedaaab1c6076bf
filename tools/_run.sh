mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:resample -c 2 -f -o gpurun_out/prof_resample python tools/bench_stages.py --iters 4 > gpurun_out/ncu_rs.log 2>&1; tail -2 gpurun_out/ncu_rs.log
