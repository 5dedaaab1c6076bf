#!/bin/bash
# bench + ncu evidence for one round (1 GPU).  Outputs -> gpurun_out/
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "lstm" --timeout 300 > gpurun_out/pytest_lstm.log 2>&1; tail -n 2 gpurun_out/pytest_lstm.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_64.json 2> gpurun_out/bench_64.err; tail -c 600 gpurun_out/bench_64.err
python bench.py --steps 3 --warmup 3 --chunks 128 --no-cpu-baseline > gpurun_out/bench_128.json 2> gpurun_out/bench_128.err
SMALL="python bench.py --chunks 16 --steps 1 --warmup 1 --no-cpu-baseline"
$SMALL > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_launches.log 2>&1
$SMALL > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 6 -c 6 -o gpurun_out/prof_tc_gemm $SMALL > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -n 20
