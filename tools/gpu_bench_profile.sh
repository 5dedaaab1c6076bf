#!/bin/bash
# bench + ncu evidence for one round (1 GPU).  Outputs -> gpurun_out/
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_64.json 2> gpurun_out/bench_64.err; tail -c 400 gpurun_out/bench_64.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
SMALL="python bench.py --chunks 64 --steps 1 --warmup 1 --no-cpu-baseline"
$SMALL > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_launches.log 2>&1
$SMALL > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"tc_gemm|conv_halo|lstm_cluster|attention_tc|logmel|conv1_mma|conv1_kernel|add_layernorm" -c 24 -f -o gpurun_out/prof_step64 $SMALL > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out | tail -n 8
