#!/bin/bash
# One kernel under ncu --set full (1 GPU): bash tools/gpu_ncu_kernel.sh <kernel regex> <count> <out name>
mkdir -p gpurun_out
SMALL="python bench.py --steps 1 --warmup 1 --no-configs --no-cpu-baseline --lanes 1"
$SMALL > gpurun_out/plain_k.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$1" -c ${2:-2} -f -o gpurun_out/${3:-prof_kernel} $SMALL > gpurun_out/ncu_kernel.log 2>&1
tail -2 gpurun_out/ncu_kernel.log
