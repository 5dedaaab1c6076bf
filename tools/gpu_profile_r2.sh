#!/bin/bash
# Round-2 ncu evidence (1 GPU).  Outputs -> gpurun_out/ ; summarised by tools/summarize_ncu.py into profiles/.
mkdir -p gpurun_out
SMALL="python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline --lanes 1"
$SMALL > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches.csv $SMALL > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
$SMALL > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"tc_gemm|conv_halo|lstm_cluster|attention_tc|logmel|conv1_mma|conv1_kernel|add_layernorm|pack_roll|notes_scan" -c 26 -f -o gpurun_out/prof_r2_step60 $SMALL > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out | tail -n 6
