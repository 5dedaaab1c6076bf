#!/bin/bash
# Run the GPU parity suite file by file (separate processes: a CUDA fault in one file must not poison the
# others), each under its own timeout.  Logs -> gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name, file, extra pytest args
  timeout 1500 python -m pytest "$2" -q -m gpu --timeout 600 --timeout-method=thread ${3} \
      > "gpurun_out/test_$1.log" 2>&1
  echo "$1 exit=$?" | tee -a gpurun_out/summary.txt
  tail -n 4 "gpurun_out/test_$1.log"
}
: > gpurun_out/summary.txt
run kernels tests/test_gpu_kernels.py
run model tests/test_gpu_model.py
run audio tests/test_audio.py
run cached tests/test_cached.py
cat gpurun_out/summary.txt
[ -f gpurun_out/parity_report.jsonl ] && cat gpurun_out/parity_report.jsonl
