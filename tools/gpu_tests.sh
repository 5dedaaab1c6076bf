#!/bin/bash
# Run the GPU parity suite category by category (separate processes: a CUDA fault in one
# category must not poison the others), each under its own timeout.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name, pytest -k expression, file
  timeout 600 python -m pytest "$3" -q -m gpu -k "$2" --timeout 300 --timeout-method=thread -x \
      > "gpurun_out/test_$1.log" 2>&1
  echo "$1 exit=$?" | tee -a gpurun_out/summary.txt
  tail -n 3 "gpurun_out/test_$1.log"
}
: > gpurun_out/summary.txt
run gemm "gemm" tests/test_gpu_kernels.py
run conv "conv" tests/test_gpu_kernels.py
run lstm "lstm" tests/test_gpu_kernels.py
run attn "attention" tests/test_gpu_kernels.py
run ints "notes or f1 or sigmoid" tests/test_gpu_kernels.py
run frontend "frontend or logmel" tests/test_gpu_model.py
run model "model or transcribe" tests/test_gpu_model.py
run audio "resampler or transcribe_audio" tests/test_audio.py
run cached "bucketed" tests/test_cached.py
cat gpurun_out/summary.txt
