bash tools/gpu_tests.sh > gpurun_out/tests_o.log 2>&1; cat gpurun_out/summary.txt
bash tools/gpu_bench_profile.sh
