mkdir -p gpurun_out
bash tools/gpu_tests.sh > gpurun_out/tests_o.log 2>&1; cat gpurun_out/summary.txt; grep -h "FAILED\|Error\|passed\|failed" gpurun_out/test_*.log | head -20
