mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k lstm -x --timeout 120 2>&1 | tail -5
AMT_LSTM_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/trace.json 2> gpurun_out/trace.err; grep "trace" gpurun_out/trace.err | head -3
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -c 300 gpurun_out/bench_a.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_a.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
print([(s["stage"], s["ms_per_launch"]) for s in d["stages"]][:8])
PY
