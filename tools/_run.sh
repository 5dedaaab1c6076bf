mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k conv --timeout 60 2>&1 | tail -8
timeout 200 python -m pytest tests/test_gpu_model.py -q -m gpu -k "golden or canonical" --timeout 120 2>&1 | tail -3
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -c 300 gpurun_out/bench_a.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_a.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
print([(s["stage"], s["ms_per_launch"], s.get("tflops")) for s in d["stages"] if s["stage"].startswith(("res","freq","conv"))])
PY
