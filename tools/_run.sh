mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_model.py -q -m gpu -k "streaming" --timeout 120 2>&1 | tail -5
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -c 600 gpurun_out/bench_a.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_a.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"], d["clocks"], d["notes_last_step"])
PY
