mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k conv --timeout 120 2>&1 | tail -15
for v in halo; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; tail -c 300 gpurun_out/bench_$v.err
done
python - <<'PY'
import json
for v in ("halo",):
    d=json.loads(open(f"gpurun_out/bench_{v}.json").read().strip().splitlines()[-1])
    print(v, d["value"], d["ms_per_step"], d["e2e"]["value"])
    print([(s["stage"], s["ms_per_launch"], s.get("tflops")) for s in d["stages"] if s["stage"].startswith(("res","freq","conv"))])
PY
