mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 400 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["gathered_notes"])
PY
