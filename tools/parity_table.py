#!/usr/bin/env python
"""gpurun_out/parity_report.jsonl (written by tests/test_gpu_model.py::_report on the GPU box) -> the markdown table of
profiles/r2_parity.md:  python tools/parity_table.py gpurun_out/parity_report.jsonl"""
import json
import sys

rows = {}
for line in open(sys.argv[1]):
    line = line.strip()
    if line:
        d = json.loads(line)
        rows.setdefault(d.pop("test"), {}).update(d)
print("| case | max-abs prob | mean-abs prob | max-abs logit | thresholded cells that differ |\n|---|---|---|---|---|")
f = lambda v, fmt: "" if v is None else format(v, fmt)
for k, d in rows.items():
    fl = d.get("flips")
    print(f"| {k} | {f(d.get('prob_max'), '.2e')} | {f(d.get('prob_mean'), '.2e')} | {f(d.get('logit_max'), '.2e')} | "
          f"{'' if fl is None else format(100 * fl, '.3f') + ' %'} |")
