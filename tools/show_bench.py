#!/usr/bin/env python
"""Print the headline fields of bench.py JSON lines: python tools/show_bench.py gpurun_out/a.json [...]"""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(f, "ERR", e)
        continue
    ss = d.get("single_stream", {})
    print(f"{f}: value {d['value']} ({d['ms_per_step']} ms)  e2e {d['e2e']['value']}  single-stream {ss.get('value')}  "
          f"sm_mhz {d['clocks']['sm_mhz']}  fell_back {d.get('fell_back_to_one_lane')}  lanes {d['config'].get('lanes')}")
    print("   " + "  ".join(f"{s['stage']} {s['ms_per_launch']}" for s in d["stages"][:14]))
