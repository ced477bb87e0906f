#!/bin/bash
# bench with the persistent and the per-tile schedule (A/B), layer tables in gpurun_out/
mkdir -p gpurun_out
for P in 1 0; do
python - <<PY 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('persistent=$P', 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'hbm_frac', round(d['roofline']['frac'],3), 'tc', round(d['roofline_tensor']['achieved']))"
import sys, runpy
sys.path.insert(0, ".")
from bcos_b200 import _lib
_lib.load().bcosk_set_persistent($P)
sys.argv = ["bench.py", "--steps", "10", "--warmup", "3", "--no-cpu-baseline", "--layer-table", "gpurun_out/layers_persist$P.json"]
runpy.run_path("bench.py", run_name="__main__")
PY
done
python - <<'PY'
import json
a = {r["name"]: r for r in json.load(open("gpurun_out/layers_persist1.json"))["rows"]}
b = {r["name"]: r for r in json.load(open("gpurun_out/layers_persist0.json"))["rows"]}
rows = sorted(a.values(), key=lambda r: -r["ms"])[:26]
for r in rows:
    o = b[r["name"]]
    print(f"{r['name']:34s} persist {r['ms']:.3f} ms  per-tile {o['ms']:.3f} ms   {r.get('gbs', 0) or 0:6.0f} / {o.get('gbs', 0) or 0:6.0f} GB/s  {r.get('tflops', 0) or 0:6.1f} / {o.get('tflops', 0) or 0:6.1f} TF/s")
PY
