#!/bin/bash
# A/B bench over library switches: usage  bench_ab.sh "name:persistent:light:light_k_iters[:autotune[:flat_stem[:cluster[:parity_dgrad[:late_in_stages[:pdl]]]]]]" ...
mkdir -p gpurun_out
for cfg in "$@"; do
IFS=: read NAME P L LK AT FS CL PD LI PDL WK <<< "$cfg"
python - <<PY 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$NAME', 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'hbm_frac', round(d['roofline']['frac'],3), 'tc', round(d['roofline_tensor']['achieved']))"
import sys, runpy
sys.path.insert(0, ".")
from bcos_b200 import _lib
from bcos_b200.engine.base import PlanBase
_lib.load().bcosk_set_persistent($P)
_lib.load().bcosk_set_light($L)
_lib.load().bcosk_set_cluster(${CL:-1})
_lib.load().bcosk_set_late_input(${LI:-4})
_lib.load().bcosk_set_pdl(${PDL:-0})
PlanBase.light_k_iters = $LK
PlanBase.wide_k_iters = ${WK:-0}
PlanBase.autotune_default = bool(${AT:-1})
PlanBase.flat_stem = bool(${FS:-1})
PlanBase.flat_3x3 = bool(${FS:-1})
PlanBase.parity_dgrad = bool(${PD:-1})
sys.argv = ["bench.py", "--steps", "10", "--warmup", "3", "--no-cpu-baseline", "--layer-table", "gpurun_out/layers_$NAME.json"]
runpy.run_path("bench.py", run_name="__main__")
PY
done
