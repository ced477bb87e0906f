#!/bin/bash
# regular GPU check: kernel parity, network parity, bench (+layer table)
mkdir -p gpurun_out
run() { local name=$1 t=$2; shift 2; echo "=== $name"; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
run kernels 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x
TAILN=14 run resnet 900 python -m pytest tests/test_resnet_gpu.py -m gpu -q -s
grep -E "parity mode|REPORT" gpurun_out/resnet.log | cut -c1-330
TAILN=3 run bench 900 python bench.py --steps 10 --warmup 3 --layer-table gpurun_out/layers.json ${BENCH_ARGS}
