#!/bin/bash
# First-contact run on a B200 box: each stage in its own process (a trapped kernel poisons the CUDA context),
# each under a timeout; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, cmd...
  local name=$1 t=$2; shift 2
  echo "=== $name"; timeout $t "$@" > gpurun_out/$name.log 2>&1; local rc=$?
  echo "rc=$rc"; tail -n ${TAILN:-25} gpurun_out/$name.log
}
run a_tile 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -s -x -k test_a_tile
run elementwise 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -s -x -k test_elementwise
run fwd_first 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -s -x -k "test_igemm_forward and 1x1_c64"
run fwd_all 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -s -k "test_igemm_forward"
run dgrad_all 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -s -k "dgrad"
run resnet 900 python -m pytest tests/test_resnet_gpu.py -m gpu -q -s
run smoke 300 python __graft_entry__.py smoke
run bench 900 python bench.py --steps 5 --warmup 3 --layer-table gpurun_out/layers.json
