#!/bin/bash
# run each listed pytest node in its own process (a device fault poisons the CUDA context of the process)
for t in "$@"; do
  echo "=== $t"
  timeout 120 python -m pytest "tests/test_kernels_gpu.py" -x -q -m gpu -s -k "$t" 2>&1 | grep -E "passed|failed|AcceleratorError:|BcoskError|AssertionError|^flat_|^E  " | head -12
done
