#!/bin/bash
# quick_bench over the kernel variants built into tools/var_*.so
for f in tools/var_*.so; do
  echo "== $f"
  SIMUSCOP_CUDA_LIB=$PWD/$f python tools/quick_bench.py 256 2 2>&1 | tail -1
done
