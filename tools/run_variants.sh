#!/bin/bash
for f in tools/var_*.so; do
  for b in 2097152 524288; do
  echo "== $f batch $b"
  QB_BATCH=$b SIMUSCOP_CUDA_LIB=$PWD/$f python tools/quick_bench.py 256 3 2>&1 | tail -1
  done
done
