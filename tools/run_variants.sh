#!/bin/bash
# usage: tools/run_variants.sh [names...]   (default: every tools/var_*.so)
names="$@"; [ -z "$names" ] && names=$(ls tools/var_*.so | sed 's/.*var_\(.*\)\.so/\1/')
for n in $names; do
  echo "== $n: $(SIMUSCOP_CUDA_LIB=$PWD/tools/var_$n.so python tools/quick_bench.py 256 3 2>&1 | tail -1)"
done
