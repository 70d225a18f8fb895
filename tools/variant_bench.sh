#!/bin/bash
# Bench several builds of libpm (make variant ...) back to back on one GPU:  tools/variant_bench.sh tag name1 name2 ...
# ("main" = lib/libpm.so).  Output: gpurun_out/<tag>_<name>.json
tag=$1; shift
for v in "$@"; do
  lib=computational-fluid-dynamics_b200/lib/libpm_$v.so
  [ "$v" = main ] && lib=computational-fluid-dynamics_b200/lib/libpm.so
  PM_LIB=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-secondary --no-parity $BENCH_ARGS > gpurun_out/${tag}_$v.json 2> gpurun_out/${tag}_$v.err
  echo "$v rc=$? $(python -c "import json;d=json.load(open('gpurun_out/${tag}_$v.json'));print(d['ms_per_step'], d['value'], d['roofline']['ms_per_launch'])" 2>&1 | tail -1)"
done
