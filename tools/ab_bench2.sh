#!/bin/bash
# like ab_bench.sh with extra bench.py arguments: tools/ab_bench2.sh <tag> "<extra args>" <lib...>
tag=$1; extra=$2; shift; shift
out=gpurun_out/ab_${tag}.log
: > $out
for lib in "$@"; do
  echo "== $lib $extra" >> $out
  MPCGPU_LIB=$PWD/$lib timeout 300 python bench.py --steps ${AB_STEPS:-6} --warmup 3 $extra --no-other-configs --no-cpu-baseline 2>> $out | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('value','ms_per_step','failed_candidates','counters')}), d['roofline']['kernel_ms'])" >> $out 2>&1
done
cat $out
