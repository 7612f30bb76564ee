#!/bin/bash
# A/B runs of bench.py over several builds of libmpcgpu.so (MPCGPU_LIB override) on one GPU box.
#   tools/ab_bench.sh <tag> <lib1.so> [<lib2.so> ...]   -> gpurun_out/ab_<tag>.log
tag=$1; shift
mkdir -p gpurun_out
out=gpurun_out/ab_${tag}.log
: > $out
for lib in "$@"; do
  for pop in ${AB_POPS:-4096 16384}; do
    echo "== $lib pop $pop" >> $out
    MPCGPU_LIB=$PWD/$lib timeout 300 python bench.py --steps ${AB_STEPS:-10} --warmup 3 --pop $pop --no-other-configs --no-cpu-baseline 2>> $out | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('value','ms_per_step','failed_candidates','counters')}), d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'])" >> $out 2>&1
  done
done
cat $out
