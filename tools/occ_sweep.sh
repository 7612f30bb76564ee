#!/bin/bash
# resident runs per SM sweep (MPCGPU_RUNS_PER_SM pads the shared-memory request)
out=gpurun_out/occ_sweep.log; : > $out
for rps in 0 8 6 5 4 3; do for pop in 4096 16384; do
  echo "== rps $rps pop $pop" >> $out
  MPCGPU_RUNS_PER_SM=$rps python bench.py --steps 6 --warmup 3 --pop $pop --no-other-configs --no-cpu-baseline 2>>$out | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), d['roofline']['kernel_ms'])" >> $out
done; done
cat $out
