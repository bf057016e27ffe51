#!/bin/bash
# Profile snapshot run on the GPU box (one B200): launch list of the bench command + ncu --set full of the dominant
# kernels on one eager inner step.  usage: tools/profile_snapshot.sh <tag>   (outputs under gpurun_out/)
set -u
tag=${1:-snap}
out=gpurun_out
mkdir -p $out
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 3 --slots 1 --tasks-per-step 1 --no-graph --skip-cpu-baseline --skip-e2e \
    --gemm-mode tf32x3 > $out/${tag}_launches.log 2>&1
for grp in "tc_conv3_kernel:6" "tc_wgrad_kernel:8" "tc_conv_kernel:12" "dw_:14" "bn_bwd:8"; do
  k=${grp%%:*}; c=${grp##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -c $c -f -o $out/${tag}_$k \
      python tools/prof_step.py --gemm-mode tf32x3 > $out/${tag}_$k.log 2>&1
  ncu -i $out/${tag}_$k.ncu-rep --page raw --csv > $out/${tag}_$k.raw.csv 2>/dev/null
  rm -f $out/${tag}_$k.ncu-rep
done
ls -la $out | tail -20
