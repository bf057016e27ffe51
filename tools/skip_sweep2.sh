#!/bin/bash
# Throughput sensitivity under the real multi-slot load: drop one launcher family at a time (results are garbage,
# timing is data-independent) and see how many ms per task disappear.  usage: tools/skip_sweep2.sh <tag> [bench args]
tag=${1:-sweep}; shift
out=gpurun_out/${tag}_skip_sweep.txt
: > $out
for fam in none 'tc_conv$' tc_conv3 tc_wgrad dw_ bn_stats,bn_finalize bn_bwd se_,img_colsum reduce_partials stem_ add3,block_out,dec_bn_apply pool_,region_sums bilinear,head_,loss_,predict adam_step,tc_prep; do
  v=$(MLIIS_SKIP="$fam" timeout 300 python bench.py --skip-cpu-baseline --skip-e2e --steps 4 --warmup 3 "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f tasks/s  %.3f ms/task' % (d['value'], 1e3/d['value']))")
  echo "skip=$fam  $v" | tee -a $out
done
