#!/bin/bash
# bottleneck experiment: tensor-core kernels off (MLIIS_TC_DEBUG=16), then drop one launcher family at a time
for fam in none dw_ bn_stats,bn_finalize bn_bwd se_,img_colsum reduce_partials stem_ bilinear,head_,loss_ dec_bn_apply,block_out,bcast_rows,add3 adam_step gemm_,transpose; do
  v=$(MLIIS_TC_DEBUG=16 MLIIS_SKIP=$fam timeout 200 python bench.py --skip-cpu-baseline --skip-e2e --steps 4 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f tasks/s  %.3f ms/task  launches/task %d' % (d['value'], 1e3/d['value'], d['gpu_launches']/4/16))")
  echo "skip=$fam  $v"
done
