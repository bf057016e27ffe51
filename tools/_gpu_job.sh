timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py -q -x 2>&1 | grep -v "^$" | tail -4
B="python bench.py --steps 4 --warmup 3 --skip-cpu-baseline --skip-kernels"
run() { tag=$1; shift; "$@" 2>gpurun_out/r02zm_$tag.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag value %.2f e2e %.2f' % (d['value'], d['e2e']['value']), {k:(round(v['meta_steps_per_s'],3), v['task_slots_per_rank']) for k,v in (d['meta_train'] or {}).items()})"; }
run base $B
MLIIS_TRAIN_GROUP=10 run r20g10 $B --meta-slots 20
MLIIS_TRAIN_GROUP=20 run r40g20 $B --meta-slots 40
MLIIS_TRAIN_GROUP=5 run f5g5 $B --meta-slots 5
