timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_group.py tests/test_gpu_parity2.py -q -x 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_host.py -q -x -k "slot" 2>&1 | tail -2
B="python bench.py --steps 5 --warmup 3 --skip-cpu-baseline --skip-kernels --skip-e2e"
run() { tag=$1; shift; "$@" 2>gpurun_out/r03b_$tag.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$tag value %.2f launches %d' % (d['value'], d['gpu_launches']), {k:round(v['meta_steps_per_s'],3) for k,v in (d['meta_train'] or {}).items()})"; }
run ticket $B
MLIIS_BN_TICKET=0 run noticket $B
