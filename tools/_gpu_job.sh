timeout 1800 python -m pytest tests/test_gpu_host.py tests/test_gpu_parity2.py -q -x 2>&1 | tail -4
python bench.py --steps 6 --warmup 3 --skip-cpu-baseline --skip-kernels 2>gpurun_out/r02zd_bench.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2f e2e %.2f' % (d['value'], d['e2e']['value'])); print('meta', {k:(round(v['meta_steps_per_s'],3), round(v['tasks_per_s'],1), v['task_slots_per_rank']) for k,v in d['meta_train'].items()})"
tail -3 gpurun_out/r02zd_bench.err
