nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r02zs_bench_2gpu.json 2> gpurun_out/r02zs_bench_2gpu.err; tail -3 gpurun_out/r02zs_bench_2gpu.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02zs_bench_2gpu.json').read().strip().splitlines()[-1])
    print('n_gpus', d['n_gpus'], 'value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'])
    print('meta', json.dumps({k:(v['meta_steps_per_s'], v['tasks_per_s'], v['allreduce_ms'], v['ranks_without_tasks']) for k,v in d['meta_train'].items()}))
except Exception as e:
    print('no json', e)
PY
timeout 600 python -m pytest tests -q -m gpu -k "nccl or two_gpu or 2gpu or dist" 2>&1 | tail -3
