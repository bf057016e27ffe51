nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_dist.py -q -x 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r02q_bench_2gpu.json 2> gpurun_out/r02q_bench_2gpu.err; tail -5 gpurun_out/r02q_bench_2gpu.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02q_bench_2gpu.json').read().strip().splitlines()[-1])
    print('n_gpus', d['n_gpus'], 'value', d['value'], 'e2e', d['e2e']['value'])
    print('meta', json.dumps(d['meta_train']))
except Exception as e:
    print('no json', e)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -2 | cut -c1-300
