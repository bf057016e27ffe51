# multi-GPU validation job: `gpurun --gpus 2 --timeout 1500 -- bash tools/_gpu_job2.sh`
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -c 1500 gpurun_out/bench_2gpu.json
timeout 600 python -m pytest tests/test_gpu_dist.py -q 2>&1 | tail -3
