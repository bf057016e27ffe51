timeout 600 python -m pytest tests/test_gpu_dist.py -q -x 2>&1 | tail -15
