timeout 600 python -m pytest tests/test_gpu_multiclass.py -q 2>&1 | tail -4
