timeout 600 python -m pytest tests/test_gpu_host.py -q -x -k "slot_parallel_meta_step_equals" 2>&1 | tail -12
