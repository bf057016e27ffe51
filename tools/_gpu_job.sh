timeout 170 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_group.py -q -x 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-160
