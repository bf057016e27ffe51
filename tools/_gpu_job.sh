python -m pytest tests/test_gpu_kernels.py -q > gpurun_out/r02h_kernels.log 2>&1; tail -15 gpurun_out/r02h_kernels.log
python -m pytest tests/test_gpu_host.py tests/test_gpu_group.py "tests/test_gpu_parity2.py::test_meta_step_theta_parity_vs_oracle" -q > gpurun_out/r02h_pytest.log 2>&1; tail -15 gpurun_out/r02h_pytest.log
python bench.py --steps 6 --warmup 3 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; tail -c 6000 gpurun_out/r02h_bench.json; tail -5 gpurun_out/r02h_bench.err
