timeout 1500 python -m pytest tests/test_gpu_group.py tests/test_gpu_kernels.py tests/test_gpu_host.py -q -x -s 2>&1 | grep -v "^$" | tail -8
ncu --metrics gpu__time_duration.sum,sm__cycles_active.sum,launch__grid_size --clock-control none --csv \
    --log-file gpurun_out/r02zj_smtime_g16.csv python tools/prof_step.py --gemm-mode tf32x3 --group 16 > gpurun_out/r02zj_prof.log 2>&1
python tools/sm_time.py gpurun_out/r02zj_smtime_g16.csv > gpurun_out/r02zj_sm_time_g16.md; head -30 gpurun_out/r02zj_sm_time_g16.md
