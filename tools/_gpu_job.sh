timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "tc_wgrad or loss_and_gradients or tensor_core_modes" 2>&1 | grep -v "^$" | tail -3
ncu --metrics gpu__time_duration.sum,sm__cycles_active.sum,launch__grid_size --clock-control none --csv \
    --log-file gpurun_out/r02zq_smtime_g16.csv python tools/prof_step.py --gemm-mode tf32x3 --group 16 > gpurun_out/r02zq_prof.log 2>&1
python tools/sm_time.py gpurun_out/r02zq_smtime_g16.csv > gpurun_out/r02zq_sm_time_g16.md; head -5 gpurun_out/r02zq_sm_time_g16.md
python bench.py --steps 4 --warmup 3 --skip-cpu-baseline --skip-kernels --skip-meta-train 2>gpurun_out/r02zq.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2f e2e %.2f' % (d['value'], d['e2e']['value']))"
