timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -q -x 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --skip-cpu-baseline --skip-meta-train > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02p_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['kernel_ms'], 'launches', d['gpu_launches'])
for o in d['roofline_hbm']: print('%-66s %8.1f GB/s  frac %.3f  %.1f us' % (o['kernel'], o['GBps'], o['frac'], o['ms']*1e3))
PY
tail -3 gpurun_out/r02p_bench.err
echo "--- ncu: HBM kernels, 6 slots per launch"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'dw_|bn_|img_reduce|se_fc|loss_|adam_kernel|tc_conv_kernel|reduce_partials' -c 80 -f -o gpurun_out/r02p_hbm python tools/prof_hbm.py > gpurun_out/r02p_hbm.log 2>&1; tail -2 gpurun_out/r02p_hbm.log
ncu -i gpurun_out/r02p_hbm.ncu-rep --page raw --csv > gpurun_out/r02p_hbm.raw.csv 2>/dev/null; python tools/ncu_summary.py < gpurun_out/r02p_hbm.raw.csv > gpurun_out/r02p_hbm.md; head -60 gpurun_out/r02p_hbm.md
echo "--- ncu: dominant conv"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_conv3|pool_taps' -c 6 -f -o gpurun_out/r02p_conv3 python tools/prof_dominant.py > gpurun_out/r02p_conv3.log 2>&1; tail -2 gpurun_out/r02p_conv3.log
ncu -i gpurun_out/r02p_conv3.ncu-rep --page raw --csv > gpurun_out/r02p_conv3.raw.csv 2>/dev/null; python tools/ncu_summary.py < gpurun_out/r02p_conv3.raw.csv
echo "--- SM time of one step"
timeout 900 ncu --metrics gpu__time_duration.sum,sm__cycles_active.sum,launch__grid_size --clock-control none --csv --log-file gpurun_out/r02p_smtime.csv python tools/prof_step.py --gemm-mode tf32x3 > gpurun_out/r02p_smtime.log 2>&1; python tools/sm_time.py gpurun_out/r02p_smtime.csv > gpurun_out/r02p_smtime.md; head -30 gpurun_out/r02p_smtime.md
rm -f gpurun_out/r02p_hbm.ncu-rep
