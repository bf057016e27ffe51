set -u
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
python bench.py > gpurun_out/r02zr_bench.json 2> gpurun_out/r02zr_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02zr_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['kernel_ms'], 'launches', d['gpu_launches'], 'steps', d['steps'], 'ms/step', d['ms_per_step'])
print('meta', {k:(round(v['meta_steps_per_s'],3), round(v['tasks_per_s'],1)) for k,v in d['meta_train'].items()})
print('miou', d['miou_vs_oracle']['max_abs_diff'], d['miou_vs_oracle']['mean_abs_diff'], 'cpu', d['cpu_baseline']['value'], d['clocks'])
for r in d['roofline_hbm']: print('  %-70s %.3f  %.1f us' % (r['kernel'], r['frac'], r['ms']*1e3))
PY
tail -2 gpurun_out/r02zr_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_conv3|pool_taps' -c 4 -f -o gpurun_out/r02zr_conv3 python tools/prof_dominant.py 24 > gpurun_out/r02zr_conv3.log 2>&1
ncu -i gpurun_out/r02zr_conv3.ncu-rep --page raw --csv > gpurun_out/r02zr_prof_conv3_g24.raw.csv 2>/dev/null; rm -f gpurun_out/r02zr_conv3.ncu-rep
python tools/ncu_summary.py < gpurun_out/r02zr_prof_conv3_g24.raw.csv
timeout 900 ncu --set full --clock-control none -k regex:'dw_|bn_|img_reduce|se_fc|loss_|adam_|tc_conv_kernel|tc_pw' -c 60 -f -o gpurun_out/r02zr_hbm python tools/prof_hbm.py 24 > gpurun_out/r02zr_hbm.log 2>&1
ncu -i gpurun_out/r02zr_hbm.ncu-rep --page raw --csv > gpurun_out/r02zr_prof_hbm_g24.raw.csv 2>/dev/null; rm -f gpurun_out/r02zr_hbm.ncu-rep
python tools/ncu_summary.py < gpurun_out/r02zr_prof_hbm_g24.raw.csv > gpurun_out/r02zr_ncu_hbm_g24.md; wc -l gpurun_out/r02zr_ncu_hbm_g24.md
