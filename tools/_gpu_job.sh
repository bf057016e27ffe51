timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python bench.py > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02z_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['kernel_ms'], 'launches', d['gpu_launches'], 'steps', d['steps'], 'ms/step', d['ms_per_step'])
print('engine', d['engine'])
print('meta', {k:(v['meta_steps_per_s'], v['tasks_per_s']) for k,v in d['meta_train'].items()})
print('miou', d['miou_vs_oracle']['max_abs_diff'], d['miou_vs_oracle']['mean_abs_diff'], 'cpu', d['cpu_baseline']['value'], 'clocks', d['clocks'])
for o in d['roofline_hbm']: print('%-66s %8.1f GB/s  frac %.3f  %.1f us' % (o['kernel'], o['GBps'], o['frac'], o['ms']*1e3))
PY
tail -2 gpurun_out/r02z_bench.err
