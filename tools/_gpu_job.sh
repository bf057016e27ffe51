set -u
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
python bench.py > gpurun_out/r03a_bench.json 2> gpurun_out/r03a_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r03a_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['kernel_ms'], 'launches', d['gpu_launches'], 'steps', d['steps'], 'ms/step', d['ms_per_step'])
print('meta', {k:(round(v['meta_steps_per_s'],3), round(v['tasks_per_s'],1)) for k,v in d['meta_train'].items()})
print('miou', d['miou_vs_oracle']['max_abs_diff'], d['miou_vs_oracle']['mean_abs_diff'], 'cpu', d['cpu_baseline']['value'], d['clocks'])
for r in d['roofline_hbm']:
    if 'dw_' in r['kernel']: print('  %-70s %.3f  %.1f us' % (r['kernel'], r['frac'], r['ms']*1e3))
PY
tail -2 gpurun_out/r03a_bench.err
