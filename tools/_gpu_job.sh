timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02v_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['kernel_ms'], 'launches', d['gpu_launches'], 'steps', d['steps'])
print('meta', {k:(v['meta_steps_per_s'], v['tasks_per_s']) for k,v in d['meta_train'].items()})
print('miou', d['miou_vs_oracle']['max_abs_diff'], d['miou_vs_oracle']['mean_abs_diff'], 'cpu', d['cpu_baseline']['value'])
for o in d['roofline_hbm']: print('%-66s %8.1f GB/s  frac %.3f  %.1f us' % (o['kernel'], o['GBps'], o['frac'], o['ms']*1e3))
PY
tail -2 gpurun_out/r02v_bench.err
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02v_launches_bench.csv python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-meta-train --skip-kernels --slots 2 --tasks-per-step 2 > gpurun_out/r02v_ncu_bench.log 2>&1; tail -2 gpurun_out/r02v_ncu_bench.log | cut -c1-300
python tools/summarize_launches.py gpurun_out/r02v_launches_bench.csv > gpurun_out/r02v_launches_bench.md; head -30 gpurun_out/r02v_launches_bench.md
