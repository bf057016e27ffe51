timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -q -x -k "dw or loss_and_gradients or train_steps or full_size" 2>&1 | grep -v "^$" | tail -3
python bench.py --steps 6 --warmup 3 --skip-cpu-baseline --skip-meta-train > gpurun_out/r02zt_bench.json 2>gpurun_out/r02zt.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02zt_bench.json').read().strip().splitlines()[-1])
print('value %.2f e2e %.2f' % (d['value'], d['e2e']['value']))
for r in d['roofline_hbm']:
    if 'dw_' in r['kernel']: print('  %-70s %.3f  %.1f us' % (r['kernel'], r['frac'], r['ms']*1e3))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02zt_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-meta-train --skip-kernels --skip-e2e > gpurun_out/r02zt_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/r02zt_launches_bench.csv > gpurun_out/r02zt_launches_bench.md 2>&1; head -30 gpurun_out/r02zt_launches_bench.md
