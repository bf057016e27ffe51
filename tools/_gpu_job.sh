timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x 2>&1 | tail -3
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_gpu_group.py tests/test_gpu_host.py -q 2>&1 | tail -6
grep "canonical 224" gpurun_out/parity2.log
python bench.py --steps 6 --warmup 3 > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02o_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['kernel_ms'], 'launches', d['gpu_launches'])
print('meta', {k:(v['meta_steps_per_s'], v['tasks_per_s']) for k,v in d['meta_train'].items()})
print('miou', d['miou_vs_oracle']['max_abs_diff'], d['miou_vs_oracle']['mean_abs_diff'])
for o in d['roofline_hbm']: print('%-66s %8.1f GB/s  frac %.3f  %.1f us' % (o['kernel'], o['GBps'], o['frac'], o['ms']*1e3))
PY
tail -3 gpurun_out/r02o_bench.err
for t in 148 222 296; do MLIIS_WG_TARGET=$t python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('WG_TARGET=$t value', d['value'])"; done
