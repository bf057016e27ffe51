for c in 112 224 360; do timeout 120 python tools/time_conv.py $c 2>&1 | tail -2; done
echo "--- phase timestamps (SM clocks, CTA 0)"
for c in 112 224; do MLIIS_TC_DEBUG=32 timeout 120 python tools/time_conv.py $c 2>&1 | tail -8; done
echo "--- dgrad shape (Cin=112 -> N=224)"
timeout 120 python tools/time_conv.py 112 224 2>&1 | tail -2
python - <<'PY'
import sys; sys.path.insert(0,'.')
import bench, json, torch
flush = bench._Flusher()
out = bench.hbm_rooflines(6, 6550.1, flush)
for o in out: print('%-60s %8.1f GB/s  frac %.3f  %.1f us' % (o['kernel'], o['GBps'], o['frac'], o['ms']*1e3))
PY
