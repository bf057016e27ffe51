#!/usr/bin/env python
"""The dominant decoder convolution as the engine runs it (pool_taps + tc_conv3_kernel over the 224 real channels), a
few launches - the target of `ncu --set full -k regex:'tc_conv3|pool_taps'`.  usage: prof_dominant.py [slots per launch = 24]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from mliis_b200 import native as N

bench._time_launch.__defaults__ = (2, 1)
flush = bench._Flusher()
G = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ms, fl = bench.time_dominant_kernel(N.GEMM_TF32X3, flush, G)
print("dominant conv x%d slots: %.1f us  %.1f TFLOP/s" % (G, ms * 1e3, fl / ms / 1e9))
torch.cuda.synchronize()
