#!/usr/bin/env python
"""One launch of every HBM-bound kernel of bench.py's `roofline_hbm` list, task-batched over 24 slots (argv[1]) - the target of the
multi-slot ncu capture:
    ncu --set full --clock-control none --import-source on -k regex:'dw_|bn_|img_reduce|se_fc|loss_|adam_' \
        -o gpurun_out/prof_hbm python tools/prof_hbm.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

bench._time_launch.__defaults__ = (1, 0)          # reps = 1, warm = 0: each kernel exactly once
flush = bench._Flusher()
for o in bench.hbm_rooflines(int(sys.argv[1]) if len(sys.argv) > 1 else 24, 6550.1, flush):
    print("%-60s %8.1f GB/s" % (o["kernel"], o["GBps"]))
torch.cuda.synchronize()
