#!/usr/bin/env python
"""Clocks per k-step of the tcgen05 TF32 MMA patterns the convolutions issue (mliis_tc_mma_rate): tells whether a kernel
that is 'MMA-issue-bound' sits at the pipe's rate for ITS instruction shapes or below it."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mliis_b200 import native as N

lib = N.lib()
st = torch.cuda.current_stream().cuda_stream
v = C.c_double()
print("| pattern | N | A shift rows | clk / k-step | MAC/clk/SM |")
print("|---|---:|---:|---:|---:|")
for pattern, name, ns in ((0, "1 x N", (64, 112, 128, 224, 256)), (3, "3 x N (x3)", (64, 112, 128, 224, 256)),
                          (1, "N=2n + N=n (wide x3)", (56, 64, 112, 128)), (2, "wide x3, 2 tiles", (112,))):
    for n in ns:
        for shift in ((0, 1, 3) if n == 112 else (0,)):
            N.check(lib.mliis_tc_mma_rate(4096, n, pattern, shift, C.byref(v), st))
            mults = {0: 1, 3: 3, 1: 3, 2: 3}[pattern]
            print("| %s | %d | %d | %.1f | %.0f |" % (name, n, shift, v.value, 128 * n * 8 * mults / v.value))
