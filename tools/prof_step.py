#!/usr/bin/env python
"""One inner step (B=8, 224x224) + one prediction on slot 0, eagerly - the target of the ncu captures:
    ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -c 30 -o gpurun_out/prof_tc_conv \
        python tools/prof_step.py --gemm-mode tf32x3"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mliis_b200 import native as N
from mliis_b200.engine import Engine
from mliis_b200.init import initial_bn_state, initial_variables
from mliis_b200.synthetic import make_task_arrays, parse_records

ap = argparse.ArgumentParser()
ap.add_argument("--gemm-mode", default="tf32x3")
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--group", type=int, default=1, help="task-batched launches over this many slots (inputs in the arena)")
a = ap.parse_args()
mode = {"fp32": N.GEMM_FP32, "tf32": N.GEMM_TF32, "tf32x3": N.GEMM_TF32X3}[a.gemm_mode]
G = a.group
eng = Engine(image_size=224, max_batch=8, n_slots=G, gemm_mode=mode)
for s_ in range(G):
    eng.init_state(s_, initial_variables(eng.ctx.params, 0), *initial_bn_state(eng.n_bn))
x, y = parse_records(*make_task_arrays(1, 10, 224))
if G == 1:
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    idx = torch.arange(8, dtype=torch.int32).cuda()
    for _ in range(a.steps):
        eng.train_step(0, xd, yd, 1e-3, index=idx)
    eng.predict(0, xd, yd, index=torch.arange(5, 10, dtype=torch.int32).cuda(), want_pred=False)
else:
    xs, ys, idxs = [], [], []
    for s_ in range(G):              # the same carve in every slot's staging region (uniform stride)
        stg = eng.staging(s_)
        nx, ny = x.size * 4, y.size * 4
        xs.append(stg[:nx].view(torch.float32).view(x.shape))
        o = (nx + 255) // 256 * 256
        ys.append(stg[o:o + ny].view(torch.float32).view(y.shape))
        o = (o + ny + 255) // 256 * 256
        idxs.append(stg[o:o + 32].view(torch.int32))
        xs[-1].copy_(torch.from_numpy(x)); ys[-1].copy_(torch.from_numpy(y)); idxs[-1].copy_(torch.arange(8, dtype=torch.int32))
    for _ in range(a.steps):
        N.check(eng.lib.mliis_kernel_group(G, eng.slot_stride))
        try:
            eng.train_step(0, xs[0], ys[0], 1e-3, index=idxs[0])
        finally:
            N.check(eng.lib.mliis_kernel_group(1, 0))
torch.cuda.synchronize()
print("done")
