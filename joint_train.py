#!/usr/bin/env python
"""Trains an image segmentation model on all classes jointly with SGD, on the B200 engine.

Drop-in for /root/reference/joint_train.py (same flags); the implementation lives in mliis_b200/joint_train.py.
Multi-GPU data parallelism: `torchrun --nproc-per-node N joint_train.py ...` (one NCCL all-reduce of the gradient
per step)."""
from mliis_b200.joint_train import main

if __name__ == "__main__":
    main()
