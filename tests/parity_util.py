"""Shared helpers for the GPU parity tests: build the same problem for the CPU oracle and the CUDA engine."""
import numpy as np
import torch

from oracle.efficientlab_oracle import Arch, EfficientLabOracle, OptState
from mliis_b200.synthetic import make_task_arrays, parse_records


def make_problem(size=64, n_examples=8, task_id=0, theta_seed=0):
    arch = Arch()
    theta = arch.init_theta(theta_seed, torch.float64).to(torch.float32).to(torch.float64)  # fp32-representable
    bn = arch.init_bn_state(torch.float64)
    iu8, mu8 = make_task_arrays(task_id, n_examples, size)
    images, labels = parse_records(iu8, mu8)
    return arch, theta, bn, images, labels


def split_vars(arch, flat):
    """flat TF-order vector -> list of numpy arrays with TF shapes."""
    flat = flat.detach().cpu().numpy() if isinstance(flat, torch.Tensor) else np.asarray(flat)
    return [flat[p.offset:p.offset + p.size].reshape(p.shape).astype(np.float32) for p in arch.params]


def make_engine(arch, theta, bn, size, max_batch, **kw):
    from mliis_b200.engine import Engine
    eng = Engine(image_size=size, max_batch=max_batch, **kw)
    eng.init_state(0, split_vars(arch, theta), bn[0].numpy(), bn[1].numpy())
    return eng


def rel_err(a, b):
    """max |a-b| / max |b|  (a: engine, b: oracle)."""
    a = a.detach().double().cpu() if isinstance(a, torch.Tensor) else torch.as_tensor(a).double()
    b = b.detach().double().cpu() if isinstance(b, torch.Tensor) else torch.as_tensor(b).double()
    denom = b.abs().max().item()
    return (a - b).abs().max().item() / (denom + 1e-30)


def rel_l2(a, b):
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def per_param_report(arch, g_engine_tf, g_oracle, top=8):
    """list of (rel_l2, name) sorted worst first."""
    rows = []
    for p in arch.params:
        a = g_engine_tf[p.offset:p.offset + p.size]
        b = g_oracle[p.offset:p.offset + p.size]
        rows.append((rel_l2(a, b), float(b.abs().max()), p.name))
    rows.sort(key=lambda r: -r[0])
    return rows[:top]


def oracle_state_from_engine(eng, sgd=False, dtype=torch.float64, slot=0):
    """Every global variable of the engine's slot as an oracle MetaState (theta, BN statistics, Adam slots)."""
    from oracle.meta_oracle import state_from_flat
    torch.cuda.synchronize()
    p = eng.powers(slot).cpu()
    return state_from_flat(eng.tf_order_vector(eng.theta(slot)), eng.bn_state(slot),
                           eng.tf_order_vector(eng.adam_v(slot)), float(p[0]), float(p[1]), sgd=sgd, dtype=dtype)
