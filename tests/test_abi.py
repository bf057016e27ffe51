"""The C-ABI library loads and exports every symbol include/mliis_b200.h declares; tables match the oracle; compute
entry points fail loudly without an sm_100 device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from mliis_b200 import native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "mliis_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mliis_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(N.LIB_PATH), "build the library first: python -m mliis_b200.build"
    lib = ctypes.CDLL(N.LIB_PATH)
    names = _declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "header declares %s but the library does not export it" % n


def test_binding_covers_the_header():
    bound = {s[0] for s in N.SYMBOLS}
    assert set(_declared_functions()) == bound


def test_tables_match_oracle():
    from oracle.efficientlab_oracle import Arch
    c = N.Context(N.make_config(), -1)
    a = Arch()
    assert c.n_params == a.n_params == 2071714
    assert [(p.name, p.shape, p.l2) for p in c.params] == [(p.name, p.shape, p.l2) for p in a.params]
    assert [(b.scope, b.channels, b.offset, b.fused) for b in c.bns] == [(b.name, b.channels, b.offset, b.fused) for b in a.bns]
    assert c.n_dc == len(a.dc_blocks) == 6
    # flat layout: 16-byte aligned, non-overlapping, L2 tensors first
    spans = sorted((p.offset, p.offset + p.size, p.l2) for p in c.params)
    assert all(s[0] % 4 == 0 for s in spans)
    assert all(spans[i][1] <= spans[i + 1][0] for i in range(len(spans) - 1))
    first_bn = min(s[0] for s in spans if not s[2])
    assert all(s[1] <= first_bn for s in spans if s[2])
    assert c.state_floats == 2 * c.n_theta + 2 * c.n_bn + 4
    c.close()


@pytest.mark.parametrize("size,rsd", [(224, (2, 4)), (320, (2, 4)), (64, (4,)), (96, (2, 3, 4))])
def test_plans_for_other_configs(size, rsd):
    c = N.Context(N.make_config(image_size=size, rsd=rsd), -1)
    assert c.workspace_bytes > 0 and c.n_theta >= c.n_params
    c.close()


def test_bad_configs_are_rejected():
    for kw in [dict(image_size=100), dict(max_batch=0), dict(n_slots=0), dict(rsd=(7,)), dict(final_dropout_rate=1.5)]:
        with pytest.raises(N.MliisError) as e:
            N.Context(N.make_config(**kw), -1)
        assert e.value.code == N.MLIIS_ERR_ARG


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    c = N.Context(N.make_config(), -1)
    lib = N.lib()
    rc = lib.mliis_forward(c.handle, 0, None, None, 1, 0, None, None, 0, None, None)
    assert rc == N.MLIIS_ERR_DEVICE and b"fallback" in lib.mliis_last_error()
    rc = lib.mliis_gemm_nn(None, None, None, 8, 8, 8, 0, None)
    assert rc == N.MLIIS_ERR_DEVICE
    with pytest.raises(N.MliisError):
        from mliis_b200.engine import Engine
        Engine()
    c.close()
