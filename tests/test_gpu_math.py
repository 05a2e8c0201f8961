"""The kernels' branch-free FP64 functions (csrc/fastmath.cuh) against numpy float64 on their guaranteed ranges."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(which, x):
    import torch

    from quansino_b200 import _lib

    lib = _lib.load()
    xin = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    out = torch.empty_like(xin)
    _lib.check(lib.qmc_debug_math(which, xin.data_ptr(), out.data_ptr(), xin.numel(), torch.cuda.current_stream().cuda_stream))
    return out.cpu().numpy()


def _ulp_err(got, ref):
    return np.max(np.abs(got - ref) / np.spacing(np.abs(ref)))


def test_exp():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-700, 700, 200000), rng.uniform(-130, 13, 200000), rng.uniform(-1, 1, 100000), [0.0, -700.0, 700.0]])
    assert _ulp_err(_run(0, x), np.exp(x)) <= 2.0


def test_log():
    rng = np.random.default_rng(1)
    x = np.concatenate([np.exp(rng.uniform(-40, 40, 300000)), rng.uniform(0.5, 2.0, 200000), [1.0]])
    got, ref = _run(1, x), np.log(x)
    err = np.abs(got - ref)
    assert np.all(err <= 2.0 * np.spacing(np.abs(ref)) + 1e-17)


def test_sqrt_and_reciprocal():
    rng = np.random.default_rng(2)
    x = np.exp(rng.uniform(-100, 100, 500000))
    assert _ulp_err(_run(2, x), np.sqrt(x)) <= 1.0
    assert _ulp_err(_run(3, x), 1.0 / x) <= 1.0
    assert _ulp_err(_run(3, -x), -1.0 / x) <= 1.0


def test_sincos():
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(0, 2 * np.pi, 500000), [0.0, np.pi / 2, np.pi, 3 * np.pi / 2, 2 * np.pi]])
    for which, ref in ((4, np.sin(x)), (5, np.cos(x))):
        got = _run(which, x)
        assert np.max(np.abs(got - ref)) <= 2.3e-16
