"""The device math primitives of the fast path, pinned on the GPU against mpmath/numpy."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(lib, op, a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    out = np.empty_like(a)
    lib.check(lib.lib().fb200_debug_math(0, op, lib.dptr(a), lib.dptr(b), a.size, lib.dptr(out)))
    return out


def test_branch_free_sqrt_is_ieee(gpu_lib):
    rs = np.random.RandomState(1)
    a = rs.uniform(1, 4, 400000) * 2.0 ** rs.randint(-960, 1000, 400000)
    a[:1000] = np.arange(1, 1001, dtype=float) ** 2
    mine = _run(gpu_lib, 3, a, a)
    assert np.array_equal(mine, np.sqrt(a))              # numpy sqrt is correctly rounded
    assert np.array_equal(_run(gpu_lib, 4, a, a), np.sqrt(a))


def test_div_log_ratio_atan_accuracy(gpu_lib):
    import mpmath as mp
    mp.mp.dps = 40
    rs = np.random.RandomState(0)
    n = 20000
    a = rs.uniform(-1, 1, n) * 10.0 ** rs.uniform(-20, 20, n)
    b = rs.uniform(0.5, 1, n) * 10.0 ** rs.uniform(-30, 40, n)
    q = _run(gpu_lib, 2, a, b)
    assert np.max(np.abs(q - a / b) / np.spacing(np.abs(a / b))) <= 1.0
    num = rs.uniform(1, 2, n) * 10.0 ** rs.uniform(-10, 30, n)
    den = np.abs(num * np.concatenate([1 + rs.uniform(-1, 1, n // 2) * 10.0 ** rs.uniform(-12, 0, n // 2),
                                       10.0 ** rs.uniform(-8, 8, n // 2)])) + 1e-300
    lr = _run(gpu_lib, 0, num, den)
    ex = np.array([float(mp.log(mp.mpf(x) / mp.mpf(y))) for x, y in zip(num, den)])
    assert np.max(np.abs(lr - ex) / np.spacing(np.abs(ex))) <= 2.0
    re = np.abs(rs.normal(size=n)) * 10.0 ** rs.uniform(-5, 20, n)
    im = rs.normal(size=n) * 10.0 ** rs.uniform(-5, 20, n)
    re[:100] = 0.0
    at = _run(gpu_lib, 1, im, re)
    ex = np.array([float(mp.atan2(mp.mpf(y), mp.mpf(x))) for x, y in zip(re, im)])
    assert np.max(np.abs(at - ex) / np.spacing(np.abs(ex))) <= 2.0
