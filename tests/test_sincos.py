"""The simulator's table-driven sin/cos (roborugby_b200/csrc/rr_sincos.cuh), host build, against glibc
(what the reference uses through CPython's math module) and against mpmath at 200 bits.

The routine is made of IEEE +,-,*,fma only, so the bits checked here are the bits the GPU produces."""
import math

import numpy as np
import pytest


def _args():
    rng = np.random.default_rng(1)
    rot = (rng.integers(0, 361, 60000) + rng.integers(-600, 600, 60000) * 0.6) % 360.0
    return {
        "uniform": rng.uniform(-1.6, 7.9, 60000),
        "360-rot": np.radians(360 - rot),   # MyUtils.py:284
        "rot": np.radians(rot),             # RR_Robot.py:182
        "rot+90": np.radians(rot + 90),     # RR_Robot.py:167
        "rot-90": np.radians(rot - 90),     # RR_Robot.py:174
        "special": np.radians(np.array([0.0, 45, 90, 135, 180, 225, 270, 315, 360, 0.6, 1.2, 359.4, 89.4, 90.6])),
    }


def test_sincos_vs_glibc_and_mpmath():
    from emul.emul import sincos
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 200
    total = differ = mine_wrong = 0
    for name, x in _args().items():
        s, c = sincos(x)
        gs = np.array([math.sin(v) for v in x]); gc = np.array([math.cos(v) for v in x])
        assert np.max(np.abs(s - gs) / np.spacing(np.maximum(np.abs(gs), 1e-300))) <= 1.0, name
        assert np.max(np.abs(c - gc) / np.spacing(np.maximum(np.abs(gc), 1e-300))) <= 1.0, name
        total += 2 * len(x)
        for got, ref, fn in ((s, gs, mp.sin), (c, gc, mp.cos)):
            bad = np.nonzero(got != ref)[0]
            differ += len(bad)
            for i in bad:  # wherever the two differ, ours must be the correctly rounded one
                if got[i] != float(fn(mp.mpf(float(x[i])))):
                    mine_wrong += 1
    print(f"{total} evaluations: {differ} differ from glibc by 1 ulp ({100.0 * differ / total:.3f} %), "
          f"{mine_wrong} of those are not the correctly rounded value")
    assert mine_wrong == 0
    assert differ < 0.004 * total


def test_sincos_exact_cases_and_fallback():
    from emul.emul import sincos
    s, c = sincos(np.array([0.0, -0.0]))
    assert s[0] == 0.0 and c[0] == 1.0 and c[1] == 1.0
    x = np.array([1e3, -1e5, np.inf, np.nan, 17.0])
    s, c = sincos(x)
    assert np.allclose(s[[0, 1, 4]], np.sin(x[[0, 1, 4]]), rtol=0, atol=1e-15)
    assert np.isnan(s[2]) and np.isnan(s[3])


def test_sincos_grid_path_equals_dd_path():
    """rr_sincos_grid's fast path (multiples of 0.2 degrees plus rounding drift) against rr_sincos_dd, which the test
    above pins against mpmath: the two agree except in the few cases per million where the exact value lies
    within rr_sincos_dd's 2^-13 ulp of a rounding boundary, and there the grid path (2^-100 before its final
    rounding) must be the correctly rounded one.  Off-grid arguments take the dd path: identical bits."""
    from emul.emul import sincos
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 300
    rng = np.random.default_rng(7)
    n = 400000
    deg = rng.integers(-450, 2251, n) * 0.2
    # accumulated drift of thousands of float additions, up to the fast path's limit and beyond it
    drift = rng.standard_normal(n) * 10.0 ** rng.uniform(-17, -10.5, n)
    fams = {
        "exact grid": np.radians(deg),
        "drifted": np.radians(deg) + drift,
        "walk": None,
        "off grid": rng.uniform(-1.6, 7.9, n),
    }
    rot = np.empty(n); r = 37.0
    steps = rng.choice([0.6, -0.6, 1.2, -1.2], n)
    for i in range(n):           # the simulator's own heading arithmetic (MyUtils.py:279)
        r = (r + steps[i] + 720.0) % 360.0
        rot[i] = r
    fams["walk"] = np.concatenate([np.radians(rot), np.radians(360.0 - rot), np.radians(rot + 90.0), np.radians(rot - 90.0),
                                   np.radians((rot + 45.0 + 720.0) % 360.0)])
    for name, x in fams.items():
        s, c = sincos(x)
        s2, c2 = sincos(x, dd_only=True)
        for got, dd, fn in ((s, s2, mp.sin), (c, c2, mp.cos)):
            bad = np.nonzero(got != dd)[0]
            assert len(bad) <= 1e-5 * len(x), (name, len(bad))
            if name == "off grid":
                assert len(bad) == 0
            for i in bad:
                assert got[i] == float(fn(mp.mpf(float(x[i])))), (name, x[i])
    # and directly against mpmath on a sample of the simulator's own arguments
    x = fams["walk"][:: 97]
    s, c = sincos(x)
    assert all(s[i] == float(mp.sin(mp.mpf(float(x[i])))) and c[i] == float(mp.cos(mp.mpf(float(x[i])))) for i in range(len(x)))
