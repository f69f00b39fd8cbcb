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
