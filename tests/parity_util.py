"""Comparison rules shared by the CPU (host-emulated kernel source) and GPU parity tests.

Bar (BASELINE.json north_star): integer outputs — step counter, thrust flags, history-valid flags,
done, naughty count, raised-exception flag — bit-exact; floats within REL_TOL relative with an
ABS_TOL floor.  The north star allows 1e-4 relative per step; the kernels reproduce the reference's
IEEE operation order, so the tests hold them to 1e-9 and report how many records are bit-identical.
Known sources of last-bit differences: CUDA libm sin/cos/atan vs glibc, sqrt vs
glibc pow(x, .5), and the reference's module-global scratch rect whose low bits depend on call
history.
"""
import numpy as np

REL_TOL = 1e-9
ABS_TOL = 1e-9
INT_KEYS = ("rflag", "step")
FLOAT_KEYS = ("rob", "rhist", "ball")


def close(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    with np.errstate(invalid="ignore"):
        ok = np.abs(a - b) <= ABS_TOL + REL_TOL * np.maximum(np.abs(a), np.abs(b))
    return bool(np.all(ok | both_nan | both_inf))


def max_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    with np.errstate(invalid="ignore"):
        d = np.abs(a - b)
    d = d[np.isfinite(d)]
    return float(d.max()) if d.size else 0.0


def compare_record(got_state, got_out, want_state, want_out):
    """Returns (ok, bit_exact, max_abs_err, reason)."""
    reasons = []
    for k in INT_KEYS:
        if not np.array_equal(np.asarray(got_state[k]), np.asarray(want_state[k])):
            reasons.append(f"int:{k}")
    for k in ("done", "naughty"):
        if int(got_out[k]) != int(want_out[k]):
            reasons.append(f"int:{k}")
    worst = 0.0
    exact = not reasons
    for k in FLOAT_KEYS:
        if not close(got_state[k], want_state[k]):
            reasons.append(f"float:{k}")
        worst = max(worst, max_err(got_state[k], want_state[k]))
        exact &= np.array_equal(np.asarray(got_state[k]), np.asarray(want_state[k]))
    for k in ("obs_h", "obs_g", "rew"):
        if not close(got_out[k], want_out[k]):
            reasons.append(f"float:{k}")
        worst = max(worst, max_err(got_out[k], want_out[k]))
        exact &= np.array_equal(np.asarray(got_out[k], np.float64), np.asarray(want_out[k], np.float64), equal_nan=True)
    return (not reasons), bool(exact), worst, ",".join(reasons)
