"""TEST INFRASTRUCTURE ONLY (oracle) -- numpy restatement of the float32 update of the sparse stepper
(scrutiny_b200/csrc/step_kernel.cuh ``next_state``), the arithmetic form of developCell (MatriSeq.py:529-535) it uses:

    m = K z,  K = 2 log2(e)           (operator entries, bias and noise radius are pre-scaled by K on the host)
    r = 1 / (1 + 2^m)                 (tanh z = 1 - 2 r)
    x' = proj * fma(r, -2 dt, fma(x, 1 - dt, dt)) + const

``perturb`` models the device's approximate ex2 / rcp units: each result is multiplied by (1 + e), |e| <= ``ulps`` * 2^-23.
tests/test_stepper_update_cpu.py bounds the distance to the reference's float64 expression.
"""
import numpy as np

K = np.float32(2.88539008177792681472)


def next_state_f32(z, x, dt, proj, const, ulps=0.0, rng=None):
    f = np.float32
    z, x, proj, const = (np.asarray(a, dtype=f) for a in (z, x, proj, const))
    dt = f(dt)

    def perturb(v):
        if not ulps:
            return v
        e = rng.uniform(-ulps, ulps, size=v.shape) * 2.0 ** -23
        return (v.astype(np.float64) * (1.0 + e)).astype(f)

    with np.errstate(over="ignore"):
        m = (K * z).astype(f)
        e2 = perturb(np.exp2(m).astype(f))
        r = perturb((f(1) / (e2 + f(1))).astype(f))
    inner = (x.astype(np.float64) * np.float64(f(1) - dt) + np.float64(dt)).astype(f)          # fma: one rounding
    mid = (r.astype(np.float64) * np.float64(f(-2) * dt) + inner.astype(np.float64)).astype(f)
    return (proj.astype(np.float64) * mid.astype(np.float64) + const.astype(np.float64)).astype(f)


def next_state_f64(z, x, dt, proj, const):
    """MS:529-535 for one element, float64."""
    return proj * (x + dt * (np.tanh(z) - x)) + const
