"""The float32 arithmetic form of developCell the sparse stepper uses (oracle/stepper_f32_ref.py restates
step_kernel.cuh next_state) against the reference's float64 expression (MS:529-535): one teacher-forced step stays within
the 5e-7 the GPU parity tests allow, also with the approximate ex2 / rcp units modelled as +-2 ulp, and the form is exact in
both saturation limits (no NaN from 2^m overflowing)."""
import numpy as np

from oracle import stepper_f32_ref as ref


def test_one_step_error_of_the_float32_form():
    rng = np.random.RandomState(0)
    k = 400000
    z = np.concatenate([rng.uniform(-12, 12, k // 2), rng.normal(0, 1.5, k // 2)])
    x = rng.uniform(-1, 1, k)
    proj = (rng.rand(k) > 0.05).astype(float)
    const = np.where(proj == 0, np.sign(rng.rand(k) - 0.5), 0.0)
    for dt in (0.01, 0.05):
        want = ref.next_state_f64(z.astype(np.float32).astype(np.float64), x.astype(np.float32).astype(np.float64), dt, proj, const)
        exact_units = ref.next_state_f32(z, x, dt, proj, const)
        assert np.max(np.abs(exact_units - want)) <= 1.5e-7            # the state's own float32 rounding dominates
        approx_units = ref.next_state_f32(z, x, dt, proj, const, ulps=2.0, rng=rng)
        assert np.max(np.abs(approx_units - want)) <= 2.0e-7
        # the tanh part alone: |(1 - 2 r) - tanh z| with +-2 ulp units, before it is scaled by dt
        f = np.float32
        m = (ref.K * z.astype(f)).astype(f)
        r = (f(1) / (np.exp2(m).astype(f) + f(1))).astype(np.float64)
        assert np.max(np.abs((1.0 - 2.0 * r) - np.tanh(z.astype(f).astype(np.float64)))) <= 4e-7


def test_saturation_limits_are_exact():
    z = np.array([-1e4, -200.0, -60.0, 60.0, 200.0, 1e4, 3e38, -3e38])
    x = np.full(len(z), 0.25)
    got = ref.next_state_f32(z, x, 0.01, np.ones(len(z)), np.zeros(len(z)))
    want = ref.next_state_f64(z, x, 0.01, 1.0, 0.0)
    assert np.isfinite(got).all() and np.max(np.abs(got - want)) <= 6e-8
