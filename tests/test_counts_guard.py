"""Host-side check of the error budget behind the count kernel's float32 screening (csrc/aux_kernels.cuh,
poisson_small_f32): the multiplication method of numpy's legacy Poisson sampler (MS:884) is evaluated with product and
threshold in float32, and a comparison is only taken when the product lies outside threshold * (1 +- 1e-4); everything
else is redone in float64 on the device.  Here the float32 arithmetic is replayed with numpy float32 operations, the two
ex2.approx evaluations are modelled as the correctly rounded result perturbed by up to +-2 ulp (the PTX bound), and the
outcome is compared with the C oracle on the same Philox-style uniforms: whenever the screening answers, it must answer
what the float64 sampler answers, and the log-distance between the two (product / threshold) ratios must stay far inside
the guard.  (The kernel itself is compared entry for entry with the float64 sampler in tests/test_gpu_parity.py.)"""
import numpy as np

from oracle import poisson_ref

GUARD = np.float32(1e-4)
LOG2E32 = np.float32(1.4426950408889634)


def ex2_model(t, rng):
    """float32 2^t as ex2.approx.ftz delivers it in the worst case: correctly rounded, then moved by up to 2 ulp."""
    with np.errstate(over="ignore", invalid="ignore"):   # arguments outside the screened range overflow; they are declined
        y = np.exp2(t.astype(np.float64)).astype(np.float32)
        ulp = np.spacing(y)
        return (y + rng.randint(-2, 3, size=y.shape).astype(np.float32) * ulp).astype(np.float32)


def screen(x32, shift, size, words, rng, exp_scale=1.0):
    """Replays counts_kernel<T, FAST=true>'s screening.  Returns (count or -1, log-ratio error at the decision)."""
    esl = exp_scale * 1.4426950408889634
    t2 = ((x32.astype(np.float64) + shift) * esl).astype(np.float32)
    sf32 = size.astype(np.float32)
    ok = (size >= 0.05) & (size <= 8.0) & (t2 > np.float32(-64)) & (t2 < np.float32(8))
    lam_f = (ex2_model(t2, rng) * sf32).astype(np.float32)
    ok &= lam_f < np.float32(9.999)
    enlam = ex2_model((lam_f * (-LOG2E32)).astype(np.float32), rng)
    hi = (enlam * (np.float32(1) + GUARD)).astype(np.float32)
    lo = (enlam * (np.float32(1) - GUARD)).astype(np.float32)
    # one FMA on the device: a single rounding of the exact (v_rounded * 2^-32 + 2^-33)
    u = (words.astype(np.float32).astype(np.float64) * 2.0 ** -32 + 2.0 ** -33).astype(np.float32)
    prod = np.cumprod(u, axis=1, dtype=np.float32)                  # sequential float32 multiplies
    above = prod > hi[:, None]
    first = np.where(above.all(axis=1), -1, np.argmin(above, axis=1))
    pf = prod[np.arange(len(prod)), np.maximum(first, 0)]
    out = np.where((first >= 0) & (pf < lo), first, -1)
    out = np.where(ok, out, -1)
    return out, lam_f, enlam, prod


def test_screening_never_disagrees_with_the_float64_sampler():
    rng = np.random.RandomState(2024)
    entries, draws = 400000, 64
    # lambda spread over the whole screened range, dense near the critical ends (tiny and just below 10)
    loglam = np.concatenate([rng.uniform(np.log(1e-9), np.log(9.99), entries // 2),
                             rng.uniform(np.log(3.0), np.log(9.9989), entries // 4),
                             np.log(rng.gamma(0.35, 1 / 0.19, entries // 4))])
    size = rng.normal(1, 0.14, entries) ** 3
    size[::1000] = rng.choice([0.01, 0.049, 8.5, -0.3, 0.0], size=len(size[::1000]))     # outside the screened range
    x32 = np.tanh(rng.randn(entries)).astype(np.float32)
    shift = loglam - np.log(np.abs(size) + 1e-300) - x32.astype(np.float64)
    words = rng.randint(0, 2 ** 32, size=(entries, draws), dtype=np.uint64).astype(np.uint32)
    words[::777, 0] = rng.choice([0, 1, 2 ** 32 - 1, 2 ** 32 - 200, 511], size=len(words[::777, 0]))   # extreme uniforms

    got, lam_f, enlam_f, prod_f = screen(x32, shift, size, words, rng)

    # the float64 sampler on the same uniforms (C oracle; lambda exactly as the exact device path forms it)
    lam = np.maximum(np.exp(1.0 * (x32.astype(np.float64) + shift)) * size, 0.0)
    U = (words.astype(np.float64) + 0.5) / 2.0 ** 32
    want, used = poisson_ref.poisson_batch(lam, U)
    answered = got >= 0
    assert answered.mean() > 0.97                                     # lambda >= 9.999 and odd sizes decline, little else
    assert (lam[answered] < 10.0).all() and (lam[answered] > 0).all()
    assert np.array_equal(got[answered], want[answered])
    small = (lam > 0) & (lam < 9.999) & (size >= 0.05) & (size <= 8.0)
    declined = small & ~answered
    assert declined.sum() / small.sum() < 1e-3                        # ~2e-4 expected: the guard band is cheap

    # the error budget itself: (prod_f / enlam_f) against (prod / enlam) in float64 at every draw that is compared
    sel = np.nonzero(small & (want >= 0))[0]
    k = want[sel]
    p64 = np.cumprod(U[sel], axis=1)[np.arange(len(sel)), k]
    ratio64 = np.log(p64) + lam[sel]
    ratio32 = np.log(prod_f[sel, k].astype(np.float64)) - np.log(enlam_f[sel].astype(np.float64))
    assert np.max(np.abs(ratio32 - ratio64)) < 2.5e-5                 # bound derived in aux_kernels.cuh: 2e-5; guard 1e-4
    assert np.max(np.abs(lam_f[sel].astype(np.float64) - lam[sel])) < 6.8e-6


def test_declined_inputs():
    """Inputs outside the screened ranges are never answered (the device then runs the float64 sampler)."""
    rng = np.random.RandomState(1)
    words = rng.randint(0, 2 ** 32, size=(6, 64), dtype=np.uint64).astype(np.uint32)
    x32 = np.zeros(6, dtype=np.float32)
    shift = np.array([0.0, 0.0, 0.0, np.log(9.9995), -60.0, np.nan])
    size = np.array([0.04, 8.1, -1.0, 1.0, 1.0, 1.0])
    got, *_ = screen(x32, shift, size, words, rng)
    assert (got == -1).all()


# ---------------------------------------------------------------------------------------------------------------------
# float32 screening of the PTRS sampler (lambda >= 10): aux_kernels.cuh poisson_ptrs_f32
# ---------------------------------------------------------------------------------------------------------------------
F = np.float32


def approx(y, rng, ulps=2):
    """a MUFU result in the worst case: the correctly rounded float32 value moved by up to `ulps` ulp"""
    y = y.astype(np.float32)
    return (y + rng.randint(-ulps, ulps + 1, size=y.shape).astype(np.float32) * np.spacing(np.abs(y))).astype(np.float32)


def lg2_model(v, rng):
    """lg2.approx.ftz: correctly rounded +- 2 ulp and +- 2^-22 absolute"""
    with np.errstate(divide="ignore"):
        y = np.log2(v.astype(np.float64))
    return (approx(y, rng) + (rng.randint(-1, 2, size=v.shape) * 2.0 ** -22).astype(np.float32)).astype(np.float32)


def fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + np.asarray(c, dtype=np.float64)).astype(np.float32)


def ptrs_screen(lam, words, rng, lgam):
    """Replays poisson_ptrs_f32 (vectorised over entries; draws in sequence).  Returns count or -1."""
    n = len(lam)
    lam = lam.astype(np.float32)
    slam = approx(np.sqrt(lam.astype(np.float64)), rng, 1)
    b = fma32(F(2.53) * np.ones(n, F), slam, F(0.931))
    a = fma32(F(0.02483) * np.ones(n, F), b, F(-0.059))
    vr = fma32(F(-3.6224) * np.ones(n, F), approx(1.0 / (b - F(2.0)).astype(np.float64), rng, 1), F(0.9277))
    g0 = fma32(F(2e-6) * np.ones(n, F), lam, F(3e-4))
    a2 = (a + a).astype(F)
    out = np.full(n, -2, dtype=np.int64)          # -2 = still drawing
    ln2 = F(0.69314718055994531)
    for d in range(16):
        live = out == -2
        if not live.any():
            break
        u = fma32(words[:, 2 * d].astype(np.float32), F(2.3283064365386963e-10) * np.ones(n, F), F(1.16415321826934814e-10))
        V = fma32(words[:, 2 * d + 1].astype(np.float32), F(2.3283064365386963e-10) * np.ones(n, F), F(1.16415321826934814e-10))
        U = (u - F(0.5)).astype(F)
        us = (F(0.5) - np.abs(U)).astype(F)
        decline = (np.abs(us - F(0.07)) < F(1e-6)) | (np.abs(us - F(0.013)) < F(1e-6))
        rus = approx(1.0 / us.astype(np.float64), rng, 1)
        q = (a2 * rus).astype(F)
        x = fma32((q + b).astype(F), U, (lam + F(0.43)).astype(F))
        g = fma32(q, fma32(F(2e-7) * np.ones(n, F), rus, F(1e-5)), g0)
        kf = np.floor(x)
        fr = (x - kf).astype(F)
        k_exact = (fr > g) & (fr < F(1.0) - g)
        res = np.full(n, -2, dtype=np.int64)      # -2 continue drawing, -1 decline, >= 0 answer
        undecided = ~decline
        res[decline] = -1
        fast_zone = undecided & (us > F(0.07))
        acc = fast_zone & (V < vr - F(2e-6))
        res[acc] = np.where(k_exact[acc], kf[acc].astype(np.int64), -1)
        amb = fast_zone & ~acc & (V < vr + F(2e-6))
        res[amb] = -1
        rest = undecided & ~acc & ~amb
        neg = rest & (x < g)
        res[neg & ~(x < -g)] = -1                 # sign of kf uncertain
        rest &= ~neg                              # x < -g: next draw (res stays -2)
        sq = rest & (us < F(0.013))
        nxt = sq & (V > us + F(1e-6))
        res[sq & ~nxt & (V > us - F(1e-6))] = -1
        rest &= ~nxt & (res == -2)
        bad = rest & (~k_exact | (kf >= len(lgam)))
        res[bad] = -1
        rest &= ~bad
        if rest.any():
            r = np.nonzero(rest)[0]
            inva = fma32(F(1.1328) * np.ones(len(r), F), approx(1.0 / (b[r] - F(3.4)).astype(np.float64), rng, 1), F(1.1239))
            inner = fma32(a[r], (rus[r] * rus[r]).astype(F), b[r])
            lhs = ((lg2_model(V[r], rng) + lg2_model(inva, rng) - lg2_model(inner, rng)).astype(F) * ln2).astype(F)
            rhs = (fma32(kf[r].astype(F), (lg2_model(lam[r], rng) * ln2).astype(F), -lam[r]) - lgam[kf[r].astype(np.int64)]).astype(F)
            G = (fma32(F(4e-6) * np.ones(len(r), F), kf[r].astype(F), F(3e-3)) + F(2e-7) * approx(1.0 / V[r].astype(np.float64), rng, 1)
                 + F(1e-6) * rus[r]).astype(F)
            res[r] = np.where(lhs < rhs - G, kf[r].astype(np.int64), np.where(lhs > rhs + G, -2, -1))
        out[live] = res[live]
    out[out == -2] = -1
    return out


def test_ptrs_screening_never_disagrees_with_the_float64_sampler():
    from math import lgamma
    rng = np.random.RandomState(77)
    entries, pairs = 300000, 16
    lam = np.concatenate([rng.uniform(10.0011, 40.0, entries // 2), np.exp(rng.uniform(np.log(10.0011), np.log(999.0), entries // 2))])
    lam_f = (lam * (1.0 + rng.uniform(-8e-7, 8e-7, entries))).astype(np.float32)     # the kernel's float32 rate: within 1e-6
    words = rng.randint(0, 2 ** 32, size=(entries, 2 * pairs), dtype=np.uint64).astype(np.uint32)
    words[::501, 0] = rng.choice([0, 1, 2 ** 32 - 1, 2 ** 31, 2 ** 31 - 1, 300647710, 55834574], size=len(words[::501, 0]))  # us ~ 0, 0.5, 0.07, 0.013
    lgam = np.array([lgamma(k + 1.0) for k in range(2048)]).astype(np.float32)
    got = ptrs_screen(lam_f, words, rng, lgam)
    U = (words.astype(np.float64) + 0.5) / 2.0 ** 32
    want, used = poisson_ref.poisson_batch(lam, U)
    answered = got >= 0
    assert answered.mean() > 0.97, answered.mean()
    ok = want >= 0                                                   # the oracle itself ran out of the 32 uniforms: ignore
    assert np.array_equal(got[answered & ok], want[answered & ok])
    assert (answered & ~ok).sum() == 0
