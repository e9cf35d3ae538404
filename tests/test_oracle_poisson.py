"""Pin oracle/poisson_oracle.c against numpy's legacy np.random.poisson (what MatriSeq.py:884
calls): with the uniforms the legacy generator itself would consume (random_sample of the same
seed), the restated sampler must return the same integers.  numpy's PTRS uses its own log-gamma,
ours uses libm lgamma, so a handful of borderline acceptances may differ for lam >= 10; below 10
the match must be exact."""
import numpy as np

from oracle import poisson_ref


def test_multiplication_branch_matches_numpy_exactly():
    lam = np.random.RandomState(1).gamma(0.8, 2.5, size=20000)
    lam = lam[lam < 10.0]
    want = np.array([np.random.RandomState(77 + 0).poisson(1.0)])  # warm import path
    rs = np.random.RandomState(5)
    want = np.array([rs.poisson(l) for l in lam])
    u = np.random.RandomState(5).random_sample(int(want.sum() + len(lam) + 16))
    got, used = poisson_ref.poisson_stream(lam, u)
    assert used == want.sum() + len(lam)          # k+1 uniforms per draw
    assert np.array_equal(got, want)


def test_ptrs_branch_matches_numpy():
    lam = np.random.RandomState(2).uniform(10.0, 400.0, size=5000)
    rs = np.random.RandomState(9)
    want = np.array([rs.poisson(l) for l in lam])
    u = np.random.RandomState(9).random_sample(8 * len(lam))
    got, used = poisson_ref.poisson_stream(lam, u)
    assert used > 0
    assert np.mean(got == want) > 0.999            # identical stream unless an acceptance test ties
    if not np.array_equal(got, want):
        first = np.nonzero(got != want)[0][0]
        assert first > 100                          # streams stay in lock-step for a long prefix


def test_zero_and_negative_lambda():
    got, used = poisson_ref.poisson_batch(np.array([0.0, -3.0, np.nan]), np.full((3, 4), 0.5))
    assert list(got) == [0, 0, 0] and list(used) == [0, 0, 0]


def test_out_of_uniforms_is_flagged():
    got, used = poisson_ref.poisson_batch(np.array([50.0]), np.full((1, 1), 0.5))
    assert got[0] == -1


def test_extension_oracle_gamma_poisson_and_dropout_distributions():
    """oracle_counts_ext_batch (NB via Marsaglia-Tsang Gamma x Poisson, Bernoulli dropout; default-off extensions of the
    read-out): with both switches off it IS the Poisson oracle; switched on, its output follows scipy's negative binomial /
    the zero-inflated law (chi-square on the pmf), for shapes below and above 1."""
    from scipy import stats
    from oracle import poisson_ref
    rng = np.random.RandomState(3)
    entries = 200000
    U = rng.random_sample((entries, 96))
    for lam0 in (0.7, 6.0, 40.0):
        lam = np.full(entries, lam0)
        base, used0 = poisson_ref.poisson_batch(lam, U)
        same, used1 = poisson_ref.counts_ext_batch(lam, U, 0.0, 0.0)
        assert np.array_equal(base, same) and np.array_equal(used0, used1)
        for phi, p in ((0.4, 0.0), (2.5, 0.0), (0.0, 0.25), (0.4, 0.25)):
            got, used = poisson_ref.counts_ext_batch(lam, U, phi, p)
            assert (got >= 0).all() and used.max() <= 96
            k = np.arange(got.max() + 1)
            pmf = stats.nbinom.pmf(k, 1.0 / phi, 1.0 / (1.0 + phi * lam0)) if phi > 0 else stats.poisson.pmf(k, lam0)
            pmf = (1 - p) * pmf
            pmf[0] += p
            obs = np.bincount(got, minlength=len(k))
            exp = pmf * entries
            m = exp > 8
            chi = ((obs[m] - exp[m]) ** 2 / exp[m]).sum()
            assert stats.chi2.sf(chi, m.sum() - 1) > 1e-4, (lam0, phi, p, chi)
