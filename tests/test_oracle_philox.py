"""The generator restated in oracle/philox_ref.py (what the GPU tests compare the device words and the free-running noise
with) pinned on the CPU: Random123's known-answer vectors for Philox4x32-10, and the distribution of both Box-Muller maps
and of the sparse stepper's noise stream v2 (moments, Kolmogorov-Smirnov, tails, independence across cells / steps / genes)."""
import numpy as np
from scipy import stats

from oracle import philox_ref as pr


def test_philox_known_answer_vectors():
    # Random123 kat_vectors, "philox4x32 10"
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = pr.philox4x32_10(*[np.array([c], dtype=np.uint32) for c in ctr], *key)
        assert tuple(int(v[0]) for v in got) == want
    # vectorised call == element by element
    rng = np.random.RandomState(1)
    c = rng.randint(0, 2 ** 32, size=(4, 64), dtype=np.uint64).astype(np.uint32)
    all_at_once = np.stack(pr.philox4x32_10(c[0], c[1], c[2], c[3], 7, 9))
    for i in (0, 17, 63):
        one = pr.philox4x32_10(c[0, i:i + 1], c[1, i:i + 1], c[2, i:i + 1], c[3, i:i + 1], 7, 9)
        assert [int(v[0]) for v in one] == [int(v) for v in all_at_once[:, i]]


def _words(count, seed):
    i = np.arange(count, dtype=np.uint64)
    return pr.philox4x32_10(i & np.uint64(0xFFFFFFFF), i >> np.uint64(32), 5, 0, seed, 0)


def test_both_box_muller_maps_give_standard_normals():
    w = _words(250000, 11)
    a, b = pr.box_muller(w[0], w[1], 1.0)                    # 23-bit uniforms (dense stepper, count kernel)
    c, d = pr.box_muller16(w[2], 1.0)                        # 16-bit halves of one word (sparse stepper, stream v2)
    e, f = pr.box_muller16(w[3], 1.0)
    for z in (np.concatenate([a, b]), np.concatenate([c, d, e, f])):
        assert abs(z.mean()) < 4 / np.sqrt(z.size) and abs(z.std() - 1) < 4e-3
        assert abs(stats.skew(z)) < 0.02 and abs(stats.kurtosis(z)) < 0.04
        assert stats.kstest(z, "norm").pvalue > 1e-3
        assert abs((np.abs(z) > 3).mean() - 2 * stats.norm.sf(3)) < 3e-4          # tail mass beyond 3 sigma
    assert abs(np.corrcoef(a, b)[0, 1]) < 0.01 and abs(np.corrcoef(c, d)[0, 1]) < 0.01 and abs(np.corrcoef(c, e)[0, 1]) < 0.01
    # the 16-bit radius ends at sqrt(2 ln 65536) = 4.71 sigma: its second moment is short of 1 by 2^-16 (1 + ln 65536) / ... < 2e-4
    k = np.arange(65536, dtype=np.float64)
    assert abs(np.mean(-2.0 * np.log(1.0 - k / 65536.0)) / 2.0 - 1.0) < 2e-4
    assert np.max(np.hypot(c, d)) <= np.sqrt(2 * np.log(65536.0)) + 1e-12


def test_stepper_noise_stream_is_a_pure_function_with_independent_draws():
    n_pad, sigma, seed = 512, 0.05, 0xABCDEF0123
    perm = np.arange(n_pad)
    perm[500:] = -1                                           # 500 genes, 12 padding positions
    base = pr.stepper_noise(7, 3, n_pad, perm, sigma, seed)
    assert base.shape == (500,) and np.array_equal(base, pr.stepper_noise(7, 3, n_pad, perm, sigma, seed))
    assert len(np.unique(base)) == 500                        # no value handed to two genes
    other = [pr.stepper_noise(8, 3, n_pad, perm, sigma, seed), pr.stepper_noise(7, 4, n_pad, perm, sigma, seed),
             pr.stepper_noise(7, 3, n_pad, perm, sigma, seed + 1), pr.stepper_noise(7, 3, n_pad, perm, sigma, seed, stream=1),
             pr.stepper_noise(7 + 2 ** 32, 3, n_pad, perm, sigma, seed)]
    for o in other:
        assert abs(np.corrcoef(base, o)[0, 1]) < 0.2 and not np.array_equal(base, o)
    # a gene order is only a relabelling: the value belongs to the internal position
    shuffled = np.random.RandomState(0).permutation(500)
    perm2 = perm.copy()
    perm2[:500] = shuffled
    assert np.array_equal(pr.stepper_noise(7, 3, n_pad, perm2, sigma, seed)[shuffled], base)
    rows = np.stack([pr.stepper_noise(c, s, n_pad, perm, sigma, seed) for c in range(60) for s in range(20)]) / sigma
    z = rows.ravel()
    assert abs(z.mean()) < 4 / np.sqrt(z.size) and abs(z.std() - 1) < 6e-3 and stats.kstest(z, "norm").pvalue > 1e-3
    # draws that share a Philox call or a word are uncorrelated: the two lanes of a pair (words 0, 1 against 2, 3), the cos and
    # sin branch of one word (genes 32 apart), the two words of a lane's half (genes 64 apart)
    for g0, g1 in ((0, 1), (0, 32), (0, 64), (5, 37), (130, 131)):
        assert abs(np.corrcoef(rows[:, g0], rows[:, g1])[0, 1]) < 0.12, (g0, g1)       # 1200 samples: 4 sigma = 0.115
