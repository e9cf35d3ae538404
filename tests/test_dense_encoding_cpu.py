"""The operand encodings of the dense stepper, restated in numpy (oracle/dense_split_ref.py): how far the three-product
split-precision form of X W^T is from the float64 product -- the representation error that DESIGN.md section 3.2 quotes
(about 22 significant bits per operand for both encodings, the fp16 split no worse than 3xTF32)."""
import numpy as np
import pytest

from oracle import dense_split_ref as ref


def c4_like(rng, n, cells, alpha=2.21):
    W = np.stack([rng.beta(rng.randint(1, 11), rng.randint(1, 11), size=n) * rng.choice([-1.0, 1.0]) for _ in range(n)])
    W *= rng.rand(n, n) < 0.5
    X = rng.uniform(-1, 1, size=(cells, n))
    return X, W / alpha


@pytest.mark.parametrize("kind", ["f16", "tf32"])
def test_three_products_carry_fp32_class_accuracy(kind):
    rng = np.random.RandomState(4)
    X, W = c4_like(rng, 600, 64)
    exact = X.astype(np.float32).astype(np.float64) @ W.astype(np.float32).astype(np.float64).T
    got = ref.product(X, W, kind)
    scale = np.abs(X).astype(np.float64) @ np.abs(W).astype(np.float64).T      # sum of |terms|: what the error is relative to
    rel = np.max(np.abs(got - exact) / scale)
    assert rel < 2.0 ** -20, rel                                               # dropped lo * lo term + lo rounding
    one = (ref.split_f16(X, ref.STATE_SCALE)[0].astype(np.float64) / ref.STATE_SCALE) @ \
        (ref.split_f16(W, ref.w_scale(W))[0].astype(np.float64) / ref.w_scale(W)).T
    assert np.max(np.abs(one - exact) / scale) > 30 * rel                      # a single 11-bit product is far worse


def test_fp16_split_is_no_worse_than_tf32_split_and_handles_small_and_large_values():
    rng = np.random.RandomState(9)
    X, W = c4_like(rng, 400, 32)
    X[:, :50] *= 1e-4                      # states near zero: lo terms in fp16's subnormal range
    W[:, 100:150] *= 1e-3
    W[7, 3] = 40.0                         # one large weight sets the scale
    exact = X.astype(np.float32).astype(np.float64) @ W.astype(np.float32).astype(np.float64).T
    scale = np.abs(X).astype(np.float64) @ np.abs(W).astype(np.float64).T + 1e-30
    e16 = np.max(np.abs(ref.product(X, W, "f16") - exact) / scale)
    e32 = np.max(np.abs(ref.product(X, W, "tf32") - exact) / scale)
    assert e16 < 2.0 ** -19 and e32 < 2.0 ** -19 and e16 < 4 * e32, (e16, e32)
    # representation of a state entry: relative 2^-23, or 2^-31 absolute where lo falls into fp16's subnormal range
    xh, xl = ref.split_f16(X, ref.STATE_SCALE)
    x32 = X.astype(np.float32).astype(np.float64)
    back = (xh.astype(np.float64) + xl.astype(np.float64)) / ref.STATE_SCALE
    assert np.all(np.abs(back - x32) <= 2.0 ** -23 * np.abs(x32) + 2.0 ** -31)


def test_scale_puts_the_largest_weight_below_fp16_max():
    for wmax in (1e-6, 0.3, 1.0, 17.0, 3e4):
        W = np.array([[wmax, -wmax / 3]])
        s = ref.w_scale(W)
        assert 2 ** 14 <= wmax * s < 2 ** 15 + 1
        hi, lo = ref.split_f16(W, s)
        assert np.isfinite(hi.astype(np.float32)).all()
