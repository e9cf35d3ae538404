"""oracle/regnet_port.py must reproduce, bit for bit, what the unmodified reference's RegNetInference produced
(tests/golden/regnet.npz, made by oracle/make_golden.py): this pins the oracle for SURVEY section 8 row f4."""
import numpy as np
import pytest

from oracle import regnet_port as rp


@pytest.mark.parametrize("name", ["a", "b", "c", "d"])
@pytest.mark.parametrize("ranks", [rp.average_ranks, rp.average_ranks_fast])
def test_convert_to_s_matches_reference(golden, name, ranks):
    g = golden("regnet")
    m = g["in_" + name]
    got = rp.convert_to_s(m.shape[0], m.shape[1], m.copy(), ranks=ranks)
    assert np.array_equal(got, g["s_" + name])


def test_pairs_match_reference_through_cv_cost(golden):
    """getCVCost (RNI:755-797) returns only the mean squared error over the pairs it finds; equal values for
    three thresholds pin the pair set (the identical cells 5 and 9 must be skipped)."""
    g = golden("regnet")
    S = g["s_c"]
    for t, want in zip(g["thresh"], g["cv_cost"]):
        pairs = rp.find_cell_pairs(S, g["pt"], g["types"], t)
        assert pairs == rp.find_cell_pairs_fast(S, g["pt"], g["types"], t)
        assert [9, 5] not in pairs and [5, 9] not in pairs
        assert rp.cv_cost(t, g["W"], S, g["pt"], g["types"], pairs) == want


def test_live_reference_when_present():
    from oracle import reference_harness as rh
    if not rh.available():
        pytest.skip("reference tree not present (GPU box)")
    rni = rh.load_reference_rni()
    rs = np.random.RandomState(5)
    m = rs.poisson(0.7, size=(25, 40)).astype(np.float64)
    assert np.array_equal(rni.convertToS(25, 40, m.copy()), rp.convert_to_s(25, 40, m.copy()))
