"""GPU parity of the RegNetInference front end through the C ABI: bit-exact against the
reference's own outputs (tests/golden/regnet.npz) and against the pinned oracle on seeded inputs."""
import numpy as np
import pytest

from oracle import regnet_port as rp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from scrutiny_b200 import engine
    e = engine.Engine(0, "f32")
    yield e
    e.close()


@pytest.mark.parametrize("name", ["a", "b", "c", "d"])
def test_rank_transform_matches_reference_golden(eng, golden, name):
    g = golden("regnet")
    m = g["in_" + name].copy()
    eng.rank_transform(m)
    assert np.array_equal(m, g["s_" + name])


def test_rank_transform_counting_and_sorting_paths_against_oracle(eng):
    rs = np.random.RandomState(3)
    n, cells = 300, 1500
    m = rs.poisson(np.exp(rs.normal(0.0, 1.5, size=(n, 1))), size=(n, cells)).astype(np.float64)
    m[:, 7] += 0.5                    # one non-integer cell -> sorting path
    m[3, 11] = 5000.0                 # one value beyond the histogram range -> sorting path
    m[:, 20] = 0.0                    # an all-tied cell
    m[17, :] = 2.0                    # an all-tied gene (before the first pass)
    want = rp.convert_to_s(n, cells, m.copy())
    got = eng.rank_transform(m.copy())
    st = eng.rank_transform_stats()
    assert st["counting_cells"] == cells - 2 and st["sorted_cells"] == 2
    assert np.array_equal(got, want)
    c = rs.standard_normal(size=(n, cells))
    c[:, 5] = c[:, 6]
    c[40:60, 9] = c[39, 9]
    assert np.array_equal(eng.rank_transform(c.copy()), rp.convert_to_s(n, cells, c.copy()))
    assert eng.rank_transform_stats()["sorted_cells"] == cells


def test_rank_transform_edges(eng):
    from scrutiny_b200.engine import ScrutinyError
    rs = np.random.RandomState(4)
    for n, cells in [(1, 1), (1, 9), (9, 1), (2, 2), (129, 3), (4096, 5), (4097, 4), (8192, 3)]:
        m = rs.randint(0, 4, size=(n, cells)).astype(np.float64)
        assert np.array_equal(eng.rank_transform(m.copy()), rp.convert_to_s(n, cells, m.copy())), (n, cells)
        c = rs.standard_normal(size=(n, cells)).round(1)
        assert np.array_equal(eng.rank_transform(c.copy()), rp.convert_to_s(n, cells, c.copy())), (n, cells)
    # strided rows (ld > cells): only the view is touched
    big = rs.poisson(2.0, size=(50, 80)).astype(np.float64)
    ref = big.copy()
    view = big[:, 10:50]
    eng.rank_transform(view)
    assert np.array_equal(view, rp.convert_to_s(50, 40, ref[:, 10:50].copy()))
    assert np.array_equal(big[:, :10], ref[:, :10]) and np.array_equal(big[:, 50:], ref[:, 50:])
    with pytest.raises(ScrutinyError):
        eng.rank_transform(np.zeros((8193, 2)))


def test_rank_transform_bench_shape_properties(eng):
    """n = 3000 genes x 20000 cells of Poisson counts (the shape MatriSeq writes): exact against the oracle on a
    sample of genes and cells, and the rank-sum checksum on every gene."""
    rs = np.random.RandomState(6)
    n, cells = 3000, 20000
    lam = np.exp(rs.normal(0.0, 1.2, size=(n, 1)))
    m = rs.poisson(lam, size=(n, cells)).astype(np.float64)
    got = eng.rank_transform(m.copy())
    # every gene row holds a permutation-with-ties of ranks 1..cells: sum r = cells (cells + 1) / 2
    r = (got + 1.0) * 0.5 * cells - 1.0
    assert np.allclose(r.sum(axis=1), cells * (cells + 1) / 2, rtol=0, atol=1e-3)
    first = np.empty_like(m)
    for c in range(cells):
        first[:, c] = rp.average_ranks_fast(m[:, c])
    for g in rs.choice(n, size=40, replace=False):
        want = 2 * ((rp.average_ranks_fast(2 * ((first[g] + 1) / n) - 1) + 1) / cells) - 1
        assert np.array_equal(got[g], want)


def test_cell_pairs_match_reference_golden(eng, golden):
    g = golden("regnet")
    S = g["s_c"]
    for t in g["thresh"]:
        want = rp.find_cell_pairs(S, g["pt"], g["types"], t)
        got = eng.cell_pairs(S, g["pt"], g["types"], t)
        assert got.tolist() == want
        assert rp.cv_cost(t, g["W"], S, g["pt"], g["types"], got.tolist()) == g["cv_cost"][list(g["thresh"]).index(t)]


def test_cell_pairs_against_oracle(eng):
    rs = np.random.RandomState(8)
    n, cells = 40, 2500
    S = rp.convert_to_s(n, cells, rs.poisson(1.0, size=(n, cells)).astype(np.float64))
    S[:, 100] = S[:, 200]
    S[:, 300] = S[:, 200]
    pt = np.round(rs.rand(cells) * 3.0, 3)
    types = rs.randint(-1, 5, size=cells).astype(np.int64)     # -1: matched by no type loop
    types[[100, 200, 300]] = 2
    pt[100], pt[200], pt[300] = 1.0, 1.001, 1.002
    for thresh in (0.0, 0.004, 0.05):
        want = rp.find_cell_pairs_fast(S, pt, types, thresh)
        got = eng.cell_pairs(S, pt, types, thresh)
        assert got.shape == (len(want), 2) and got.tolist() == want
    assert [200, 100] not in want and [300, 200] not in want
    assert eng.cell_pairs(S, pt, np.full(cells, -1), 1.0).shape == (0, 2)
    one = eng.cell_pairs(S[:, :1].copy(), pt[:1], np.zeros(1, dtype=np.int64), 1.0)
    assert one.shape == (0, 2)


def test_module_surface(eng, tmp_path):
    """readCellStates / convertToS / findCellPairs of scrutiny_b200.RegNetInference on artefacts written the
    way MatriSeq.saveData writes them."""
    import scrutiny_b200.RegNetInference as rni
    rs = np.random.RandomState(9)
    n, cells = 30, 200
    counts = rs.poisson(1.3, size=(n, cells)).astype(np.float64)
    pt = 0.01 * rs.randint(0, 400, size=cells)
    types = rs.randint(0, 3, size=cells).astype(np.int64)
    gc = rs.standard_normal((n, n)) * (rs.rand(n, n) < 0.1)
    np.save(tmp_path / "synthscRNAseq.npy", counts)
    np.save(tmp_path / "synthPseudotimes.npy", pt)
    np.save(tmp_path / "cellTypesRecord.npy", types)
    np.save(tmp_path / "gc.npy", gc)
    rni.init(inputDir=str(tmp_path), outputDir=None, verbose=False)
    n2, c2, S, pt2, ty2, W = rni.readCellStates()
    assert (n2, c2) == (n, cells) and np.array_equal(W, gc)
    assert np.array_equal(S, rp.convert_to_s(n, cells, counts.copy()))
    assert rni.findCellPairs(S, pt2, ty2, 0.05) == rp.find_cell_pairs_fast(S, pt2, ty2, 0.05)
    with pytest.raises(ValueError):
        rni.convertToS(n + 1, cells, counts)
