"""CPU-side checks of the C ABI: the library loads, exports every symbol the header declares,
refuses to compute without a GPU, and the packed operator (gene permutation, sliced-ELL tiles,
long-row chunks, warp assignment) reproduces W x exactly as a plain product does."""
import os
import re

import numpy as np
import pytest

from scrutiny_b200 import _lib, engine
from scrutiny_b200.network import dense_to_csr, trrust_adjacency

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header():
    header = open(os.path.join(ROOT, "include", "scrutiny_b200.h")).read()
    declared = set(re.findall(r"\b(scr_[a-z0-9_]+)\s*\(", header))
    declared -= {"scr_ctx"}
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), "library does not export " + name
    assert declared == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert lib.scr_abi_version() == 2


def test_no_cpu_fallback():
    eng = engine.Engine(device=-1)          # packing-only context
    eng.set_network(np.eye(8) * 0.5)
    with pytest.raises(engine.ScrutinyError) as e:
        eng.alloc_cells(4)
    assert e.value.code == -2
    with pytest.raises(engine.ScrutinyError):
        eng.run_phase([0], [1], np.ones(8), np.zeros(8))


def power_law_like(n, seed, heavy=0):
    rng = np.random.RandomState(seed)
    W = np.zeros((n, n))
    deg = np.minimum((0.5 * (1 - rng.rand(n)) ** (-1 / 1.05) + 0.5).astype(int), n)
    regs = rng.choice(n, size=min(n, max(2, n // 3)), replace=False)
    for r in range(n):
        k = min(deg[r], len(regs))
        W[r, rng.choice(regs, size=k, replace=False)] = rng.randn(k)
    for r in range(heavy):
        W[rng.randint(n), :] = rng.randn(n) * (rng.rand(n) < 0.6)
    return W


@pytest.mark.parametrize("n,seed,heavy", [(1, 0, 0), (31, 1, 0), (64, 2, 1), (129, 3, 2), (500, 4, 3),
                                          (1000, 5, 0), (3000, 6, 2)])
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_packed_operator_matches_dense(n, seed, heavy, prec):
    W = power_law_like(n, seed, heavy)
    eng = engine.Engine(device=-1, precision=prec)
    eng.set_network(W)
    rng = np.random.RandomState(99)
    for _ in range(2):
        x = rng.randn(n)
        y = eng.host_matvec(x)
        ref = W @ x
        assert np.allclose(y, ref, rtol=1e-12, atol=1e-12 * max(1.0, np.abs(ref).max()))
    lay = eng.layout()
    assert lay["n_pad"] % 128 == 0 and lay["n_pad"] >= n
    assert lay["smem_bytes"] <= 232448


@pytest.mark.parametrize("n,seed,heavy", [(129, 3, 2), (1000, 5, 0), (3000, 6, 2)])
def test_bank_aware_gene_order_is_a_valid_packing(n, seed, heavy):
    """The default gene order exchanges genes inside aligned groups of 8 positions to lower the modelled shared-memory
    bank conflicts of the gather (SCR_BANK_OPT=0 switches that off).  The packed operator must still be W exactly, with
    the same block structure, and the model must not get worse."""
    W = power_law_like(n, seed, heavy)
    os.environ["SCR_BANK_OPT"] = "0"
    try:
        base = engine.Engine(device=-1)
        base.set_network(W)
    finally:
        del os.environ["SCR_BANK_OPT"]
    eng = engine.Engine(device=-1)
    eng.set_network(W)
    x = np.random.RandomState(7).randn(n)
    ref = W @ x
    assert np.allclose(eng.host_matvec(x), ref, rtol=1e-12, atol=1e-12 * max(1.0, np.abs(ref).max()))
    a, b = base.layout(), eng.layout()
    for key in ("n_pad", "pingpong_genes", "tile_slots", "chunks", "smem_bytes", "gather_wavefronts_floor"):
        assert a[key] == b[key], key
    assert b["gather_wavefronts_floor"] <= b["gather_wavefronts_model"] <= a["gather_wavefronts_model"]
    # the annealed refinement (default) is at least as good as the pairwise descent alone, and the order is a pure
    # function of the pattern: a second context (order cache hit) and a rescaled network (same pattern) get the same one
    os.environ["SCR_BANK_ANNEAL"] = "0"
    try:
        pairwise = engine.Engine(device=-1)
        pairwise.set_network(W)
    finally:
        del os.environ["SCR_BANK_ANNEAL"]
    assert b["gather_wavefronts_model"] <= pairwise.layout()["gather_wavefronts_model"]
    again = engine.Engine(device=-1)
    again.set_network(W / 3.0)
    assert np.array_equal(again.gene_order(), eng.gene_order()) and sorted(eng.gene_order()[:n]) == list(range(n))
    if n >= 1000:
        assert b["gather_wavefronts_model"] < 0.9 * a["gather_wavefronts_model"]


@pytest.mark.parametrize("n,seed,heavy", [(129, 3, 2), (1000, 5, 0), (3000, 6, 2)])
def test_hub_first_rows_are_the_same_operator_with_fewer_modelled_wavefronts(n, seed, heavy):
    """Round 2: every row lists its regulators hub first, rows of equal class and in-degree are ordered by the regulators
    they read, and entries are exchanged inside rows where the gather then touches fewer rows per bank group
    (operator_pack.h; SCR_ROW_SORT=0 keeps CSR order).  The packed operator is still W exactly, same block structure, the
    model does not get worse, the order stays a pure function of the pattern; and on the bench-shaped TRRUST adjacency the
    model drops by more than a tenth."""
    W = power_law_like(n, seed, heavy)
    os.environ["SCR_ROW_SORT"] = "0"
    try:
        base = engine.Engine(device=-1)
        base.set_network(W)
    finally:
        del os.environ["SCR_ROW_SORT"]
    eng = engine.Engine(device=-1)
    eng.set_network(W)
    x = np.random.RandomState(11).randn(n)
    ref = W @ x
    assert np.allclose(eng.host_matvec(x), ref, rtol=1e-12, atol=1e-12 * max(1.0, np.abs(ref).max()))
    a, b = base.layout(), eng.layout()
    for key in ("n_pad", "pingpong_genes", "tile_slots", "chunks", "smem_bytes"):
        assert a[key] == b[key], key
    assert b["gather_wavefronts_floor"] <= b["gather_wavefronts_model"] <= a["gather_wavefronts_model"]
    again = engine.Engine(device=-1)
    again.set_network(W * 0.5)
    assert np.array_equal(again.gene_order(), eng.gene_order()) and sorted(eng.gene_order()[:n]) == list(range(n))
    assert np.allclose(again.host_matvec(x), 0.5 * ref, rtol=1e-12, atol=1e-12 * max(1.0, np.abs(ref).max()))
    if n == 3000:
        from scrutiny_b200 import network
        import random
        np.random.seed(1337)
        random.seed(1337)
        gc, _ = network.generate_gene_correlation_matrix(0, "TRRUST")
        os.environ["SCR_ROW_SORT"] = "0"
        try:
            t0 = engine.Engine(device=-1)
            t0.set_network(gc / 0.1)
        finally:
            del os.environ["SCR_ROW_SORT"]
        t1 = engine.Engine(device=-1)
        t1.set_network(gc / 0.1)
        assert t1.layout()["gather_wavefronts_model"] < 0.9 * t0.layout()["gather_wavefronts_model"]
        xt = np.random.RandomState(3).randn(gc.shape[0])
        assert np.allclose(t1.host_matvec(xt), (gc / 0.1) @ xt, rtol=1e-12, atol=1e-10)


@pytest.mark.parametrize("n,seed,heavy", [(1, 0, 0), (64, 2, 1), (129, 3, 2), (1000, 5, 0), (3000, 6, 2)])
def test_block_records_describe_the_warp_lists(n, seed, heavy):
    """The 16-byte records the stepper walks (step_kernel.cuh): every 128-gene block exactly once, slot ranges tiling the
    operator, and the flags that replace per-block comparisons in the kernel."""
    eng = engine.Engine(device=-1)
    eng.set_network(power_law_like(n, seed, heavy))
    lay = eng.layout()
    rec, bounds = eng.block_records()
    nb = lay["n_pad"] // 128
    assert rec.shape == (nb, 4) and bounds[0] == 0 and bounds[-1] == nb and (np.diff(bounds) >= 0).all()
    assert sorted(rec[:, 0].tolist()) == [128 * b for b in range(nb)]             # each block once
    by_gene = rec[np.argsort(rec[:, 0])]
    assert by_gene[0, 1] == 0 and by_gene[-1, 2] == lay["tile_slots"]
    assert np.array_equal(by_gene[1:, 1], by_gene[:-1, 2])                        # slot ranges tile the operator
    assert ((by_gene[:, 2] - by_gene[:, 1]) % 128 == 0).all()
    pp = (rec[:, 3] & 1) != 0
    assert np.array_equal(pp, rec[:, 0] < lay["pingpong_genes"])                  # ping-ponged = read by some row
    long_blocks = sorted(rec[(rec[:, 3] & 2) != 0, 0].tolist())
    assert long_blocks == [128 * b for b in range(len(long_blocks))]              # long rows are a prefix of the order
    assert (len(long_blocks) > 0) == (lay["chunks"] > 0)
    for w in range(len(bounds) - 1):
        mine = rec[bounds[w]:bounds[w + 1]]
        mpp = (mine[:, 3] & 1) != 0
        assert not (~mpp[:-1] & mpp[1:]).any()                                    # regulator blocks first
        last = (mine[:, 3] & 4) != 0
        first = (mine[:, 3] & 8) != 0
        if mpp.any():
            assert last.sum() == 1 and last[mpp.sum() - 1]                        # arrival after the last regulator block
            assert first.sum() == 1 and first[0]                                  # wait before the first regulator store
        else:
            assert not last.any() and not first.any()
        lng = (mine[:, 3] & 2) != 0
        assert not (lng & ~mpp).any()


def test_dense_and_empty_networks():
    rng = np.random.RandomState(3)
    W = rng.randn(200, 200)                       # fully dense: no long-row chunks expected
    eng = engine.Engine(device=-1)
    eng.set_network(W)
    assert eng.layout()["chunks"] == 0
    x = rng.randn(200)
    assert np.allclose(eng.host_matvec(x), W @ x, rtol=1e-12, atol=1e-10)
    eng.set_network(np.zeros((40, 40)))           # no edges at all
    assert np.array_equal(eng.host_matvec(np.ones(40)), np.zeros(40))


def test_trrust_layout_keeps_two_groups_resident():
    adj = trrust_adjacency()
    rng = np.random.RandomState(0)
    W = adj * rng.rand(*adj.shape)
    eng = engine.Engine(device=-1)
    eng.set_network(csr=dense_to_csr(W), n=adj.shape[0])
    lay = eng.layout()
    x = rng.randn(adj.shape[0])
    assert np.allclose(eng.host_matvec(x), W @ x, rtol=1e-12, atol=1e-10)
    assert lay["groups_per_cta"] >= 2 and lay["pingpong_genes"] <= 1024, lay


def test_bad_arguments_are_reported():
    eng = engine.Engine(device=-1)
    with pytest.raises(engine.ScrutinyError) as e:
        eng.set_network(csr=(np.array([0, 2, 1]), np.array([0, 1]), np.array([1.0, 2.0])), n=2)
    assert e.value.code == -1
    with pytest.raises(engine.ScrutinyError):
        eng.set_network(csr=(np.array([0, 1, 2]), np.array([0, 5]), np.array([1.0, 2.0])), n=2)


def test_csr_with_unsorted_duplicate_and_zero_entries():
    """scr_set_network_csr takes what scipy calls a non-canonical CSR: columns in any order, a column listed twice in a row
    (entries add up, as in ``gc.dot``), explicit zeros, empty rows, self loops."""
    eng = engine.Engine(device=-1)
    rowptr = np.array([0, 3, 5, 5, 6], dtype=np.int32)
    col = np.array([2, 0, 2, 1, 1, 3], dtype=np.int32)
    val = np.array([1.0, 2.0, 0.5, 3.0, 0.0, -4.0])
    eng.set_network(csr=(rowptr, col, val), n=4)
    x = np.array([1.0, 10.0, 100.0, 1000.0])
    assert np.array_equal(eng.host_matvec(x), [152.0, 30.0, 0.0, -4000.0])


def test_packing_fuzz_against_dense_product():
    """Random operators of every shape the packer distinguishes (sizes off the 128-gene block, empty rows, rows longer than
    the cap, hubs read by most rows, a few fully dense rows) in both precisions: the packed operator is the matrix."""
    rng = np.random.RandomState(2024)
    for case in range(40):
        n = int(rng.choice([1, 2, 7, 33, 127, 128, 129, 255, 300, 641, 1025]))
        W = np.zeros((n, n))
        density = rng.choice([0.0, 0.002, 0.02, 0.2])
        mask = rng.rand(n, n) < density
        W[mask] = rng.randn(int(mask.sum()))
        for _ in range(int(rng.randint(0, 3))):                       # hubs: one regulator read by most rows
            W[rng.rand(n) < 0.8, rng.randint(n)] = rng.randn()
        for _ in range(int(rng.randint(0, 3))):                       # long rows
            W[rng.randint(n), :] = rng.randn(n) * (rng.rand(n) < rng.choice([0.3, 1.0]))
        W[rng.rand(n) < 0.1, :] = 0.0                                  # empty rows
        eng = engine.Engine(device=-1, precision=["f32", "f64"][case % 2])
        eng.set_network(W)
        lay = eng.layout()
        assert lay["n_pad"] % 128 == 0 and lay["n_pad"] >= n
        perm = eng.gene_order()
        assert sorted(p for p in perm if p >= 0) == list(range(n))
        x = rng.randn(n)
        ref = W @ x
        assert np.allclose(eng.host_matvec(x), ref, rtol=1e-12, atol=1e-12 * max(1.0, np.abs(ref).max())), (case, n)
        eng.close()
