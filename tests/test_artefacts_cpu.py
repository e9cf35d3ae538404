"""Host-side pieces added in round 2 that need no GPU: the compact / sharded count artefact and the
Poisson(T) member-step draw of the prepared mode (MS:590-591)."""
import json
import os

import numpy as np
import pytest

from scrutiny_b200 import artefacts
from scrutiny_b200.simulate import poisson_quantile


def compact_of(dense_cells_by_genes, dtype):
    esc = artefacts.ESCAPE[np.dtype(dtype)]
    rows, genes = np.nonzero(dense_cells_by_genes >= esc)
    ovf = np.stack([rows, genes, dense_cells_by_genes[rows, genes]], axis=1).astype(np.int64)
    return np.minimum(dense_cells_by_genes, esc).astype(dtype), ovf


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
def test_count_matrix_reads_like_the_reference_array(dtype):
    rng = np.random.RandomState(1)
    genes, cells = 37, 101
    full = rng.poisson(2.0, size=(cells, genes)).astype(np.int64)
    full[rng.randint(cells, size=9), rng.randint(genes, size=9)] = [255, 256, 300, 65535, 65536, 70000, 254, 1000, 99999]
    edges = [0, 40, 41, 101]
    blocks, ovfs = zip(*[compact_of(full[a:b], dtype) for a, b in zip(edges[:-1], edges[1:])])
    m = artefacts.CountMatrix(blocks, ovfs)
    ref = full.T.astype(np.float64)                          # what MatriSeq.transformData returns: (genes, cells) float64
    assert m.shape == ref.shape and m.dtype == np.float64 and len(m) == genes
    assert np.array_equal(np.asarray(m), ref)
    assert np.array_equal(m[:, 40], ref[:, 40]) and np.array_equal(m[5, :], ref[5, :])
    assert np.array_equal(m[3:9, 38:45], ref[3:9, 38:45]) and np.array_equal(m[:, [0, 100, 41]], ref[:, [0, 100, 41]])
    assert m[7, 99] == ref[7, 99] and m.sum() == ref.sum()
    assert np.array_equal(m.rows(39, 43, np.int64), full[39:43])
    assert m.nbytes_compact() < ref.nbytes // (3 if dtype == np.uint16 else 7)   # tiny matrix: the overflow rows weigh in


def test_shards_manifest_and_save_data(tmp_path):
    import scrutiny_b200.MatriSeq as ms
    rng = np.random.RandomState(2)
    genes, cells = 20, 64
    full = rng.poisson(3.0, size=(cells, genes)).astype(np.int64)
    full[10, 3] = 4000
    parts = [compact_of(full[:30], np.uint8), compact_of(full[30:], np.uint8)]
    for r, (blk, ovf) in enumerate(parts):
        artefacts.save_shard(str(tmp_path), blk, ovf, r)
    man = artefacts.write_manifest(str(tmp_path), genes, [30, 34], np.uint8)
    assert man["cells"] == cells and json.load(open(tmp_path / "synthscRNAseq.manifest.json"))["dtype"] == "uint8"
    back = artefacts.load(str(tmp_path))
    assert np.array_equal(np.asarray(back), full.T.astype(np.float64))
    # saveData with a CountMatrix writes shards + manifest and (small matrix) the reference's dense artefact too
    out = tmp_path / "run"
    ms.init(outputDir=str(out), seed=3, verbose=False)
    m = artefacts.CountMatrix([p[0] for p in parts], [p[1] for p in parts])
    ms.saveData(m, np.zeros(cells, dtype="int64"), np.eye(genes), np.ones((2, genes)), np.zeros((2, genes)), np.ones(cells),
                np.linspace(0, 1, cells))
    assert np.array_equal(np.load(out / "synthscRNAseq.npy"), full.T.astype(np.float64))
    assert np.array_equal(np.asarray(artefacts.load(str(out))), full.T.astype(np.float64))
    assert os.path.isfile(out / "synthscRNAseq.gzip") and os.path.isfile(out / "synthPseudotimes.npy")
    # above the dense limit only the shards are written
    old = artefacts.DENSE_LIMIT
    artefacts.DENSE_LIMIT = 100
    try:
        big = tmp_path / "big"
        ms.init(outputDir=str(big), seed=3, verbose=False)
        ms.saveData(m, np.zeros(cells, dtype="int64"), np.eye(genes), np.ones((2, genes)), np.zeros((2, genes)), np.ones(cells),
                    np.linspace(0, 1, cells))
        assert not os.path.exists(big / "synthscRNAseq.npy") and artefacts.has_manifest(str(big))
        assert os.path.isfile(big / "cellTypesRecord.npy")
    finally:
        artefacts.DENSE_LIMIT = old


def test_poisson_quantile_is_the_inverse_cdf():
    from scipy import stats
    u = np.concatenate([np.linspace(0, 1, 5001)[:-1], [1 - 1e-12]])
    for lam in (0, 0.3, 7, 80, 480):
        k = poisson_quantile(u, lam)
        if lam == 0:
            assert not k.any()
            continue
        want = stats.poisson.ppf(u, lam)
        want[u == 0] = 0
        assert np.abs(k - want).max() <= 1 and (k != want).mean() < 2e-3   # ties at cdf jumps may round either way
    rng = np.random.RandomState(4)
    k = poisson_quantile(rng.rand(400000), 412)
    assert abs(k.mean() - 412) < 0.2 and abs(k.var() / 412 - 1) < 0.02
