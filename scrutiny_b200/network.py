"""Host-side producers of the simulation inputs: the gene regulatory matrix, generated as CSR.

Mirrors ``generateGeneCorrelationMatrix`` of the reference (MatriSeq.py:34-172): same arguments, same ``(gc, tf)``
result and -- because it draws from ``numpy.random`` / ``random`` in the same order -- the same matrix for the same
``init(seed=...)``.  The reference builds two dense (n, n) arrays (MS:58, 115-127); here the adjacency is kept as one
sorted column list per gene and the weights are filtered row by row as they are drawn, so memory is O(nnz + n)
(``generate_csr``); the dense matrix of the reference's signature is only formed at the API boundary
(``generate_gene_correlation_matrix``).  This stays on the host: it runs once per simulation.
"""
import os
import random

import numpy as np

_TRRUST_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "TRRUST_edges.npz")
_NETWORK_KINDS = ("SimulatedPowerLaw", "TRRUST", "Random")


def _trrust_rows():
    """Per gene the sorted regulator list of the TRRUST network (what MatriSeq.py:99 loads from TRRUST_Network.npy as a
    dense 0/1 matrix, rows = regulated gene); rebuilt from the packaged edge list."""
    z = np.load(_TRRUST_FILE)
    n = int(z["n"])
    rows, cols = z["rows"].astype(np.int64), z["cols"].astype(np.int64)
    order = np.lexsort((cols, rows))
    rows, cols = rows[order], cols[order]
    keep = np.ones(len(rows), dtype=bool)
    keep[1:] = (rows[1:] != rows[:-1]) | (cols[1:] != cols[:-1])
    rows, cols = rows[keep], cols[keep]
    bounds = np.searchsorted(rows, np.arange(n + 1))
    return n, [cols[bounds[g]:bounds[g + 1]] for g in range(n)]


def trrust_adjacency():
    """Dense 0/1 (2895, 2895) TRRUST adjacency (tests and tools)."""
    n, rows = _trrust_rows()
    adj = np.zeros((n, n))
    for g, cols in enumerate(rows):
        adj[g, cols] = 1.0
    return adj


def _power_law_degree(exponent):
    # discrete power law via inverse transform of a continuous Pareto, rounded (MS:74-78)
    return int(0.5 * (1 - np.random.uniform()) ** (-1 / (exponent - 1)) + 0.5)


def _adjacency_rows(n, kind, exponent, sparsity):
    """(n, [sorted regulator ids of gene 0, gene 1, ...]) drawing exactly what MS:66-108 draws."""
    if kind == "TRRUST":
        return _trrust_rows()
    rows = []
    if kind == "SimulatedPowerLaw":
        degs = np.array([[_power_law_degree(exponent), _power_law_degree(exponent)]
                         for _ in range(n)], dtype=int).reshape(n, 2)
        degs = np.minimum(degs, n)
        stubs = np.repeat(np.arange(n), degs[:, 1]).tolist()       # multiset of regulators
        for gene in range(n):
            regulators = set(random.sample(stubs, degs[gene, 0]))   # duplicates collapse
            rows.append(np.array(sorted(regulators), dtype=np.int64))
    else:  # Random: Binomial(n, density) regulators per gene
        for gene in range(n):
            k = np.random.binomial(n, 1 - sparsity)
            rows.append(np.array(sorted(random.sample(range(n), k)), dtype=np.int64))
    return n, rows


def generate_csr(n, networkStructure="TRRUST", maxAlphaBeta=10, normalizeWRows=False, powerLawExponent=2.05,
                 networkSparsity=0.999):
    """-> (n, rowptr int32[n+1], col int32[nnz], val float64[nnz], tf).  Rows = regulated gene, columns ascending."""
    if networkStructure not in _NETWORK_KINDS:
        raise Exception('Invalid network structure. Parameter "networkStructure" must be one of '
                        '"SimulatedPowerLaw", "TRRUST", or "Random".')
    n, rows = _adjacency_rows(n, networkStructure, powerLawExponent, networkSparsity)   # TRRUST overrides n (MS:100)
    tf = np.unique(np.concatenate(rows)) if n else np.zeros(0, dtype=np.int64)   # regulators = non-zero out-degree (MS:136-137)
    vals = []
    for gene in range(n):                              # one Beta(a,b) law and ONE sign per row (MS:120-124)
        a = np.random.randint(1, maxAlphaBeta + 1)
        b = np.random.randint(1, maxAlphaBeta + 1)
        sign = np.random.choice([-1.0, 1.0])
        row = sign * np.random.beta(a=a, b=b, size=n)  # the reference draws the whole row, then masks it (MS:124-127)
        v = row[rows[gene]]
        keep = v != 0.0                                # a weight that is exactly 0 is no connection (count_nonzero, MS:129)
        if not keep.all():
            rows[gene], v = rows[gene][keep], v[keep]
        if normalizeWRows and len(v):
            v = v / len(v)
        vals.append(v)
    counts = np.array([len(r) for r in rows], dtype=np.int64)
    rowptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(counts, out=rowptr[1:])
    col = (np.concatenate(rows) if n else np.zeros(0)).astype(np.int32)
    val = np.ascontiguousarray(np.concatenate(vals) if n else np.zeros(0), dtype=np.float64)
    return n, rowptr, col, val, tf


def csr_to_dense(n, rowptr, col, val):
    gc = np.zeros((n, n))
    gc[np.repeat(np.arange(n), np.diff(rowptr)), col] = val
    return gc


def generate_gene_correlation_matrix(n, networkStructure="TRRUST", maxAlphaBeta=10,
                                     normalizeWRows=False, powerLawExponent=2.05,
                                     networkSparsity=0.999):
    """The reference's signature and result: dense ``gc`` (n, n) float64 and the regulator list ``tf``."""
    n, rowptr, col, val, tf = generate_csr(n, networkStructure, maxAlphaBeta, normalizeWRows, powerLawExponent,
                                           networkSparsity)
    return csr_to_dense(n, rowptr, col, val), tf


def dense_to_csr(gc):
    """(rowptr int32[n+1], col int32[nnz], val float64[nnz]) of a dense (n, n) matrix or a scipy sparse matrix."""
    if hasattr(gc, "tocsr"):
        m = gc.tocsr()
        m.sort_indices()
        m.eliminate_zeros()
        return m.indptr.astype(np.int32), m.indices.astype(np.int32), np.ascontiguousarray(m.data, dtype=np.float64)
    gc = np.asarray(gc)
    rows, cols = np.nonzero(gc)
    n = gc.shape[0]
    rowptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(np.bincount(rows, minlength=n), out=rowptr[1:])
    return rowptr, cols.astype(np.int32), np.ascontiguousarray(gc[rows, cols], dtype=np.float64)
