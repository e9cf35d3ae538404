"""Host-side producers of the simulation inputs: the gene regulatory matrix and its CSR form.

Mirrors ``generateGeneCorrelationMatrix`` of the reference (MatriSeq.py:34-172): same
arguments, same ``(gc, tf)`` result and -- because it draws from ``numpy.random`` /
``random`` in the same order -- the same matrix for the same ``init(seed=...)``.
This stays on the host: it runs once per simulation and is O(n^2) at most.
"""
import os
import random

import numpy as np

_TRRUST_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "TRRUST_edges.npz")
_NETWORK_KINDS = ("SimulatedPowerLaw", "TRRUST", "Random")


def trrust_adjacency():
    """Dense 0/1 (2895, 2895) TRRUST adjacency, rows = regulated gene, cols = regulator
    (what MatriSeq.py:99 loads from TRRUST_Network.npy); rebuilt from the packaged edge list."""
    z = np.load(_TRRUST_FILE)
    n = int(z["n"])
    adj = np.zeros((n, n))
    adj[z["rows"].astype(np.int64), z["cols"].astype(np.int64)] = 1.0
    return adj


def _power_law_degree(exponent):
    # discrete power law via inverse transform of a continuous Pareto, rounded (MS:74-78)
    return int(0.5 * (1 - np.random.uniform()) ** (-1 / (exponent - 1)) + 0.5)


def _adjacency(n, kind, exponent, sparsity):
    if kind == "TRRUST":
        return trrust_adjacency()
    adj = np.zeros((n, n))
    if kind == "SimulatedPowerLaw":
        degs = np.array([[_power_law_degree(exponent), _power_law_degree(exponent)]
                         for _ in range(n)], dtype=int).reshape(n, 2)
        degs = np.minimum(degs, n)
        stubs = np.repeat(np.arange(n), degs[:, 1]).tolist()       # multiset of regulators
        for gene in range(n):
            regulators = set(random.sample(stubs, degs[gene, 0]))   # duplicates collapse
            adj[gene, list(regulators)] = 1
    else:  # Random: Binomial(n, density) regulators per gene
        for gene in range(n):
            k = np.random.binomial(n, 1 - sparsity)
            adj[gene, random.sample(range(n), k)] = 1
    return adj


def generate_gene_correlation_matrix(n, networkStructure="TRRUST", maxAlphaBeta=10,
                                     normalizeWRows=False, powerLawExponent=2.05,
                                     networkSparsity=0.999):
    if networkStructure not in _NETWORK_KINDS:
        raise Exception('Invalid network structure. Parameter "networkStructure" must be one of '
                        '"SimulatedPowerLaw", "TRRUST", or "Random".')
    adj = _adjacency(n, networkStructure, powerLawExponent, networkSparsity)
    n = adj.shape[0]                                   # TRRUST overrides n (MS:100)
    gc = np.empty((n, n))
    for gene in range(n):                              # one Beta(a,b) law and ONE sign per row (MS:120-124)
        a = np.random.randint(1, maxAlphaBeta + 1)
        b = np.random.randint(1, maxAlphaBeta + 1)
        sign = np.random.choice([-1.0, 1.0])
        gc[gene] = sign * np.random.beta(a=a, b=b, size=n)
    gc *= adj
    if normalizeWRows:
        fan_in = np.maximum(np.count_nonzero(gc, axis=1), 1)
        gc /= fan_in[:, None]
    tf = np.nonzero(adj.sum(axis=0))[0]                # regulators = non-zero out-degree (MS:136-137)
    return gc, tf


def dense_to_csr(gc):
    """(rowptr int32[n+1], col int32[nnz], val float64[nnz]) of a dense (n, n) matrix."""
    gc = np.asarray(gc)
    rows, cols = np.nonzero(gc)
    n = gc.shape[0]
    rowptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(np.bincount(rows, minlength=n), out=rowptr[1:])
    return rowptr, cols.astype(np.int32), np.ascontiguousarray(gc[rows, cols], dtype=np.float64)
