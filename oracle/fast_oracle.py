"""TEST INFRASTRUCTURE ONLY (oracle) -- vectorised float64 restatement, batched over cells.

Same arithmetic as ``oracle/matriseq_port.py`` (itself pinned to the reference), but the
normal noise is *looked up* by (cell, step-of-that-cell, gene) instead of drawn from a
sequential stream, which is the contract of the CUDA supplied-noise mode and what makes
runs independent of batching.  ``tests/test_oracle_consistency.py`` checks it against the
port with noise recorded from a seeded port run.
"""
import numpy as np
import scipy.sparse as sp


def as_operator(gc):
    """dense ndarray or scipy sparse -> something with ``@`` (CSR when sparse enough)."""
    if sp.issparse(gc):
        return gc.tocsr()
    gc = np.asarray(gc, dtype=np.float64)
    if np.count_nonzero(gc) < 0.05 * gc.size:
        return sp.csr_matrix(gc)
    return gc


def step_batch(W, X, proj, const, dt, noise, mu=0.0):
    """One step (MS:529-535) for a batch.  X, noise: (cells, n).  Returns new (cells, n)."""
    Z = (W @ (X - mu).T).T + noise
    T = np.tanh(Z)
    X1 = X + dt * (T - X)
    return proj * X1 + const + mu


def run_phase(W, X, steps, proj, const, dt, noise, step_base=None, mu=0.0, rows=None):
    """Step row ``i`` of X ``steps[i]`` times (MS:587-605 for one node).
    ``noise[rows[i], step_base[i] + s, :]`` is the noise of local step ``s``.  In place."""
    m = X.shape[0]
    steps = np.asarray(steps)
    rows = np.arange(m) if rows is None else np.asarray(rows)
    step_base = np.zeros(m, dtype=np.int64) if step_base is None else np.asarray(step_base)
    for s in range(int(steps.max()) if m else 0):
        act = np.nonzero(steps > s)[0]
        X[act] = step_batch(W, X[act], proj, const, dt, noise[rows[act], step_base[act] + s], mu)
    return X


def probe(W, X, proj, const, dt, noise, thresh, max_iter, avg_num, mu=0.0, scale=None):
    """Convergence probe of MS:650-669 / MS:456-479 for a batch of sampled cells.
    X: (k, n) copies; noise: (k, max_iter, n).  ``scale`` (k,) optionally divides W per row
    (the alpha ladder of MS:461).  Returns (iters[k], normdiff[k, max_iter], final X)."""
    k, n = X.shape
    X = X.copy()
    iters = np.zeros(k, dtype=np.int64)
    nd = np.zeros((k, max_iter))
    live = np.ones(k, dtype=bool)
    for it in range(max_iter):
        act = np.nonzero(live)[0]
        if act.size == 0:
            break
        Xa = X[act]
        drive = (W @ (Xa - mu).T).T
        if scale is not None:
            drive = drive / scale[act, None]
        new = proj * (Xa + dt * (np.tanh(drive + noise[act, it]) - Xa)) + const + mu
        nd[act, it] = np.mean(np.abs(new - Xa) / dt, axis=1)
        X[act] = new
        iters[act] = it + 1
        if it >= avg_num - 1:
            dist = nd[act, it - avg_num + 1:it + 1].mean(axis=1)
            live[act[dist <= thresh]] = False
    return iters, nd, X


def iterations_from_probe(iters, avg_num):
    """MS:672-673."""
    return int(int(np.sum(iters)) / len(iters)) - avg_num


def gene_shift(S, target_sorted):
    """MS:851-862: per-gene additive shift that replaces each gene mean (over cells) by the
    rank-matched sorted target.  S: (n, cells).  Returns shift (n,)."""
    means = np.mean(S, axis=1)
    order = np.argsort(means)
    shift = np.empty(S.shape[0])
    shift[order] = target_sorted - means[order]
    return shift


def poisson_lambda(S, shift, exp_scale, size_factors):
    """MS:873-883: lambda = max(exp(expScale*(S+shift)) * sizeFactor, 0).  S: (n, cells)."""
    lam = np.exp(exp_scale * (S + shift[:, None])) * size_factors[None, :]
    return np.maximum(lam, 0.0)
