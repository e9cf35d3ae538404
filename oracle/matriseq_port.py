"""TEST INFRASTRUCTURE ONLY (oracle) -- faithful CPU restatement of the MatriSeq hot path.

This file is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it.

It restates, per cell and in float64 exactly like the reference does, the functions
on the simulation path of ``RNAscrutiny/RNAscrutiny/MatriSeq.py`` (cited as ``MS``),
consuming ``numpy.random`` (legacy MT19937 ``RandomState``) and Python's ``random``
in the *same call order*, so that a seeded run reproduces the reference's output
bit for bit.  That is how the oracle is pinned: ``tests/golden/*.npz`` hold outputs
of the unmodified reference run in the build container (``oracle/make_golden.py``)
and ``tests/test_oracle_golden.py`` replays them through this port.

Parity status: the reference ships no tests and does not pin numpy;
the pin is "outputs of the reference itself run here" (python 3.12.3, numpy 2.3.5).

All random draws go through a ``Draws`` object so tests can record or replace the
normal noise (supplied-noise parity mode) without touching the arithmetic.
"""
import random as _pyrandom

import numpy as np


class Draws:
    """RNG adapter.  Default: the same global generators the reference uses
    (``np.random.*`` seeded by MS:27 and ``random.*`` seeded by MS:28)."""

    def normal_vec(self, scale, n, tag=None):      # MS:531-532
        return np.random.normal(loc=0, scale=scale, size=n)

    def uniform_scalar(self):                       # MS:707 (stray draw), MS:74-78
        return np.random.uniform()

    def randint(self, lo, hi):                      # MS:589, MS:121-122
        return np.random.randint(lo, hi)

    def poisson(self, lam):                         # MS:591, MS:884
        return np.random.poisson(lam)

    def sample(self, population, k):                # MS:92,107,226,326,442,644
        return _pyrandom.sample(population, k)

    def py_randint(self, lo, hi):                   # MS:336
        return _pyrandom.randint(lo, hi)


class RecordingDraws(Draws):
    """Same streams, but keeps every normal vector in draw order."""

    def __init__(self):
        self.normals = []

    def normal_vec(self, scale, n, tag=None):
        v = super().normal_vec(scale, n, tag)
        self.normals.append((tag, v.copy()))
        return v


class SuppliedNoise(Draws):
    """Normal noise looked up by ``tag = (stream, cell, step)`` instead of drawn.
    ``table[stream]`` is an array ``(cells, steps, n)`` of *standard-deviation-scaled*
    noise (already multiplied by noiseDisp).  Everything else still uses the global
    generators."""

    def __init__(self, table):
        self.table = table

    def normal_vec(self, scale, n, tag=None):
        stream, cell, step = tag
        return np.asarray(self.table[stream][cell, step, :n], dtype=np.float64)


_DEFAULT = Draws()


# ----------------------------------------------------------------------------
# lineage container -- restates tree.py:4-54 (only what the hot path touches)
# ----------------------------------------------------------------------------
class PortNode:
    def __init__(self, identifier, members, const, proj, mean):
        self.identifier = identifier
        self.children = []
        self.cellTypeMembers = members
        self.childCellTypeMembers = []
        self.const = const
        self.proj = proj
        self.cellTypeMean = mean


class PortTree:
    def __init__(self):
        self.nodes = {}
        self._head = None

    def head(self):
        return self._head

    def add_head(self, identifier, members, const, proj, mean):
        node = PortNode(identifier, members, const, proj, mean)
        self.nodes[identifier] = node
        self._head = node
        return node

    def add_node(self, identifier, members, const, proj, mean, parent):
        node = PortNode(identifier, members, const, proj, mean)
        self.nodes[identifier] = node
        self.nodes[parent].children.append(identifier)
        return node

    def __getitem__(self, key):
        return self.nodes[key]


# ----------------------------------------------------------------------------
# input producers (host side in the product too)
# ----------------------------------------------------------------------------
def generate_network(n, structure="TRRUST", maxAlphaBeta=10, normalizeWRows=False,
                     powerLawExponent=2.05, networkSparsity=0.999, trrust_adjacency=None,
                     draws=_DEFAULT):
    """MS:34-172.  Returns (gc, tf).  ``trrust_adjacency`` replaces the np.load at MS:99."""
    adj = np.zeros((n, n))
    if structure == "SimulatedPowerLaw":                      # MS:69-95
        indeg = np.empty(n, dtype=int)
        outdeg = np.empty(n, dtype=int)
        expo = -1 / (powerLawExponent - 1)
        for g in range(n):
            indeg[g] = int(0.5 * (1 - draws.uniform_scalar()) ** expo + 0.5)
            outdeg[g] = int(0.5 * (1 - draws.uniform_scalar()) ** expo + 0.5)
        indeg[indeg > n] = n
        outdeg[outdeg > n] = n
        pool = []
        for g in range(n):
            pool += [g] * outdeg[g]
        for row in range(n):
            picks = draws.sample(pool, indeg[row])
            adj[row, list(set(picks))] = 1
    elif structure == "TRRUST":                                # MS:97-101
        adj = np.array(trrust_adjacency, dtype=np.float64)
        n = adj.shape[0]
    elif structure == "Random":                                # MS:103-108
        for row in range(n):
            k = np.random.binomial(n, (1 - networkSparsity))
            adj[row, draws.sample(range(n), k)] = 1
    else:                                                      # MS:110-113
        raise Exception('Invalid network structure. Parameter "networkStructure" must be one of '
                        '"SimulatedPowerLaw", "TRRUST", or "Random".')

    values = np.zeros((n, n))
    for row in range(n):                                       # MS:120-124 (per ROW despite the name)
        a = draws.randint(1, maxAlphaBeta + 1)
        b = draws.randint(1, maxAlphaBeta + 1)
        values[row, :] = np.random.choice([-1.0, 1.0]) * np.random.beta(a=a, b=b, size=n)
    gc = adj * values                                          # MS:127
    if normalizeWRows:                                         # MS:130-133
        cnt = np.count_nonzero(gc, axis=1)
        cnt[cnt == 0] = 1
        gc /= cnt[:, np.newaxis]
    tf = np.nonzero(np.sum(adj, axis=0))[0]                    # MS:136-137
    return gc, tf


def next_proj_const(n, proj0, const0, constProp, tf, draws=_DEFAULT):
    """MS:300-337 with rangeTransformation = 2x-1 (Matri-seqFinal.py:406-408)."""
    k = int(len(tf) * constProp)
    zeroed = draws.sample(np.ndarray.tolist(tf), k)
    proj = np.copy(proj0)
    proj[zeroed] = 0
    const = np.copy(const0)
    for g in range(n):
        if proj[g] == 0 and proj0[g] != 0:
            const[g] = 2 * draws.py_randint(0, 1) - 1
    return proj, const


def form_tree(n, cells, parents, numCells, constProps, tf, means, draws=_DEFAULT):
    """MS:175-297.  Returns (tree, projMatrix, constMatrix)."""
    tree = PortTree()
    ntypes = len(parents)
    projM = np.zeros((ntypes, n))
    constM = np.zeros((ntypes, n))
    head = tree.add_head(-1, [], np.zeros(n), np.ones(n), 0)   # MS:216-217
    head.childCellTypeMembers = list(range(cells))             # MS:218
    members = [[] for _ in range(ntypes)]
    pool = range(cells)
    for t in range(ntypes):                                    # MS:223-228
        idx = draws.sample(pool, numCells[t])
        members[t] = idx
        pool = list(set(pool) - set(idx))

    def grow(parent, proj0, const0):                           # MS:243-297
        kids = np.where(np.array(parents) == parent)[0]
        below = []
        for t in kids:
            proj, const = next_proj_const(n, proj0, const0, constProps[t], tf, draws)
            projM[t, :] = proj
            constM[t, :] = const
            tree.add_node(t, members[t], const, proj, means[t], parent)
            below += grow(t, proj, const)
        tree[parent].childCellTypeMembers = below
        return below + members[parent]   # parent == -1 indexes the last type, like MS:297; value unused

    grow(-1, np.ones(n), np.zeros(n))
    return tree, projM, constM


def initial_states(n, cells, sameInitialCell=True, geneMeanLevels=0):
    """MS:340-365."""
    out = np.zeros((n, cells))
    base = (2.0 * np.random.rand(n) - 1.0) + geneMeanLevels
    for j in range(cells):
        if sameInitialCell:
            out[:, j] = base
        else:
            out[:, j] = (2.0 * np.random.rand(n) - 1.0) + geneMeanLevels
    return out


# ----------------------------------------------------------------------------
# the hot path
# ----------------------------------------------------------------------------
def develop_cell(n, gc, dt, x, proj, const, noiseDisp, mu, draws=_DEFAULT, tag=None):
    """One Euler/tanh step of one cell, MS:503-537 (tau == 1, MS:529)."""
    target = np.tanh(gc.dot(x - mu) + draws.normal_vec(noiseDisp, n, tag))   # MS:530-532
    moved = x + dt * (target - x)                                             # MS:533-534 (tau == 1)
    return np.multiply(proj, moved) + const + mu                               # MS:535


def cell_next_state(gc, n, dt, proj, const, x, iters, noiseDisp, mu, draws=_DEFAULT,
                    cell=None, step0=0, stream="step"):
    """``iters`` sequential steps of one cell, MS:683-736 (diagnostics dropped, the
    stray uniform draw at MS:707 kept because it advances the global stream)."""
    draws.uniform_scalar()                                                     # MS:707
    for s in range(iters):                                                     # MS:709-717
        x = develop_cell(n, gc, dt, x, proj, const, noiseDisp, mu, draws,
                         tag=(stream, cell, step0 + s))
    return x


def probe_one_cell(n, gc, dt, x, proj, const, noiseDisp, mu, thresh, maxIter, avgNum,
                   draws=_DEFAULT, cell=None, stream="probe", trace=None):
    """Inner while-loop shared by MS:657-669 and MS:464-479: step until the mean over the
    last ``avgNum`` values of mean(|dx|)/dt drops to ``thresh`` or ``maxIter`` is hit.
    Returns (iterations, final state)."""
    diffs = [0] * maxIter
    it = 0
    dist = float("inf")
    while it < maxIter and dist > thresh:
        new = develop_cell(n, gc, dt, x, proj, const, noiseDisp, mu, draws, tag=(stream, cell, it))
        diffs[it] = np.mean(np.absolute(new - x) / dt)
        if it >= avgNum - 1:
            dist = np.mean(diffs[(it - avgNum + 1):(it + 1)])
        if trace is not None:
            trace.append((cell, it, diffs[it], dist))
        x = new
        it += 1
    return it, x


def optimal_iterations(n, gc, dt, members, proj, const, mu, thresh, noiseDisp, states,
                       maxIter, avgNum, numToSample, draws=_DEFAULT, trace=None, sampled=None):
    """MS:616-680.  ``sampled`` (optional list) receives the chosen columns."""
    cols = draws.sample(members, numToSample)                                  # MS:644
    if sampled is not None:
        sampled.extend(cols)
    total = 0
    for i in range(numToSample):                                               # MS:650
        it, _ = probe_one_cell(n, gc, dt, states[:, cols[i]].copy(), proj, const, noiseDisp, mu,
                               thresh, maxIter, avgNum, draws, cell=cols[i], trace=trace)
        total += it
    return int(total / numToSample) - avgNum                                   # MS:672-673


def all_next_states(gc, node, tree, states, typesRecord, totalIters, timeStep=0.01,
                    noiseDisp=0.05, convergenceThreshold=0.1, maxNumIterations=500,
                    convDistAvgNum=20, cellsToSample=50, cellDevelopmentMode=True,
                    geneMeanLevels=0, forward_kwargs=False, draws=_DEFAULT, log=None):
    """MS:540-612.  ``forward_kwargs=False`` reproduces reference quirk R1: the recursive
    call at MS:611-612 passes no keyword arguments, so every non-head node runs with the
    defaults of MS:546-554.  ``log`` (optional list) receives (type, T) per stepped node."""
    n = gc.shape[0]
    members = node.cellTypeMembers
    typesRecord[members] = int(node.identifier)                                # MS:576
    if len(members) != 0:                                                      # MS:579-580
        T = optimal_iterations(n, gc, timeStep, members, node.proj, node.const, geneMeanLevels,
                               convergenceThreshold, noiseDisp, states, maxNumIterations,
                               convDistAvgNum, cellsToSample, draws)
        if log is not None:
            log.append((int(node.identifier), int(T)))
        for m in members:                                                      # MS:587-597
            k = draws.randint(0, T + 1) if cellDevelopmentMode else draws.poisson(T)
            states[:, m] = cell_next_state(gc, n, timeStep, node.proj, node.const, states[:, m], k,
                                           noiseDisp, geneMeanLevels, draws, cell=m,
                                           step0=int(totalIters[m]))
            totalIters[m] += k
        for d in node.childCellTypeMembers:                                    # MS:598-605
            states[:, d] = cell_next_state(gc, n, timeStep, node.proj, node.const, states[:, d], T,
                                           noiseDisp, geneMeanLevels, draws, cell=d,
                                           step0=int(totalIters[d]))
            totalIters[d] += T
    for child in node.children:                                                # MS:607-612
        if forward_kwargs:
            all_next_states(gc, tree[child], tree, states, typesRecord, totalIters, timeStep,
                            noiseDisp, convergenceThreshold, maxNumIterations, convDistAvgNum,
                            cellsToSample, cellDevelopmentMode, geneMeanLevels, True, draws, log)
        else:
            all_next_states(gc, tree[child], tree, states, typesRecord, totalIters,
                            forward_kwargs=False, draws=draws, log=log)


def optimal_alpha(n, gc, tree, states, numToSample, timeStep=0.01, convergenceThreshold=0.1,
                  noiseDisp=0.05, maxNumIterations=500, convDistAvgNum=20, initialAlpha=0.01,
                  alphaStep=0.02, geneMeanLevels=0, draws=_DEFAULT, log=None):
    """MS:368-500 under python-3 semantics (true division at MS:489)."""
    cells = states.shape[1]
    rate = 0
    alpha = initialAlpha - alphaStep if initialAlpha >= alphaStep else 0      # MS:421-424
    head = tree.head()
    if len(head.children) == 1:                                                # MS:429-432
        proj, const = tree[head.children[0]].proj, tree[head.children[0]].const
    else:                                                                      # MS:433-435 (quirk R2)
        proj, const = np.zeros(n), np.zeros(n)
    while rate < 1:                                                            # MS:438
        cols = draws.sample(range(cells), numToSample)                         # MS:442
        if rate > 0:                                                           # MS:449-451
            alphaStep = 0.01
        alpha = alpha + alphaStep
        wins = 0
        scaled = gc / alpha                                                    # MS:461
        for i in range(numToSample):
            it, _ = probe_one_cell(n, scaled, timeStep, states[:, cols[i]].copy(), proj, const,
                                   noiseDisp, geneMeanLevels, convergenceThreshold,
                                   maxNumIterations, convDistAvgNum, draws, cell=cols[i],
                                   stream="alpha")
            if it != maxNumIterations:                                         # MS:480-481
                wins += 1
        rate = wins / numToSample                                              # MS:489
        if log is not None:
            log.append((alpha, rate))
    return alpha


def transform_data(n, cells, S, cellRadiusDisp=0.14, geneScale=0.0, expScale=1.0, expLambda=0.0,
                   shape=0.35, rate=0.19, gammaDistMean=True, draws=_DEFAULT, capture=None):
    """MS:816-891.  Mutates ``S`` up to the exp like the reference; returns (counts, sizeFactors).
    ``capture`` (optional dict) receives the intermediate lambda matrix and gene shifts."""
    means = np.mean(S, axis=1)                                                 # MS:851
    order = np.argsort(means)                                                  # MS:852
    if gammaDistMean:                                                          # MS:854-862
        target = np.sort(np.log(np.random.gamma(shape, scale=1 / rate, size=n)) / expScale)
        for r in range(len(order)):
            g = order[r]
            S[g, :] += target[r] - np.mean(S[g, :])
    else:                                                                      # MS:863-871
        target = np.sort((geneScale + expLambda) / expScale + -1 * np.random.exponential(expLambda, size=n))
        for r in range(len(order)):
            S[order[r], :] += target[r]
    lam = np.exp(expScale * S)                                                 # MS:873
    size = np.random.normal(loc=1, scale=cellRadiusDisp, size=cells) ** 3      # MS:875-876
    for j in range(cells):                                                     # MS:877-878
        lam[:, j] *= size[j]
    if capture is not None:
        capture["lambda"] = lam.copy()
        capture["target_sorted"] = target
        capture["order"] = order
    for g in range(n):                                                         # MS:880-885 gene-major draw order
        for j in range(cells):
            if lam[g, j] < 0:
                lam[g, j] = 0
            lam[g, j] = draws.poisson(lam[g, j])
    return lam, size
