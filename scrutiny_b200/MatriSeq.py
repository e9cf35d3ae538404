"""Drop-in for the simulation path of ``RNAscrutiny.MatriSeq`` running on B200.

Same module-level function surface, argument names, defaults, return values, in-place mutation
and output artefacts as the reference (RNAscrutiny/RNAscrutiny/MatriSeq.py, cited ``MS``); the
work of ``developCell`` / ``findCellNextState`` / ``findAllNextStates`` /
``findOptimalIterations`` / ``findOptimalAlpha`` / ``transformData`` runs in the CUDA library
behind ``include/scrutiny_b200.h``.  There is no CPU path: without the library or a GPU these
functions raise.

Usage is the reference's (example.py)::

    import scrutiny_b200.MatriSeq as ms
    ms.init(outputDir="out", seed=1337)
    gc, tf = ms.generateGeneCorrelationMatrix(n=3000, networkStructure="SimulatedPowerLaw")
    tree, projMatrix, constMatrix = ms.formCellTypeTree(...)
    states = ms.generateInitialStates(n, cells)
    alpha = ms.findOptimalAlpha(n, gc, tree, states, numToSample=50); gc /= alpha
    ms.findAllNextStates(gc, tree.head(), tree, states, cellTypesRecord, totalCellIterations)
    counts, cellSizeFactors = ms.transformData(n, cells, states)
    ms.saveData(counts, cellTypesRecord, gc, projMatrix, constMatrix, cellSizeFactors, pseudotimes)

Deliberate differences (DESIGN.md "Reference quirks"):
  * random numbers inside the GPU kernels come from a counter-based Philox generator keyed by
    ``init(seed)``; host-side draws (sampling, step counts, gamma targets, size factors) still use
    ``numpy.random`` / ``random`` like the reference.  Seed-for-seed equality with the reference's
    MT19937 noise is impossible for a parallel run; parity is checked with supplied noise.
  * keyword arguments of ``findAllNextStates`` ARE forwarded to child nodes (the reference drops
    them at MS:611-612); pass ``reference_compat=True`` to reproduce the reference's behaviour.
  * ``rangeTransformation`` exists (the packaged reference forgot it: NameError at MS:336).
  * extra trailing arguments with defaults: ``precision`` ("f32" default, "f64" parity path), ``alpha=`` on
    ``generateGeneCorrelationMatrix`` (skip the search), ``device``, ``compact=`` / ``capture=`` on ``transformData``.
  * the CUDA context holding the packed network is cached between calls (keyed on the ``gc`` buffer and its contents), so
    per-cell calls of ``developCell`` / ``findCellNextState`` do not rebuild the operator; ``init`` drops it.
"""
import os
import random

import numpy as np

from . import artefacts as _artefacts
from . import network as _network
from .engine import STREAM_ALPHA, Engine
from .lineage import Tree
from .simulate import Comm, LineageSimulation

__outputDir = None
__verbose = None
__seed = 1337
__device = 0
__precision = "f32"
__call_counter = 0


def init(outputDir='./scRutiNy/MatriSeq', seed=1337, verbose=True, device=0, precision="f32"):
    """MS:15-31: output directory, seeds of both host generators, verbosity."""
    global __outputDir, __verbose, __seed, __device, __precision, __call_counter
    __outputDir = outputDir
    if outputDir is not None:
        os.makedirs(outputDir, exist_ok=True)
    np.random.seed(seed)
    random.seed(seed)
    __verbose = verbose
    __seed = int(seed)
    __device = device
    __precision = precision
    __call_counter = 0
    _drop_engines()


def _next_seed():
    """A fresh Philox key per API call, reproducible from init(seed)."""
    global __call_counter
    __call_counter += 1
    return (__seed << 20) + __call_counter


def rangeTransformation(x):
    """Matri-seqFinal.py:406-408 (missing from the packaged module)."""
    return 2 * x - 1


# ------------------------------------------------------------------------------------------
# input producers (host)
# ------------------------------------------------------------------------------------------
def generateGeneCorrelationMatrix(n, networkStructure='TRRUST', maxAlphaBeta=10, normalizeWRows=False,
                                  powerLawExponent=2.05, networkSparsity=0.999, alpha=None, sparse=False):
    """MS:34-172 -> (gc, tf).  The network is generated as CSR (O(nnz + n) memory, same random stream as the reference);
    ``gc`` is the reference's dense (n, n) float64 array unless ``sparse=True`` (extension), which returns a
    ``scipy.sparse.csr_matrix`` every function of this module accepts in its place.  ``alpha`` (extension): divide gc by it
    right away."""
    n, rowptr, col, val, tf = _network.generate_csr(n, networkStructure, maxAlphaBeta, normalizeWRows, powerLawExponent,
                                                    networkSparsity)
    if __verbose:
        print("Number of Connections: ", len(val))
        print("Network Density: ", float(len(val)) / n ** 2)
    if alpha is not None:
        val = val / alpha
    if sparse:
        import scipy.sparse as sp
        return sp.csr_matrix((val, col, rowptr), shape=(n, n)), tf
    return _network.csr_to_dense(n, rowptr, col, val), tf


def getNextProjConst(n, projInit, constInit, constProp, transcriptionFactors):
    """MS:300-337: clamp a random ``constProp`` share of the regulators to +-1."""
    frozen = random.sample(np.ndarray.tolist(np.asarray(transcriptionFactors)),
                           int(len(transcriptionFactors) * constProp))
    proj = np.array(projInit, dtype=float, copy=True)
    proj[frozen] = 0
    const = np.array(constInit, dtype=float, copy=True)
    for gene in np.nonzero((proj == 0) & (np.asarray(projInit) != 0))[0]:     # ascending, like MS:333
        const[gene] = rangeTransformation(random.randint(0, 1))
    return proj, const


def formCellTypeTree(n, cells, cellTypeParents, cellTypeNumCells, cellTypeConstProps, transcriptionFactors,
                     cellTypeMeans, scalable=False):
    """MS:175-297 -> (tree, projMatrix, constMatrix).  Cells get their FINAL type up front.
    Default: the reference's own assignment, draw for draw (``random.sample`` from the shrinking pool, MS:223-228), which
    costs a Python set difference over all cells per type.  ``scalable=True`` (extension, what bench.py and 10^6-cell runs
    use): one ``numpy.random.permutation`` of the cells split by ``cellTypeNumCells`` -- the same distribution (a uniformly
    random assignment with the given type sizes), members kept as sorted int64 arrays instead of Python lists."""
    kinds = len(cellTypeParents)
    tree = Tree()
    projMatrix = np.zeros((kinds, n))
    constMatrix = np.zeros((kinds, n))
    if scalable:
        root = tree.add_head(-1, np.zeros(0, dtype=np.int64), np.zeros(n), np.ones(n), 0)
        perm = np.random.permutation(cells)
        bounds = np.cumsum([0] + [int(c) for c in cellTypeNumCells])
        assigned = [np.sort(perm[bounds[k]:bounds[k + 1]]) for k in range(kinds)]
        join = lambda parts: np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)   # noqa: E731
    else:
        root = tree.add_head(-1, [], np.zeros(n), np.ones(n), 0)
        pool = range(cells)
        assigned = []
        for kind in range(kinds):
            chosen = random.sample(pool, cellTypeNumCells[kind])
            assigned.append(chosen)
            pool = list(set(pool) - set(chosen))
        join = lambda parts: [c for part in parts for c in part]                                # noqa: E731
    parents = np.array(cellTypeParents)

    def attach(parent, proj_up, const_up):
        gathered = []
        for kind in np.where(parents == parent)[0]:
            proj, const = getNextProjConst(n, proj_up, const_up, cellTypeConstProps[kind], transcriptionFactors)
            projMatrix[kind], constMatrix[kind] = proj, const
            tree.add_node(int(kind), assigned[kind], const, proj, cellTypeMeans[kind], parent)
            gathered += attach(int(kind), proj, const) + [assigned[kind]]
        tree[parent].setChildCellTypeMembers(join(gathered))
        return gathered

    attach(-1, np.ones(n), np.zeros(n))
    if not scalable:
        root.setChildCellTypeMembers(list(tree[-1].childCellTypeMembers))
    return tree, projMatrix, constMatrix


def addTreeChildren(n, cellTypeTree, cellTypeParents, cellTypeMembers, cellTypeConstProps, projInit, constInit, projMatrix,
                    constMatrix, cellTypeMeans, parentCellType, transcriptionFactors):
    """MS:243-297, the reference's public recursion helper (``formCellTypeTree`` above uses its own closure): hang every
    type whose parent is ``parentCellType`` under it -- proj / const inherited through ``getNextProjConst``, rows of
    ``projMatrix`` / ``constMatrix`` filled in place -- recurse, record on the parent the cells of everything below it, and
    return those cells followed by ``cellTypeMembers[parentCellType]`` (for the head, -1, that is the LAST type's list: the
    reference's negative index, MS:297, kept)."""
    below = []
    for kind in np.flatnonzero(np.asarray(cellTypeParents) == parentCellType):
        proj, const = getNextProjConst(n, projInit, constInit, cellTypeConstProps[kind], transcriptionFactors)
        projMatrix[kind, :] = proj
        constMatrix[kind, :] = const
        cellTypeTree.add_node(kind, cellTypeMembers[kind], const, proj, cellTypeMeans[kind], parentCellType)
        below += addTreeChildren(n, cellTypeTree, cellTypeParents, cellTypeMembers, cellTypeConstProps, proj, const, projMatrix,
                                 constMatrix, cellTypeMeans, kind, transcriptionFactors)
    cellTypeTree[parentCellType].setChildCellTypeMembers(below)
    return below + list(cellTypeMembers[parentCellType])


def generateInitialStates(n, cells, sameInitialCell=True, geneMeanLevels=0):
    """MS:340-365 -> (n, cells) float64."""
    first = (2.0 * np.random.rand(n) - 1.0) + geneMeanLevels
    if sameInitialCell:
        return np.repeat(first[:, None], cells, axis=1)
    out = np.empty((n, cells))
    for j in range(cells):
        out[:, j] = (2.0 * np.random.rand(n) - 1.0) + geneMeanLevels
    return out


# ------------------------------------------------------------------------------------------
# GPU-backed hot path
# ------------------------------------------------------------------------------------------
_engine_cache = {}        # (device, precision) -> [fingerprint of gc, Engine]: one context per device and precision is kept


def _fingerprint(gc):
    """Cheap identity of a dense network argument: buffer address and shape plus two content checks that any rescaling
    or edit of the matrix changes (full sum, bytes of a strided sample) -- a few ms at n = 3000 against the ~0.5 s of
    converting, packing and uploading the operator again."""
    if hasattr(gc, "tocsr"):                                                     # scipy sparse: its three arrays
        m = gc.tocsr()
        return ("csr", m.shape, m.data.__array_interface__["data"][0], float(m.data.sum()), hash(m.indices.tobytes()),
                hash(m.indptr.tobytes()))
    gc = np.asarray(gc)
    flat = gc.reshape(-1)
    return (gc.__array_interface__["data"][0], gc.shape, gc.dtype.str, float(flat.sum()),
            hash(flat[::max(1, flat.size // 65536)].tobytes()))


def _engine_for(gc, cells, states=None, precision=None, device=None):
    """The context holding ``gc``: developCell / findCellNextState / findOptimalIterations are called once per cell or
    per node by user code (MS:503-537, 683-736), so the packed operator is kept between calls and only rebuilt when the
    matrix changes."""
    dev = __device if device is None else device
    key = (dev, precision or __precision)
    fp = _fingerprint(gc)
    slot = _engine_cache.get(key)
    if slot is None or slot[0] != fp:
        if slot is not None:
            slot[1].close()
        eng = Engine(dev, key[1])
        eng.set_network(gc)
        _engine_cache[key] = slot = [fp, eng]
    eng = slot[1]
    if eng.cells != cells:
        eng.alloc_cells(cells)
    if states is not None:
        eng.set_states(states)
    return eng


def _drop_engines():
    for slot in _engine_cache.values():
        slot[1].close()
    _engine_cache.clear()


def developCell(n, gc, timeStep, initialState, proj, const, noiseDisp, geneMeanLevels, precision=None):
    """MS:503-537: one step of one cell.  ``geneMeanLevels``: scalar or per-gene vector, like the reference's broadcast."""
    return findCellNextState(gc, n, timeStep, 0, proj, const, initialState, 1, noiseDisp, geneMeanLevels, False,
                             precision=precision)


def findCellNextState(gc, n, timeStep, i, proj, const, initialState, cellIterations, noiseDisp, geneMeanLevels,
                      showPlots, precision=None):
    """MS:683-736: ``cellIterations`` steps of one cell -> new state (n,)."""
    eng = _engine_for(gc, 1, np.asarray(initialState, dtype=np.float64)[:, None], precision)
    eng.run_phase([0], [int(cellIterations)], proj, const, dt=timeStep, sigma=noiseDisp, mu=geneMeanLevels,
                  seed=_next_seed())
    return eng.get_states()[:, 0]


def findOptimalIterations(n, gc, timeStep, cellTypeMembers, proj, const, geneMeanLevels, convergenceThreshold,
                          noiseDisp, currentStates, maxNumIterations, convDistAvgNum, numToSample, precision=None):
    """MS:616-680 -> int."""
    cols = random.sample(cellTypeMembers, numToSample)
    eng = _engine_for(gc, numToSample, np.asarray(currentStates)[:, cols], precision)
    it = eng.probe(np.arange(numToSample), proj, const, dt=timeStep, sigma=noiseDisp, mu=geneMeanLevels,
                   thresh=convergenceThreshold, max_iter=maxNumIterations, avg_num=convDistAvgNum, seed=_next_seed())
    T = int(int(it.sum()) / numToSample) - convDistAvgNum
    if __verbose:
        print("Average Iterations: ", T)
    return T


def findOptimalAlpha(n, gc, cellTypeTree, currentStates, numToSample, timeStep=0.01, convergenceThreshold=0.1,
                     noiseDisp=0.05, maxNumIterations=500, convDistAvgNum=20, initialAlpha=0.01, alphaStep=0.02,
                     showAlphaGeneMeans=False, geneMeanLevels=0, precision=None, batch=24):
    """MS:368-500 -> alpha: smallest value on the 0.02 / 0.01 ladder for which every sampled cell
    converges within ``maxNumIterations``.  The ladder is evaluated ``batch`` rungs at a time on the
    GPU; the sequence of candidates, the per-rung cell samples and the final state of ``random`` are
    exactly those of the sequential search (python-3 semantics of MS:489)."""
    currentStates = np.asarray(currentStates)
    cells = currentStates.shape[1]
    head = cellTypeTree.head()
    if len(head.children) == 1:                                                  # MS:429-432
        first = cellTypeTree[head.children[0]]
        proj, const = first.proj, first.const
    else:                                                                        # MS:433-435
        proj, const = np.zeros(n), np.zeros(n)
    alpha = initialAlpha - alphaStep if initialAlpha >= alphaStep else 0         # MS:421-424
    seed = _next_seed()
    cpg = 2 if (precision or __precision) == "f64" else 4
    per = -(-numToSample // cpg) * cpg                                           # rung padded to whole cell groups
    samples, states_after = [], []
    rung = 0
    rate = 0.0
    while True:
        while len(samples) < rung + batch:                                       # MS:442, in order
            samples.append(random.sample(range(cells), numToSample))
            states_after.append(random.getstate())
        if rate > 0:                                                             # MS:449-451 (sticky)
            alphaStep = 0.01
        step = alphaStep
        ladder, a = [], alpha
        for _ in range(batch):
            a = a + step
            ladder.append(a)
        eng = _engine_for(gc, per * batch, precision=precision)
        block = np.zeros((n, per * batch))
        for j in range(batch):
            block[:, j * per:j * per + numToSample] = currentStates[:, samples[rung + j]]
        eng.set_states(block)
        # every slot is probed (padding slots hold zeros and are ignored) so that a cell group never
        # mixes two rungs of the ladder
        it = eng.probe(np.arange(per * batch), proj, const, dt=timeStep, sigma=noiseDisp, mu=geneMeanLevels,
                       thresh=convergenceThreshold, max_iter=maxNumIterations, avg_num=convDistAvgNum, seed=seed,
                       stream=STREAM_ALPHA + rung, w_div=np.repeat(ladder, per)).reshape(batch, per)[:, :numToSample]
        rates = (it != maxNumIterations).sum(axis=1) / numToSample               # MS:480-481, 489
        done = False
        for j in range(batch):
            alpha, rate = ladder[j], rates[j]
            if __verbose:
                print("Alpha: ", alpha)
                print("Success Rate: ", rate)
                print("Average Iterations: ", int(int(it[j].sum()) / numToSample) - convDistAvgNum)
            rung += 1
            if rate >= 1:
                done = True
                break
            if rate > 0 and step != 0.01:                                        # first partial success
                break                                                            # ladder switches to 0.01 steps
        if done:
            break
    random.setstate(states_after[rung - 1])
    if __verbose and (precision or __precision) == "f32" and alpha < 0.2:
        # gain 1 / alpha >= 5: float32 rounding is amplified over a few hundred steps
        print("Note: alpha = %.2f; for trajectories that follow the float64 reference step for step over hundreds of steps "
              "use init(precision=\"f64\")" % alpha)
    return alpha


def findAllNextStates(gc, cellTypeNode, cellTypeTree, currentStates, cellTypesRecord, totalCellIterations,
                      timeStep=0.01, noiseDisp=0.05, convergenceThreshold=0.1, maxNumIterations=500,
                      convDistAvgNum=20, cellsToSample=50, showPlots=False, cellDevelopmentMode=True,
                      geneMeanLevels=0, reference_compat=False, precision=None, fixed_iterations=None):
    """MS:540-612: develop every cell along the lineage below ``cellTypeNode``.  Mutates
    ``currentStates`` (n, cells), ``cellTypesRecord`` and ``totalCellIterations`` in place; returns None."""
    cells = currentStates.shape[1]
    eng = _engine_for(gc, cells, currentStates, precision)
    sim = LineageSimulation(eng, cells, comm=Comm(single=True), seed=_next_seed(), verbose=bool(__verbose))
    user = dict(timeStep=timeStep, noiseDisp=noiseDisp, convergenceThreshold=convergenceThreshold,
                maxNumIterations=maxNumIterations, convDistAvgNum=convDistAvgNum, cellsToSample=cellsToSample,
                cellDevelopmentMode=cellDevelopmentMode, geneMeanLevels=geneMeanLevels)
    first = True
    for node in cellTypeTree.walk(cellTypeNode):
        kw = user if (first or not reference_compat) else {}                     # MS:611-612 drops them
        sim.run_node(node, cellTypesRecord, totalCellIterations, fixed_iterations=fixed_iterations, **kw)
        first = False
    if currentStates.dtype == np.float64 and currentStates.strides[1] == 8:
        eng.get_states(out=currentStates)
    else:
        currentStates[...] = eng.get_states()


def transformData(n, cells, finalCellStates, dropoutProportion=0.0, cellRadiusDisp=0.14, geneScale=0.0, expScale=1.0,
                  expLambda=0.0, shape=0.35, rate=0.19, gammaDistMean=True, precision=None, capture=None, compact=None,
                  applyDropout=False, nbDispersion=0.0):
    """MS:816-891 -> (counts (n, cells) float64 holding integers, cellSizeFactors).
    ``dropoutProportion`` is accepted and unused, exactly like the reference (MS:819).
    ``applyDropout=True`` (extension, default off): every entry is independently zeroed with probability
    ``dropoutProportion`` -- the parallel form of the reference's commented-out dropout (MS:887-889); ``nbDispersion`` phi > 0
    (extension, default off): negative-binomial counts with mean lambda and variance lambda + phi lambda^2 instead of Poisson.
    ``compact`` (extension): False -> always the reference's dense float64 array; True -> an
    ``artefacts.CountMatrix`` (unsigned bytes + overflow list on the host, 1/8 of the bytes, reads like the reference's
    array); None (default) -> dense up to ``artefacts.DENSE_LIMIT`` bytes (2 GiB), compact above.
    ``capture`` (extension, parity aid): a dict that receives the Poisson rate matrix ``lambda`` (n, cells) and the
    per-gene ``shift`` the read-out used.
    Like the reference (MS:858-862) the dense path leaves ``finalCellStates`` shifted to the new gene levels in place; the
    compact path (10^6-cell runs) leaves it untouched."""
    key = ("readout", __device, precision or __precision)
    slot = _engine_cache.get(key)
    if slot is None or slot[0] != n:                                             # the read-out needs no network: a context
        if slot is not None:                                                     # of its own, kept like the stepping one
            slot[1].close()
        eng = Engine(__device, key[2])
        eng.set_genes(n)
        _engine_cache[key] = slot = [n, eng]
    eng = slot[1]
    try:
        if eng.cells != cells:
            eng.alloc_cells(cells)
        eng.set_states(finalCellStates)
        sim = LineageSimulation(eng, cells, comm=Comm(single=True), seed=_next_seed())
        ext = dict(nb_dispersion=float(nbDispersion), dropout=float(dropoutProportion) if applyDropout else 0.0)
        dense = (compact is False or (compact is None and 8 * n * cells <= _artefacts.DENSE_LIMIT) or capture is not None
                 or ext["nb_dispersion"] > 0 or ext["dropout"] > 0)
        counts, size = sim.counts(cellRadiusDisp=cellRadiusDisp, expScale=expScale, shape=shape, rate=rate,
                                  geneScale=geneScale, expLambda=expLambda, gammaDistMean=gammaDistMean, capture=capture,
                                  dtype="int32" if dense else "uint8", **ext)
    except Exception:
        _engine_cache.pop(key, None)
        eng.close()
        raise
    if dense:
        # MS:858-862 shifts the caller's matrix in place before rebinding the name at MS:873: same side effect here
        if (isinstance(finalCellStates, np.ndarray) and finalCellStates.flags.writeable
                and np.issubdtype(finalCellStates.dtype, np.floating)):
            finalCellStates += sim.last_shift[:, None]
        return np.ascontiguousarray(counts.T, dtype=np.float64), size
    return _artefacts.CountMatrix([counts[0]], [counts[1]], n), size


# ------------------------------------------------------------------------------------------
# artefacts
# ------------------------------------------------------------------------------------------
ARTEFACTS = ("synthscRNAseq", "cellTypesRecord", "gc", "projMatrix", "constMatrix", "cellSizeFactors",
             "synthPseudotimes")


def saveData(synthscRNAseq, cellTypesRecord, gc, projMatrix, constMatrix, cellSizeFactors, synthPseudotimes):
    """MS:1035-1065: the seven ``<name>.npy`` artefacts RegNetInference.readCellStates loads
    (RegNetInference.py:49-66) and their R twins ``<name>.gzip`` (what ``save(..., compress=TRUE)``
    writes through rpy2 at MS:1062-1065), produced here without R by ``rdata.write_rdata``."""
    from .rdata import write_rdata
    for name, value in zip(ARTEFACTS, (synthscRNAseq, cellTypesRecord, gc, projMatrix, constMatrix,
                                       cellSizeFactors, synthPseudotimes)):
        if isinstance(value, _artefacts.CountMatrix):
            # compact counts: shards + manifest always; the reference's dense float64 twins only while they stay small
            for r, (blk, ovf) in enumerate(zip(value.blocks, value.overflows)):
                _artefacts.save_shard(__outputDir, np.asarray(blk), ovf, r)
            _artefacts.write_manifest(__outputDir, value.genes, [b.shape[0] for b in value.blocks], value.blocks[0].dtype)
            if 8 * value.genes * value.cells > _artefacts.DENSE_LIMIT:
                continue
            value = np.asarray(value)
        np.save(os.path.join(__outputDir, name), value)
        write_rdata(os.path.join(__outputDir, name + ".gzip"), name, np.asarray(value))


def analyzeDataBeforeTransformation(*args, **kwargs):
    """MS:739-813 only draws plots; out of scope for the simulation path (no-op)."""


def analyzeSCRNAseqData(*args, **kwargs):
    """MS:894-1032 only draws plots / PCA diagnostics; out of scope (no-op)."""


def plotSCRNAseqData(*args, **kwargs):
    """MS:970-1032 only draws PCA / t-SNE plots of the count matrix; out of scope (no-op, kept so that a script written
    against the reference runs unchanged)."""
