"""Drop-in for the data-parallel front end of ``RNAscrutiny.RegNetInference`` on B200 (SURVEY.md section 8 row
f4): reading the MatriSeq artefacts, the double rank transform and the cell-pair search.  Same function names,
arguments and results as the reference (RNAscrutiny/RNAscrutiny/RegNetInference.py, cited ``RNI``); the work runs
in the CUDA library behind ``include/scrutiny_b200.h`` (``scr_rank_transform``, ``scr_cell_pairs``).  The inverse
problem itself (clustering, Lasso fits, plots: RNI:112-396, 417-754) is out of scope.

    import scrutiny_b200.RegNetInference as rni
    rni.init(inputDir="out/MatriSeq", outputDir=None)
    n, cells, S, pseudotimes, cellTypesRecord, W_real = rni.readCellStates()
    pairs = rni.findCellPairs(S, pseudotimes, cellTypesRecord, pseudotimeThresh=0.1)
"""
import os
import random

import numpy as np

from .engine import Engine

__inputDir = None
__outputDir = None
__verbose = None
__device = 0
__engine = None


def init(inputDir='./scRutiNy/MatriSeq', outputDir='./scRutiNy/RegNetInference', seed=1337, verbose=True, device=0):
    """RNI:26-46."""
    global __inputDir, __outputDir, __verbose, __device, __engine
    __inputDir = inputDir
    __outputDir = outputDir
    if outputDir is not None:
        os.makedirs(outputDir, exist_ok=True)
    np.random.seed(seed)
    random.seed(seed)
    __verbose = verbose
    if __engine is not None and device != __device:
        __engine.close()
        __engine = None
    __device = int(device)


def _eng():
    global __engine
    if __engine is None:
        __engine = Engine(__device, "f32")
    return __engine


def readCellStates():
    """RNI:49-66: loads the artefacts MatriSeq.saveData wrote and rank-transforms the count matrix.  When only the
    compact shards exist (counts above ``artefacts.DENSE_LIMIT``), the matrix is assembled from them."""
    from . import artefacts

    def artefact(name):
        return np.load(os.path.join(__inputDir, name + ".npy"))

    dense = os.path.join(__inputDir, "synthscRNAseq.npy")
    cellStates = np.load(dense) if os.path.isfile(dense) else np.asarray(artefacts.load(__inputDir))
    pseudotimes, cellTypesRecord, W_real = artefact("synthPseudotimes"), artefact("cellTypesRecord"), artefact("gc")
    n, cells = cellStates.shape
    cellStates = convertToS(n, cells, cellStates)
    if __verbose:
        for label, value in (("Num genes: %d , num cells: " % n, cells), ("Cell States: ", cellStates),
                             ("Num TF Connections: ", int(np.count_nonzero(W_real))), ("Pseudotimes: ", pseudotimes)):
            print(label, value)
    return n, cells, cellStates, pseudotimes, cellTypesRecord, W_real


def convertToS(n, cells, cellStates):
    """RNI:69-78.  In place when ``cellStates`` is a float64 array with contiguous rows (what the reference
    mutates); returns the transformed array either way."""
    if cellStates.shape != (n, cells):
        raise ValueError("cellStates must have shape (n, cells)")
    S = cellStates
    if not (isinstance(S, np.ndarray) and S.dtype == np.float64 and S.strides[1] == 8 and S.strides[0] % 8 == 0
            and S.strides[0] >= 8 * cells):
        S = np.ascontiguousarray(cellStates, dtype=np.float64)
    _eng().rank_transform(S)
    if S is not cellStates:
        cellStates[...] = S
    return cellStates


def findCellPairs(cellStates, pseudotimes, cellTypesRecord, pseudotimeThresh):
    """The pair search the reference inlines at RNI:397-416 and RNI:764-777: list of ``[cell1, cell2]`` of the
    same type with ``0 < pseudotime1 - pseudotime2 <= pseudotimeThresh`` and differing states."""
    S = cellStates
    if not (S.dtype == np.float64 and S.strides[1] == 8 and S.strides[0] % 8 == 0):
        S = np.ascontiguousarray(S, dtype=np.float64)
    return _eng().cell_pairs(S, pseudotimes, cellTypesRecord, pseudotimeThresh).tolist()
