#!/usr/bin/env python
"""The reference's walk-through (RNAscrutiny/RNAscrutiny/example/example.py:6-86) on scrutiny_b200: only the import line
changes.  Needs a B200 and the built library (python -c "import __graft_entry__ as g; g.build()").

    python examples/example.py [--genes 500] [--cells 200] [--out scRutiNy] [--precision f64]

Two root cell types make the alpha search return 0.02 (gain 50), the reference's own default regime: float64 is the
precision that follows the reference there, so it is this script's default; pass --precision f32
for throughput.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import scrutiny_b200.MatriSeq as ms                       # was: import RNAscrutiny.MatriSeq as ms
import scrutiny_b200.RegNetInference as rni               # was: import RNAscrutiny.RegNetInference as rni


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genes", type=int, default=500)
    ap.add_argument("--cells", type=int, default=200)
    ap.add_argument("--out", default="scRutiNy")
    ap.add_argument("--precision", default="f64", choices=["f32", "f64"])
    ap.add_argument("--network", default="SimulatedPowerLaw", choices=["SimulatedPowerLaw", "TRRUST", "Random"])
    args = ap.parse_args()
    n, cells = args.genes, args.cells
    sim_dir = os.path.join(args.out, "MatriSeq")

    t0 = time.perf_counter()
    ms.init(outputDir=sim_dir, verbose=True, precision=args.precision)
    gc, tf = ms.generateGeneCorrelationMatrix(n=n, networkStructure=args.network)
    n = gc.shape[0]                                       # TRRUST fixes its own gene count
    half = cells // 2
    tree, projMatrix, constMatrix = ms.formCellTypeTree(n=n, cells=cells, cellTypeParents=[-1, -1],
                                                        cellTypeNumCells=[half, cells - half], cellTypeConstProps=[0.0, 0.0],
                                                        transcriptionFactors=tf, cellTypeMeans=[0, 0])
    states = ms.generateInitialStates(n, cells)
    alpha = ms.findOptimalAlpha(n, gc, tree, states, numToSample=50)
    gc /= alpha

    cellTypesRecord = np.zeros(shape=cells, dtype="int64")
    totalCellIterations = np.zeros(shape=cells, dtype="int64")
    ms.findAllNextStates(gc, tree.head(), tree, states, cellTypesRecord, totalCellIterations)    # in place
    pseudotimes = 0.01 * totalCellIterations
    ms.analyzeDataBeforeTransformation(n, cells, states, alpha, pseudotimes, networkStructure=args.network)   # plots: no-op here
    counts, cellSizeFactors = ms.transformData(n, cells, states)
    ms.analyzeSCRNAseqData(n, cells, cellTypesRecord, counts)                                     # plots: no-op here
    ms.saveData(counts, cellTypesRecord, gc, projMatrix, constMatrix, cellSizeFactors, pseudotimes)
    t1 = time.perf_counter()
    print("alpha %.2f, %d cell-steps, mean count %.3f, zero fraction %.3f, %.2f s" %
          (alpha, int(totalCellIterations.sum()), counts.mean(), float((counts == 0).mean()), t1 - t0))
    print("artefacts:", sorted(os.listdir(sim_dir)))

    # the data-parallel front end of the inverse problem (RegNetInference.py:49-78, 397-416)
    rni.init(inputDir=sim_dir, outputDir=os.path.join(args.out, "RegNetInference"), verbose=False)
    n, cells, cellStates, pseudotimes, cellTypesRecord, W_real = rni.readCellStates()
    pseudotimes = pseudotimes / np.max(pseudotimes)
    pairs = rni.findCellPairs(cellStates, pseudotimes, cellTypesRecord, 0.1)
    print("rank-transformed states in [%.3f, %.3f], %d cell pairs within 0.1 pseudotime, %d regulatory links" %
          (cellStates.min(), cellStates.max(), len(pairs), int(np.count_nonzero(W_real))))


if __name__ == "__main__":
    main()
