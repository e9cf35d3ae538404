"""TEST INFRASTRUCTURE ONLY -- summary statistics of the artefacts the reference SHIPS as the output of its example.py
(RNAscrutiny/RNAscrutiny/example/scRutiNy/MatriSeq/*.npy: SURVEY.md section 4 lists them as the material that pins shapes,
dtypes, orientation and value ranges) -> tests/golden/example_fix_summary.json.  Build container only:

    python -m oracle.make_fix_summary
"""
import json
import os

import numpy as np

from oracle import reference_harness as rh

FIX = os.path.join(rh.REFERENCE_ROOT, "RNAscrutiny", "RNAscrutiny", "example", "scRutiNy", "MatriSeq")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "example_fix_summary.json")
NAMES = ("synthscRNAseq", "cellTypesRecord", "gc", "projMatrix", "constMatrix", "cellSizeFactors", "synthPseudotimes")


def main():
    a = {k: np.load(os.path.join(FIX, k + ".npy")) for k in NAMES}
    counts, gc = a["synthscRNAseq"], a["gc"]
    out = {k: {"shape": list(v.shape), "dtype": v.dtype.str} for k, v in a.items()}
    out["stats"] = {
        "zero_fraction": float((counts == 0).mean()),
        "mean_library_size": float(counts.sum(axis=0).mean()),
        "max_count": float(counts.max()),
        "counts_integer_valued": bool(np.array_equal(counts, np.round(counts))),
        "median_gene_mean": float(np.median(counts.mean(axis=1))),
        "pseudotime_max": float(a["synthPseudotimes"].max()),
        "pseudotime_min": float(a["synthPseudotimes"].min()),
        "size_factor_cbrt_mean": float(np.cbrt(a["cellSizeFactors"]).mean()),
        "size_factor_cbrt_sd": float(np.cbrt(a["cellSizeFactors"]).std()),
        "gc_abs_max": float(np.abs(gc).max()),
        "gc_nnz": int(np.count_nonzero(gc)),
        "cells_per_type": np.bincount(a["cellTypesRecord"]).tolist(),
        "proj_all_ones": bool((a["projMatrix"] == 1).all()),
        "const_all_zero": bool((a["constMatrix"] == 0).all()),
    }
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(out["stats"], indent=1))


if __name__ == "__main__":
    main()
