"""Compact and sharded form of the ``synthscRNAseq`` artefact.

The reference returns and saves the count matrix as float64 ``(genes, cells)`` (MatriSeq.py:891, 1050-1061): 24 GB per
10^6 cells at n = 3000.  The GPU read-out produces it as unsigned bytes ``(cells, genes)`` plus a short overflow list
(``scr_counts_compact``); this module keeps it that way on the host and on disk:

  * ``CountMatrix`` -- the compact rows (one block per shard) with the reference's view on demand: ``shape == (genes,
    cells)``, ``np.asarray(m)`` / ``m[:, j]`` / ``m[g0:g1, :]`` give float64 values exactly as the reference would hold
    them, materialised only for the part asked for;
  * ``save_shard`` / ``write_manifest`` / ``load`` -- one ``synthscRNAseq.shard<r>.npy`` (+ ``.overflow.npy``) per rank and
    a small JSON manifest; ``load`` memory-maps the shards.  ``MatriSeq.saveData`` writes the reference's own
    ``synthscRNAseq.npy`` as well whenever the float64 matrix is small enough to be worth having (<= ``DENSE_LIMIT``).
"""
import json
import os

import numpy as np

NAME = "synthscRNAseq"
DENSE_LIMIT = 2 << 30          # bytes of float64 (genes, cells) up to which the reference's dense artefact is also written
ESCAPE = {np.dtype(np.uint8): 0xFF, np.dtype(np.uint16): 0xFFFF}


class CountMatrix:
    """Counts as compact ``(cells, genes)`` row blocks + overflow triplets ``(row, gene, count)`` per block; behaves like
    the reference's float64 ``(genes, cells)`` array for reading."""

    def __init__(self, blocks, overflows=None, genes=None):
        self.blocks = list(blocks)
        self.overflows = [np.zeros((0, 3), dtype=np.int64) if o is None else np.asarray(o, dtype=np.int64).reshape(-1, 3)
                          for o in (overflows or [None] * len(self.blocks))]
        self.genes = int(genes if genes is not None else self.blocks[0].shape[1])
        self.starts = np.cumsum([0] + [b.shape[0] for b in self.blocks])
        self.cells = int(self.starts[-1])
        self.shape = (self.genes, self.cells)
        self.dtype = np.dtype(np.float64)
        self.ndim = 2

    # -- compact access ---------------------------------------------------------------------------------------
    def rows(self, c0, c1, dtype=np.int32):
        """Cells [c0, c1) as a ``(c1 - c0, genes)`` array of ``dtype`` with the overflow entries patched in."""
        out = np.empty((c1 - c0, self.genes), dtype=dtype)
        for b, (blk, ovf, s0) in enumerate(zip(self.blocks, self.overflows, self.starts[:-1])):
            lo, hi = max(c0, s0), min(c1, s0 + blk.shape[0])
            if lo >= hi:
                continue
            out[lo - c0:hi - c0] = blk[lo - s0:hi - s0]
            if len(ovf):
                sel = ovf[(ovf[:, 0] >= lo - s0) & (ovf[:, 0] < hi - s0)]
                out[sel[:, 0] + s0 - c0, sel[:, 1]] = sel[:, 2]
        return out

    def nbytes_compact(self):
        return int(sum(b.nbytes for b in self.blocks) + sum(o.nbytes for o in self.overflows))

    # -- the reference's view ---------------------------------------------------------------------------------
    def __array__(self, dtype=None, copy=None):
        out = np.empty(self.shape, dtype=dtype or np.float64)
        step = max(1, (64 << 20) // max(self.genes, 1))
        for c0 in range(0, self.cells, step):
            c1 = min(self.cells, c0 + step)
            out[:, c0:c1] = self.rows(c0, c1).T
        return out

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key, slice(None))
        gkey, ckey = key
        if isinstance(ckey, (int, np.integer)):
            c = int(ckey) % self.cells
            return self.rows(c, c + 1)[0].astype(np.float64)[gkey]
        if isinstance(ckey, slice):
            c0, c1, st = ckey.indices(self.cells)
            if st == 1:
                return self.rows(c0, max(c0, c1)).T.astype(np.float64)[gkey]
        cols = np.arange(self.cells)[ckey]
        return np.stack([self.rows(int(c), int(c) + 1)[0] for c in cols], axis=1).astype(np.float64)[gkey]

    def __len__(self):
        return self.genes

    def sum(self):
        return float(sum(self.rows(int(s), int(e), np.int64).sum() for s, e in zip(self.starts[:-1], self.starts[1:])))


def shard_paths(directory, rank):
    base = os.path.join(directory, "%s.shard%05d" % (NAME, rank))
    return base + ".npy", base + ".overflow.npy"


def save_shard(directory, counts, overflow, rank):
    """One rank's rows: ``counts`` (cells_local, genes) compact array, ``overflow`` (k, 3) int64."""
    path, opath = shard_paths(directory, rank)
    np.save(path, counts)
    np.save(opath, np.asarray(overflow, dtype=np.int64).reshape(-1, 3))
    return path


def write_manifest(directory, genes, shard_cells, dtype):
    """``shard_cells``: cells per shard in rank order (global cell ids are contiguous per shard)."""
    man = {"artefact": NAME, "layout": "cells x genes row-major per shard; escape code = dtype max, true value in .overflow.npy "
                                       "rows (local row, gene, count)",
           "genes": int(genes), "cells": int(sum(shard_cells)), "dtype": np.dtype(dtype).name,
           "shards": [{"rank": r, "cells": int(c), "file": os.path.basename(shard_paths(directory, r)[0])}
                      for r, c in enumerate(shard_cells)]}
    with open(os.path.join(directory, NAME + ".manifest.json"), "w") as f:
        json.dump(man, f, indent=1)
    return man


def has_manifest(directory):
    return os.path.isfile(os.path.join(directory, NAME + ".manifest.json"))


def load(directory, mmap=True):
    """CountMatrix over the shards a manifest lists (memory-mapped)."""
    with open(os.path.join(directory, NAME + ".manifest.json")) as f:
        man = json.load(f)
    blocks, ovfs = [], []
    for sh in man["shards"]:
        path, opath = shard_paths(directory, sh["rank"])
        blk = np.load(path, mmap_mode="r" if mmap else None)
        if blk.shape != (sh["cells"], man["genes"]):
            raise ValueError("shard %s has shape %s, manifest says %s" % (path, blk.shape, (sh["cells"], man["genes"])))
        blocks.append(blk)
        ovfs.append(np.load(opath))
    return CountMatrix(blocks, ovfs, man["genes"])
