"""TEST/DATA INFRASTRUCTURE -- extract the TRRUST regulatory adjacency the reference loads at
MatriSeq.py:97-101 (``TRRUST_Network.npy``; missing from the reference tree but shipped inside
``RNAscrutiny/dist/RNAscrutiny-1.1.tar.gz``) and store it as a compact edge list
``scrutiny_b200/data/TRRUST_edges.npz`` (8444 edges of a 2895-gene 0/1 matrix, ~20 KB) so that
``networkStructure='TRRUST'`` works on machines without the reference.

    python -m oracle.extract_trrust
"""
import os

import numpy as np

from oracle import reference_harness as rh


def main():
    adj = rh.load_trrust_adjacency()
    assert set(np.unique(adj)) == {0.0, 1.0}
    rows, cols = np.nonzero(adj)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                       "scrutiny_b200", "data", "TRRUST_edges.npz")
    np.savez_compressed(out, n=np.int32(adj.shape[0]), rows=rows.astype(np.int16),
                        cols=cols.astype(np.int16))
    print("wrote", out, adj.shape, len(rows))


if __name__ == "__main__":
    main()
