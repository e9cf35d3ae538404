"""Helpers shared by the oracle-consistency and GPU parity tests."""
import numpy as np

from oracle import matriseq_port as port


def noise_table(recorded, stream, cells, n):
    """Recorded (tag, vector) draws of a seeded port run -> array (cells, max_step+1, n)."""
    items = [(t, v) for t, v in recorded if t is not None and t[0] == stream]
    if not items:
        return np.zeros((cells, 1, n))
    depth = max(t[2] for t, _ in items) + 1
    out = np.zeros((cells, depth, n))
    for (s, cell, step), v in items:
        out[cell, step] = v
    return out


def tree_schedule(tree, log, iters, cells):
    """Per phase (DFS order of MS:607-612): (type, cell ids, steps per cell, step_base per cell).
    ``log`` = [(type, T)] from the port, ``iters`` = final totalCellIterations."""
    T = dict(log)
    done = np.zeros(cells, dtype=np.int64)
    phases = []

    def visit(node):
        mem = list(node.cellTypeMembers)
        if mem:
            t = int(node.identifier)
            below = list(node.childCellTypeMembers)
            ids = np.array(mem + below, dtype=np.int64)
            steps = np.concatenate([iters[mem] - done[mem], np.full(len(below), T[t], dtype=np.int64)])
            phases.append((t, ids, steps.astype(np.int64), done[ids].copy()))
            done[ids] += steps
        for ch in node.children:
            visit(tree[ch])

    visit(tree.head())
    return phases
