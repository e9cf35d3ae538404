"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the data-parallel front end of the reference's
RegNetInference module.  Never imported by scrutiny_b200/.

  convert_to_s     RegNetInference.py:69-78   per-cell then per-gene rank transform to (-1, 1]
  find_cell_pairs  RegNetInference.py:397-416 (and again :764-777 inside getCVCost)
  cv_cost          RegNetInference.py:778-797 (used only to pin the pair set against the reference)

The arithmetic the reference delegates to a dependency is ``scipy.stats.rankdata`` (method='average';
scipy is unpinned in RNAscrutiny/setup.py:14-16, 1.18.1 here).  Its published definition -- rank of a value =
mean of the 1-based positions its tie group occupies in the sorted order -- is restated in ``average_ranks``.
Pinned: tests/test_oracle_regnet.py replays tests/golden/regnet.npz (outputs of the unmodified reference made by
oracle/make_golden.py) bit for bit.
"""
import math

import numpy as np


def average_ranks(x):
    """scipy.stats.rankdata(x) with the default 'average' tie rule, as float64."""
    x = np.asarray(x)
    order = np.argsort(x, kind="stable")
    xs = x[order]
    n = len(xs)
    ranks = np.empty(n, dtype=np.float64)
    start = 0
    while start < n:
        end = start + 1
        while end < n and xs[end] == xs[start]:
            end += 1
        ranks[order[start:end]] = 0.5 * ((start + 1) + end)     # mean of start+1 .. end
        start = end
    return ranks


def average_ranks_fast(x):
    """Vectorised equivalent of ``average_ranks`` (used for the larger oracle cases)."""
    x = np.asarray(x)
    order = np.argsort(x, kind="stable")
    xs = x[order]
    n = len(xs)
    new = np.empty(n, dtype=bool)
    new[0] = True
    new[1:] = xs[1:] != xs[:-1]
    starts = np.flatnonzero(new)
    ends = np.append(starts[1:], n)
    grp = np.cumsum(new) - 1
    r = 0.5 * ((starts + 1) + ends)
    ranks = np.empty(n, dtype=np.float64)
    ranks[order] = r[grp]
    return ranks


def convert_to_s(n, cells, cell_states, ranks=average_ranks_fast):
    """RegNetInference.py:69-78, in place like the reference (returns the same array)."""
    for c in range(cells):
        r = ranks(cell_states[:, c])
        cell_states[:, c] = (2 * ((r + 1) / (n)) - 1)
    for g in range(n):
        r = ranks(cell_states[g, :])
        cell_states[g, :] = (2 * ((r + 1) / (cells)) - 1)
    return cell_states


def find_cell_pairs(cell_states, pseudotimes, cell_types, thresh):
    """RegNetInference.py:403-414: [cell1, cell2] with the same type, 0 < pt1 - pt2 <= thresh and states that
    differ in at least one gene; ordered by type, then cell1, then cell2 (all ascending)."""
    pairs = []
    for cell_type in range(int(max(cell_types)) + 1):
        of_type = np.flatnonzero(cell_types == cell_type)
        for c1 in of_type:
            p1 = pseudotimes[c1]
            for c2 in of_type:
                p2 = pseudotimes[c2]
                if math.fabs(p1 - p2) <= thresh and p1 > p2:
                    if np.any(cell_states[:, c1] - cell_states[:, c2] != 0):
                        pairs.append([int(c1), int(c2)])
    return pairs


def find_cell_pairs_fast(cell_states, pseudotimes, cell_types, thresh):
    """Vectorised equivalent (one cell1 at a time) for the larger oracle cases."""
    out = []
    for cell_type in range(int(np.max(cell_types)) + 1):
        of_type = np.flatnonzero(cell_types == cell_type)
        pt = pseudotimes[of_type]
        for c1, p1 in zip(of_type, pt):
            ok = (np.abs(p1 - pt) <= thresh) & (p1 > pt)
            for c2 in of_type[ok]:
                if np.any(cell_states[:, c1] - cell_states[:, c2] != 0):
                    out.append([int(c1), int(c2)])
    return out


def cv_cost(thresh, W, cell_states, pseudotimes, cell_types, pairs=None):
    """RegNetInference.py:778-797 on the pairs above."""
    if pairs is None:
        pairs = find_cell_pairs(cell_states, pseudotimes, cell_types, thresh)
    mse = 0
    for c1, c2 in pairs:
        s1, s2 = cell_states[:, c1], cell_states[:, c2]
        derivative = (s1 - s2) / (pseudotimes[c1] - pseudotimes[c2])
        actual = derivative * thresh + s2
        predicted = np.tanh(np.dot(W, s2))
        mse += np.mean((actual - predicted) ** 2)
    return mse / len(pairs)
