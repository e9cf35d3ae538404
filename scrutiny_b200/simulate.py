"""Host scheduler of the lineage simulation: turns the cell-type tree into per-node GPU phases.

Restates the control flow of ``findAllNextStates`` / ``findOptimalIterations`` /
``transformData`` (reference MatriSeq.py:540-680, 816-891) around the CUDA engine:

  for every node in depth-first order (MS:607-612):
      record the type of its member cells                               (MS:576)
      T  <- probe: sample cells, step copies to convergence             (MS:583-586, 616-680)
      members step k ~ U{0..T} (or Poisson(T)), descendants step T      (MS:587-605)

Cells are sharded over ranks by contiguous global id; the network, proj/const and every host
random draw are replicated (all ranks hold the same ``numpy.random`` / ``random`` state), the
probe's iteration sum is the only value exchanged per node, the per-gene sums the only value
exchanged in the read-out.  Noise is keyed by global cell id, so the result does not depend on
the number of ranks.
"""
import os
import random
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .engine import STREAM_PROBE


class Comm:
    """Sum all-reduce over ranks.  Single process: identity.  With torch.distributed initialised
    (NCCL on GPU boxes, gloo in CPU tests) it reduces through a tensor on ``device``."""

    def __init__(self, device=None, single=False):
        self.rank, self.world = 0, 1
        self._dist = None
        self._device = device
        if single:          # a one-process run inside a multi-rank job (the MatriSeq module's functions): no collectives
            return
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self._dist = dist
                self.rank, self.world = dist.get_rank(), dist.get_world_size()
        except ImportError:
            pass

    def allreduce_sum(self, arr):
        arr = np.asarray(arr, dtype=np.float64)
        if self._dist is None or self.world == 1:
            return arr
        import torch
        t = torch.from_numpy(arr.copy())
        if self._device is not None:
            t = t.to(self._device)
        self._dist.all_reduce(t)
        return t.cpu().numpy()


def shard_bounds(cells, rank, world):
    """Contiguous [lo, hi) of global cell ids owned by ``rank`` (multiples of 4 except the tail)."""
    per = -(-cells // world)
    per = -(-per // 4) * 4
    lo = min(cells, rank * per)
    return lo, min(cells, lo + per)


class LineageSimulation:
    """One rank's share of a MatriSeq run.  ``engine`` must hold the network and ``hi - lo`` cells
    allocated with ``global_offset = lo``."""

    def __init__(self, engine, cells, lo=0, hi=None, comm=None, seed=1337, verbose=False):
        self.eng = engine
        self.cells = int(cells)
        self.lo, self.hi = int(lo), int(cells if hi is None else hi)
        self.comm = comm or Comm()
        self.seed = int(seed)
        self.verbose = verbose
        self.node_iterations = {}
        self.phase_log = []
        self.plan = None

    # -- probe ---------------------------------------------------------------------------------
    def optimal_iterations(self, members, proj, const, timeStep=0.01, noiseDisp=0.05, convergenceThreshold=0.1,
                           maxNumIterations=500, convDistAvgNum=20, numToSample=50, geneMeanLevels=0.0, tag=0):
        """findOptimalIterations (MS:616-680) on the GPU.  ``random.sample`` consumes the same
        stream position as the reference (MS:644)."""
        cols = random.sample(list(members) if not isinstance(members, list) else members, numToSample)
        cols = np.asarray(cols, dtype=np.int64)
        mine = cols[(cols >= self.lo) & (cols < self.hi)]
        total = 0.0
        if len(mine):
            it = self.eng.probe(mine - self.lo, proj, const, dt=timeStep, sigma=noiseDisp, mu=geneMeanLevels,
                                thresh=convergenceThreshold, max_iter=maxNumIterations, avg_num=convDistAvgNum,
                                seed=self.seed + 7919 * (tag + 1), stream=STREAM_PROBE)
            total = float(it.sum())
        total = float(self.comm.allreduce_sum(np.array([total]))[0])
        return int(total / numToSample) - convDistAvgNum            # MS:672-673

    # -- one node -------------------------------------------------------------------------------
    def run_node(self, node, cellTypesRecord, totalCellIterations, timeStep=0.01, noiseDisp=0.05,
                 convergenceThreshold=0.1, maxNumIterations=500, convDistAvgNum=20, cellsToSample=50,
                 cellDevelopmentMode=True, geneMeanLevels=0.0, fixed_iterations=None):
        members = np.asarray(node.cellTypeMembers, dtype=np.int64)
        if len(members):
            cellTypesRecord[members] = int(node.identifier)                      # MS:576
        if len(members) == 0:                                                    # MS:579-580
            return 0
        if fixed_iterations is None:
            T = self.optimal_iterations(node.cellTypeMembers, node.proj, node.const, timeStep, noiseDisp,
                                        convergenceThreshold, maxNumIterations, convDistAvgNum, cellsToSample,
                                        geneMeanLevels, tag=int(node.identifier))
        else:
            T = int(fixed_iterations)
        self.node_iterations[int(node.identifier)] = T
        if self.verbose:
            print("Average Iterations: ", T)
        if cellDevelopmentMode:                                                  # MS:588-591
            k = np.random.randint(0, T + 1, size=len(members))
        else:
            k = np.random.poisson(T, size=len(members))
        below = np.asarray(node.childCellTypeMembers, dtype=np.int64)
        ids = np.concatenate([members, below])
        steps = np.concatenate([k.astype(np.int64), np.full(len(below), T, dtype=np.int64)])
        own = (ids >= self.lo) & (ids < self.hi)
        loc_ids, loc_steps = ids[own], steps[own]
        if len(loc_ids) and loc_steps.max() > 0:
            self.eng.run_phase(loc_ids - self.lo, loc_steps, node.proj, node.const,
                               step_base=totalCellIterations[loc_ids], dt=timeStep, sigma=noiseDisp,
                               mu=geneMeanLevels, seed=self.seed)
        self.phase_log.append((int(node.identifier), T, int(loc_steps.sum()) if len(loc_ids) else 0))
        totalCellIterations[ids] += steps                                        # MS:597, 605 (replicated)
        return T

    # -- whole tree ---------------------------------------------------------------------------
    def run_tree(self, tree, start, cellTypesRecord, totalCellIterations, fixed_iterations=None, **kw):
        for node in tree.walk(start):
            self.run_node(node, cellTypesRecord, totalCellIterations, fixed_iterations=fixed_iterations, **kw)

    # -- read-out -----------------------------------------------------------------------------
    def gene_shift(self, shape=0.35, rate=0.19, expScale=1.0, geneScale=0.0, expLambda=0.0, gammaDistMean=True):
        """Per-gene additive shift of MS:851-871 (global means via one all-reduce of n doubles)."""
        sums = self.comm.allreduce_sum(self.eng.gene_sums())
        means = sums / self.cells
        order = np.argsort(means)
        n = len(means)
        shift = np.empty(n)
        if gammaDistMean:
            target = np.sort(np.log(np.random.gamma(shape, scale=1 / rate, size=n)) / expScale)
            shift[order] = target - means[order]
        else:
            target = np.sort((geneScale + expLambda) / expScale + -1 * np.random.exponential(expLambda, size=n))
            shift[order] = target
        return shift

    def counts(self, cellRadiusDisp=0.14, expScale=1.0, out=None, out_device_ptr=None, capture=None, dtype="int32",
               nb_dispersion=0.0, dropout=0.0, **kw):
        """transformData (MS:816-891) for this rank's cells.  Returns (counts (cells_local, n) int32,
        all cellSizeFactors); with ``dtype`` "uint8" / "uint16" the first item is ``(compact counts, overflow rows)``.
        ``capture`` (dict, parity aid) receives ``lambda`` (n, cells_local) and ``shift``."""
        shift = self.last_shift = self.gene_shift(expScale=expScale, **kw)
        size = np.random.normal(loc=1, scale=cellRadiusDisp, size=self.cells) ** 3   # MS:875-876
        if nb_dispersion or dropout:                                             # default-off extensions (scr_counts_ext)
            got = self.eng.counts_ext(shift, size[self.lo:self.hi], nb_dispersion=nb_dispersion, dropout=dropout,
                                      exp_scale=expScale, seed=self.seed, want_lambda=capture is not None)
            if capture is not None:
                got, lam = got
                capture["lambda"] = np.ascontiguousarray(lam.T)
                capture["shift"] = shift
            return got, size
        if dtype != "int32":
            return self.eng.counts_compact(shift, size[self.lo:self.hi], exp_scale=expScale, seed=self.seed, dtype=dtype,
                                           out=out, out_device_ptr=out_device_ptr), size
        got = self.eng.counts(shift, size[self.lo:self.hi], exp_scale=expScale, seed=self.seed,
                              out=out, out_device_ptr=out_device_ptr, want_lambda=capture is not None)
        if capture is not None:
            got, lam = got
            capture["lambda"] = np.ascontiguousarray(lam.T)
            capture["shift"] = shift
        return got, size

    # =========================================================================================
    # prepared (scalable) mode: all per-cell bookkeeping is rank-local, every per-cell random
    # draw is a pure function of (seed, node, GLOBAL cell id) so that the result does not depend
    # on the number of ranks.  Distributionally identical to run_node (k ~ U{0..T}, MS:589;
    # size factor ~ N(1, sd)^3, MS:875-876); used by bench.py and for >= 10^6 cells.
    # =========================================================================================
    def prepare(self, tree):
        """Per node: global member ids (for the probe sample) and the LOCAL slots of this rank's
        members / descendant cells.  Done once per tree."""
        self.plan = []
        for node in tree.walk():
            mem = np.asarray(node.cellTypeMembers, dtype=np.int64)
            below = np.asarray(node.childCellTypeMembers, dtype=np.int64)
            mem_glob = mem[(mem >= self.lo) & (mem < self.hi)]
            below_loc = below[(below >= self.lo) & (below < self.hi)] - self.lo
            self.plan.append(dict(node=node, members=mem, mem_glob=mem_glob, mem_loc=mem_glob - self.lo, below_loc=below_loc,
                                  ids=np.concatenate([mem_glob - self.lo, below_loc]), parent=None))
        by_id = {int(p["node"].identifier): p for p in self.plan}
        for p in self.plan:
            for child in p["node"].children:
                if int(child) in by_id:
                    by_id[int(child)]["parent"] = int(p["node"].identifier)
        return self

    def _member_uniforms(self, p):
        """U[0,1) per local member of a node: a pure function of (seed, node, GLOBAL cell id)."""
        return hashed_uniform(self.seed, 2 * int(p["node"].identifier) + 11, p["mem_glob"])

    def _size_factors(self, cellRadiusDisp):
        """N(1, sd)^3 per local cell (MS:875-876), Box-Muller on hashed uniforms of the global cell id."""
        gid = np.arange(self.lo, self.hi, dtype=np.int64)
        u1 = 1.0 - hashed_uniform(self.seed, 5, gid)
        u2 = hashed_uniform(self.seed, 6, gid)
        z = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
        return (1.0 + cellRadiusDisp * z) ** 3

    def run_all(self, tree=None, counts_out=None, counts_dev_ptr=None, seed=None, timeStep=0.01, noiseDisp=0.05,
                convergenceThreshold=0.1, maxNumIterations=500, convDistAvgNum=20, cellsToSample=50,
                geneMeanLevels=0.0, fixed_iterations=None, cellRadiusDisp=0.14, expScale=1.0, do_counts=True,
                counts_blocks=1, on_counts_block=None, counts_dtype="int32", cellDevelopmentMode=True):
        """findAllNextStates + transformData for this rank's shard (tree prepared beforehand).
        Returns dict(global_cell_steps, local_cell_steps, h2d_bytes, iterations, types, size_factors).

        ``counts_blocks`` > 1 samples the count matrix in that many row blocks and calls
        ``on_counts_block(row0, row1)`` after each: the caller can pass finished rows on (NCCL gather in bench.py)
        while the next block is sampled.  ``counts_dtype`` "uint8" / "uint16" stores the counts compactly
        (scr_counts_compact); the result then carries ``overflow`` = (k, 3) int64 rows (local row, gene, count) for the
        few entries that do not fit.  ``cellDevelopmentMode=False`` draws the member step counts from Poisson(T)
        instead of U{0..T} (MS:590-591), by inversion of the same hashed uniforms.

        The per-cell draws that do not depend on a node's iteration count T (member uniforms, size factors) are
        produced by one helper thread while the main thread sits in the engine's calls (ctypes and numpy release
        the GIL), so they stay inside the run but off its critical path."""
        if self.plan is None:
            self.prepare(tree)
        if seed is not None:
            self.seed = int(seed)
            np.random.seed(self.seed % (2 ** 32))                                # replicated host draws
            random.seed(self.seed)
        nloc = self.hi - self.lo
        iters = np.zeros(nloc, dtype=np.int64)
        types = np.zeros(nloc, dtype=np.int64)
        local_steps, h2d = 0, 0
        self.node_iterations = {}
        pool = ThreadPoolExecutor(max_workers=1)
        try:
            # helper-thread work, in submission order: the type record (no dependence on T), then per node the member
            # uniforms; the iteration bookkeeping of a node is queued right before its phase is launched, so it runs
            # under that phase and the next node finds its step bases ready
            live = [p for p in self.plan if len(p["members"])]                   # MS:579-580 skips empty types

            def record_types():
                for p in live:
                    types[p["mem_loc"]] = int(p["node"].identifier)

            def bookkeeping(ids, steps, next_ids):
                iters[ids] += steps                                              # ids are unique within a node
                return iters[next_ids] if next_ids is not None else None

            typed = pool.submit(record_types)
            draws = [pool.submit(self._member_uniforms, p) for p in live]
            size_f = pool.submit(self._size_factors, cellRadiusDisp) if do_counts else None
            base_f = None
            # Probe samples of every node, drawn in the order the reference draws them (MS:644 inside the depth-first walk).
            # The probes themselves are batched: a node's sampled cells are final as soon as its PARENT's phase is, so when
            # the walk reaches the first child of a node the probes of all its children go out together (scr_probe_multi:
            # side by side on the GPU).  Same samples, same noise keys, same T as probing each node right before its phase.
            picks, T_of, stepped = None, {}, set()
            if fixed_iterations is None:
                picks = [p["members"][random.sample(range(len(p["members"])), cellsToSample)] for p in live]

            live_ids = {int(p["node"].identifier) for p in live}

            # Order of the phases: the reference walks the tree depth first (MS:607-612); subtrees touch disjoint cells and all
            # per-cell randomness is keyed by (seed, node, cell, iterations done), so any order that steps a node after its
            # parent gives the same states.  Level order (default; SCR_LEVEL_ORDER=0 keeps the depth-first walk) makes ALL
            # nodes of a tree level ready together: one round of side-by-side probes per level instead of one per parent.
            by_id = {int(p["node"].identifier): j for j, p in enumerate(live)}
            depth = []
            for p in live:
                d, up = 0, p["parent"]
                while up in by_id:
                    d, up = d + 1, live[by_id[up]]["parent"]
                depth.append(d)
            level_order = os.environ.get("SCR_LEVEL_ORDER", "1") != "0"
            sched = sorted(range(len(live)), key=lambda j: (depth[j], j)) if level_order else list(range(len(live)))

            def probe_ready(first):
                # ready = not probed yet and nothing above it is still to be stepped (an empty parent type never steps: MS:579-580)
                batch = [j for j in sched if j not in T_of
                         and (live[j]["parent"] not in live_ids or live[j]["parent"] in stepped)]
                mine = [picks[j][(picks[j] >= self.lo) & (picks[j] < self.hi)] - self.lo for j in batch]
                totals = np.zeros(len(batch))
                have = [b for b, m in enumerate(mine) if len(m)]
                if have:
                    kw = dict(dt=timeStep, sigma=noiseDisp, mu=geneMeanLevels, thresh=convergenceThreshold,
                              max_iter=maxNumIterations, avg_num=convDistAvgNum, stream=STREAM_PROBE)
                    nodes = [live[batch[b]]["node"] for b in have]
                    seeds = [self.seed + 7919 * (int(nd.identifier) + 1) for nd in nodes]
                    if len(have) == 1 or getattr(self.eng, "probe_multi", None) is None:
                        its = [self.eng.probe(mine[b], nd.proj, nd.const, seed=sd, **kw) for b, nd, sd in zip(have, nodes, seeds)]
                    else:
                        its = self.eng.probe_multi([mine[b] for b in have], [nd.proj for nd in nodes],
                                                   [nd.const for nd in nodes], seeds, **kw)
                    for b, it in zip(have, its):
                        totals[b] = float(it.sum())
                totals = self.comm.allreduce_sum(totals)
                for b, j in enumerate(batch):
                    T_of[j] = int(float(totals[b]) / cellsToSample) - convDistAvgNum

            for at, i in enumerate(sched):
                p, draw = live[i], draws[i]
                node, mem = p["node"], p["members"]
                if fixed_iterations is None:
                    if i not in T_of:
                        probe_ready(i)
                    T = T_of[i]
                else:
                    T = int(fixed_iterations)
                stepped.add(int(node.identifier))
                self.node_iterations[int(node.identifier)] = T
                ids, nmem = p["ids"], len(p["mem_loc"])
                steps = np.full(len(ids), T, dtype=np.int32)                     # descendants step T (MS:599-605)
                if cellDevelopmentMode:
                    np.minimum((draw.result() * (T + 1)).astype(np.int32), T, out=steps[:nmem])   # members k ~ U{0..T} (MS:589)
                else:
                    steps[:nmem] = poisson_quantile(draw.result(), T)                          # members k ~ Poisson(T) (MS:591)
                base = base_f.result() if base_f is not None else iters[ids]
                next_ids = live[sched[at + 1]]["ids"] if at + 1 < len(sched) else None
                base_f = pool.submit(bookkeeping, ids, steps, next_ids)
                if len(ids) and T > 0:
                    self.eng.run_phase(ids, steps, node.proj, node.const, step_base=base, dt=timeStep,
                                       sigma=noiseDisp, mu=geneMeanLevels, seed=self.seed)
                    h2d += len(ids) * 20 + 16 * len(node.proj)
                local_steps += int(steps.sum(dtype=np.int64))
            if base_f is not None:
                base_f.result()
            typed.result()
            size = None
            if do_counts:
                shift = self.gene_shift(expScale=expScale)
                size = size_f.result()
                nblk = max(1, min(int(counts_blocks), nloc))
                edges = [nloc * b // nblk for b in range(nblk + 1)]
                n_genes = len(shift)
                esz = {"int32": 4, "uint16": 2, "uint8": 1}[counts_dtype]
                overflow = []
                for r0, r1 in zip(edges[:-1], edges[1:]):
                    out_blk = None if counts_out is None else counts_out[r0:r1]
                    dev_blk = None if counts_dev_ptr is None else counts_dev_ptr + esz * n_genes * r0
                    if counts_dtype == "int32":
                        self.eng.counts(shift, size[r0:r1], exp_scale=expScale, seed=self.seed, cell_begin=r0,
                                        cell_count=r1 - r0, out=out_blk, out_device_ptr=dev_blk)
                    else:
                        _, ovf = self.eng.counts_compact(shift, size[r0:r1], exp_scale=expScale, seed=self.seed,
                                                         dtype=counts_dtype, cell_begin=r0, cell_count=r1 - r0, out=out_blk,
                                                         out_device_ptr=dev_blk)
                        ovf[:, 0] += r0
                        overflow.append(ovf)
                    if on_counts_block is not None:
                        on_counts_block(r0, r1)
                h2d += shift.nbytes + size.nbytes
        finally:
            pool.shutdown(wait=True)
        glob = float(self.comm.allreduce_sum(np.array([float(local_steps)]))[0])
        ovf_all = None
        if do_counts and counts_dtype != "int32":
            ovf_all = np.concatenate(overflow) if overflow else np.zeros((0, 3), dtype=np.int64)
        return dict(global_cell_steps=glob, local_cell_steps=local_steps, h2d_bytes=h2d, iterations=iters,
                    types=types, size_factors=size, overflow=ovf_all)


def poisson_quantile(u, lam):
    """Smallest k with P(Poisson(lam) <= k) > u, vectorised over u in [0, 1): inverse-transform draw of Poisson(lam)
    from one uniform per cell (the prepared mode's counterpart of np.random.poisson at MS:591)."""
    u = np.asarray(u, dtype=np.float64)
    if lam <= 0:
        return np.zeros(u.shape, dtype=np.int32)
    kmax = int(lam + 12.0 * np.sqrt(lam) + 25)
    k = np.arange(kmax + 1, dtype=np.float64)
    from math import lgamma
    logp = k * np.log(lam) - lam - np.array([lgamma(v + 1.0) for v in k])
    cdf = np.cumsum(np.exp(logp))
    return np.minimum(np.searchsorted(cdf, u, side="right"), kmax).astype(np.int32)


def hashed_uniform(seed, salt, ids):
    """U[0,1) as a pure function of (seed, salt, id): splitmix64 finaliser, 53 bits."""
    with np.errstate(over="ignore"):
        x = np.asarray(ids, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
        x = x + np.uint64((int(seed) * 0xD1342543DE82EF95 + int(salt) * 0xBF58476D1CE4E5B9) % (2 ** 64))
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
