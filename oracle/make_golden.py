"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

Every fixture stores the seed, the call arguments and the reference's outputs.  The
tests replay the same seeded call sequence through ``oracle/matriseq_port.py`` (which
must reproduce the outputs bit for bit) and through the CUDA path (supplied-noise
mode, noise recorded from the seeded port run).  Environment the fixtures were made
with: python 3.12.3, numpy 2.3.5 (legacy RandomState streams are frozen by NEP 19).
"""
import os
import random
import sys

import numpy as np

from oracle import reference_harness as rh

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def seed_all(ms, seed):
    ms.init(outputDir=None, seed=seed, verbose=False)


def csr_triplets(gc):
    r, c = np.nonzero(gc)
    return r.astype(np.int32), c.astype(np.int32), gc[r, c]


def case_networks(ms):
    out = {}
    seed_all(ms, 11)
    gc, tf = ms.generateGeneCorrelationMatrix(60, "SimulatedPowerLaw", powerLawExponent=2.5)
    out["pl_gc"], out["pl_tf"] = gc, tf
    seed_all(ms, 12)
    gc, tf = ms.generateGeneCorrelationMatrix(50, "Random", networkSparsity=0.9, normalizeWRows=True)
    out["rnd_gc"], out["rnd_tf"] = gc, tf
    seed_all(ms, 13)
    gc, tf = ms.generateGeneCorrelationMatrix(5, "TRRUST", maxAlphaBeta=6)
    r, c, v = csr_triplets(gc)
    out["tr_n"] = np.int64(gc.shape[0])
    out["tr_rows"], out["tr_cols"], out["tr_vals"], out["tr_tf"] = r, c, v, tf
    np.savez_compressed(os.path.join(OUT, "networks.npz"), **out)


def case_step(ms):
    """developCell (MS:503-537) one step, non-trivial proj/const/mu; noise recorded."""
    seed_all(ms, 21)
    n = 48
    gc, tf = ms.generateGeneCorrelationMatrix(n, "Random", networkSparsity=0.7)
    gc = gc / 0.5
    x0 = 2 * np.random.rand(n) - 1 + 0.3
    proj = (np.random.rand(n) > 0.2).astype(float)
    const = np.where(proj == 0, np.sign(np.random.rand(n) - 0.5), 0.0)
    st = np.random.get_state()
    noise = np.random.normal(loc=0, scale=0.05, size=n)       # what developCell will draw next
    np.random.set_state(st)
    x1 = ms.developCell(n, gc, 0.01, x0, proj, const, 0.05, 0.3)
    np.savez_compressed(os.path.join(OUT, "step.npz"), gc=gc, x0=x0, proj=proj, const=const,
                        noise=noise, x1=x1, dt=0.01, sigma=0.05, mu=0.3)


def case_cellsteps(ms):
    """findCellNextState (MS:683-736): 40 steps of one cell at gain 50 (alpha 0.02)."""
    seed_all(ms, 31)
    n = 96
    gc, tf = ms.generateGeneCorrelationMatrix(n, "SimulatedPowerLaw")
    gc = gc / 0.02
    x0 = 2 * np.random.rand(n) - 1
    proj = np.ones(n); const = np.zeros(n)
    proj[tf[:5]] = 0; const[tf[:5]] = [1, -1, 1, 1, -1]
    steps = 40
    st = np.random.get_state()
    np.random.uniform()                                       # MS:707
    noise = np.stack([np.random.normal(loc=0, scale=0.05, size=n) for _ in range(steps)])
    np.random.set_state(st)
    xT = ms.findCellNextState(gc, n, 0.01, 0, proj, const, x0, steps, 0.05, 0, False)
    np.savez_compressed(os.path.join(OUT, "cellsteps.npz"), gc=gc, x0=x0, proj=proj, const=const,
                        noise=noise, xT=xT, steps=steps, dt=0.01, sigma=0.05)


def case_probe(ms):
    """findOptimalIterations (MS:616-680) with small limits."""
    seed_all(ms, 41)
    n, cells = 64, 12
    gc, tf = ms.generateGeneCorrelationMatrix(n, "SimulatedPowerLaw")
    gc = gc / 0.3
    states = 2 * np.random.rand(n, cells) - 1
    proj = np.ones(n); const = np.zeros(n)
    members = list(range(1, cells))
    T = ms.findOptimalIterations(n, gc, 0.01, members, proj, const, 0, 0.35, 0.05, states, 120, 8, 6)
    np.savez_compressed(os.path.join(OUT, "probe.npz"), gc=gc, states=states, proj=proj, const=const,
                        members=np.array(members), T=T, seed=41, thresh=0.35, maxIter=120, avgNum=8,
                        numToSample=6)


def case_pipeline(ms):
    """The example.py call sequence (EX:7-85) at small size: 3 types, 56+50+60 cells."""
    seed_all(ms, 1337)
    n, cells = 40, 166
    gc, tf = ms.generateGeneCorrelationMatrix(n, "SimulatedPowerLaw", powerLawExponent=2.3)
    tree, projM, constM = ms.formCellTypeTree(n, cells, [-1, 0, 0], [56, 50, 60], [0.1, 0.15, 0.2],
                                              tf, [0, 0, 0])
    init = ms.generateInitialStates(n, cells)
    alpha_log_states = init.copy()
    alpha = ms.findOptimalAlpha(n, gc, tree, init, numToSample=10, maxNumIterations=120,
                                convDistAvgNum=10, convergenceThreshold=0.3)
    gc_unscaled = gc.copy()
    gc = gc / alpha
    states = init
    types = np.zeros(cells, dtype="int64")
    iters = np.zeros(cells, dtype="int64")
    ms.findAllNextStates(gc, tree.head(), tree, states, types, iters)
    final_states = states.copy()
    counts, size = ms.transformData(n, cells, states)
    members = {str(t): np.array(tree[t].cellTypeMembers) for t in range(3)}
    np.savez_compressed(os.path.join(OUT, "pipeline.npz"), gc_unscaled=gc_unscaled, tf=tf, projM=projM,
                        constM=constM, init=alpha_log_states, alpha=alpha, final_states=final_states,
                        types=types, iters=iters, counts=counts, size=size,
                        members0=members["0"], members1=members["1"], members2=members["2"],
                        head_children=np.array(tree.head().children),
                        child_members0=np.array(tree[0].childCellTypeMembers))


def case_transform_alt(ms):
    """transformData (MS:816-891), gammaDistMean=False branch and expScale != 1."""
    seed_all(ms, 51)
    n, cells = 20, 15
    S = np.tanh(np.random.normal(size=(n, cells)))
    c1, s1 = ms.transformData(n, cells, S.copy(), expScale=4.0, cellRadiusDisp=0.3)
    c2, s2 = ms.transformData(n, cells, S.copy(), gammaDistMean=False, geneScale=1.5, expLambda=0.7)
    np.savez_compressed(os.path.join(OUT, "transform_alt.npz"), S=S, c1=c1, s1=s1, c2=c2, s2=s2)


def case_poisson_mode(ms):
    """findAllNextStates with cellDevelopmentMode=False at the head's only child (quirk R1:
    kwargs reach only the node the user passes)."""
    seed_all(ms, 61)
    n, cells = 24, 52
    gc, tf = ms.generateGeneCorrelationMatrix(n, "Random", networkSparsity=0.8)
    gc = gc / 1.5
    tree, projM, constM = ms.formCellTypeTree(n, cells, [-1], [52], [0.0], tf, [0])
    states = ms.generateInitialStates(n, cells, sameInitialCell=False)
    init = states.copy()
    types = np.zeros(cells, dtype="int64"); iters = np.zeros(cells, dtype="int64")
    ms.findAllNextStates(gc, tree[0], tree, states, types, iters, cellDevelopmentMode=False,
                         maxNumIterations=90, convDistAvgNum=6, cellsToSample=7,
                         convergenceThreshold=0.4, noiseDisp=0.02, timeStep=0.05)
    np.savez_compressed(os.path.join(OUT, "poisson_mode.npz"), gc=gc, init=init, final=states,
                        types=types, iters=iters)


def case_regnet():
    """RegNetInference.convertToS (RNI:69-78) on count-like, continuous and heavily tied matrices, and
    getCVCost (RNI:755-797) -- its value pins the pair set of RNI:764-777 -- from the unmodified reference."""
    rni = rh.load_reference_rni()
    out = {}
    rs = np.random.RandomState(71)
    a = rs.poisson(1.5, size=(37, 53)).astype(np.float64)           # counts: many ties
    b = rs.standard_normal(size=(20, 31))                           # continuous: no ties
    c = rs.randint(0, 3, size=(16, 64)).astype(np.float64)          # three distinct values
    c[:, 5] = c[:, 9]                                               # two identical cells
    d = np.exp(rs.normal(2.0, 2.5, size=(45, 12))).round()         # large integer counts
    for name, m in (("a", a), ("b", b), ("c", c), ("d", d)):
        out["in_" + name] = m
        out["s_" + name] = rni.convertToS(m.shape[0], m.shape[1], m.copy())
    # pair search through getCVCost on the transformed matrix `c` (identical cells 5 and 9 present)
    S = out["s_c"]
    n, cells = S.shape
    pt = np.round(rs.rand(cells), 2)                                # rounded: equal pseudotimes occur
    pt[9] = pt[5] + 0.01
    types = rs.randint(0, 3, size=cells).astype(np.int64)
    types[9] = types[5]
    W = 0.3 * rs.standard_normal(size=(n, n))
    out["pt"], out["types"], out["W"] = pt, types, W
    th = np.array([0.05, 0.2, 0.61])
    out["thresh"] = th
    out["cv_cost"] = np.array([rni.getCVCost(t, W, S, pt, types) for t in th])
    np.savez_compressed(os.path.join(OUT, "regnet.npz"), **out)


# ---- round 2 additions: deeper lineages, two roots, per-gene mean levels -------------------------------------
TREE7 = dict(n=30, parents=[-1, 0, 0, 1, 1, 2, 2], num=[50, 52, 51, 55, 50, 53, 50], props=[0.05, 0.1, 0.1, 0.15, 0.15, 0.2, 0.2])
TWO_ROOTS = dict(n=26, parents=[-1, -1, 0, 1], num=[51, 50, 52, 50], props=[0.1, 0.1, 0.2, 0.2])


def case_pipeline7(ms):
    """The example.py call sequence on config c2's tree shape (4 leaves, 7 types, depth 3) with a 'Random' network and a
    per-gene geneMeanLevels VECTOR in generateInitialStates / findOptimalAlpha (MS:340, 368; the reference broadcasts
    x - geneMeanLevels at MS:530).  Children run with the defaults of MS:546-554 (quirk R1), so every type holds >= 50 cells."""
    seed_all(ms, 77)
    n, cells = TREE7["n"], sum(TREE7["num"])
    gc, tf = ms.generateGeneCorrelationMatrix(n, "Random", networkSparsity=0.85)
    tree, projM, constM = ms.formCellTypeTree(n, cells, TREE7["parents"], TREE7["num"], TREE7["props"], tf, [0] * 7)
    mu = 0.001 * np.cos(np.arange(n))    # MS:535 adds mu EVERY step: |mu| / dt must stay below the threshold
    init = ms.generateInitialStates(n, cells, geneMeanLevels=mu)
    init0 = init.copy()
    alpha = ms.findOptimalAlpha(n, gc, tree, init, numToSample=8, maxNumIterations=300, convDistAvgNum=10,
                                convergenceThreshold=0.25, geneMeanLevels=mu)
    gc_unscaled = gc.copy()
    gc = gc / alpha
    types = np.zeros(cells, dtype="int64")
    iters = np.zeros(cells, dtype="int64")
    ms.findAllNextStates(gc, tree.head(), tree, init, types, iters)
    final_states = init.copy()
    counts, size = ms.transformData(n, cells, init)
    out = dict(gc_unscaled=gc_unscaled, tf=tf, projM=projM, constM=constM, mu=mu, init=init0, alpha=alpha,
               final_states=final_states, types=types, iters=iters, counts=counts, size=size,
               head_children=np.array(tree.head().children))
    for t in range(7):
        out["members%d" % t] = np.array(tree[t].cellTypeMembers)
        out["below%d" % t] = np.array(tree[t].childCellTypeMembers, dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "pipeline7.npz"), **out)


def case_two_roots(ms):
    """Two root types (parents [-1, -1, 0, 1]): findOptimalAlpha probes with proj = const = 0 (MS:433-435, quirk R2), the
    walk visits both subtrees (MS:607-612)."""
    seed_all(ms, 88)
    n, cells = TWO_ROOTS["n"], sum(TWO_ROOTS["num"])
    gc, tf = ms.generateGeneCorrelationMatrix(n, "SimulatedPowerLaw", powerLawExponent=2.2, normalizeWRows=True)
    tree, projM, constM = ms.formCellTypeTree(n, cells, TWO_ROOTS["parents"], TWO_ROOTS["num"], TWO_ROOTS["props"], tf,
                                              [0] * 4)
    init = ms.generateInitialStates(n, cells, sameInitialCell=False)
    init0 = init.copy()
    alpha = ms.findOptimalAlpha(n, gc, tree, init, numToSample=6, maxNumIterations=200, convDistAvgNum=20)
    gc_unscaled = gc.copy()
    gc = gc / 0.3                                             # the walk itself at a gentler gain than the search's 0.02
    types = np.zeros(cells, dtype="int64")
    iters = np.zeros(cells, dtype="int64")
    ms.findAllNextStates(gc, tree.head(), tree, init, types, iters)
    out = dict(gc_unscaled=gc_unscaled, tf=tf, projM=projM, constM=constM, init=init0, alpha=alpha, final_states=init.copy(),
               types=types, iters=iters, head_children=np.array(tree.head().children))
    for t in range(4):
        out["members%d" % t] = np.array(tree[t].cellTypeMembers)
        out["below%d" % t] = np.array(tree[t].childCellTypeMembers, dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "two_roots.npz"), **out)


def case_mu_vector(ms):
    """findAllNextStates started AT a type node (so its keyword arguments apply, quirk R1) with a per-gene geneMeanLevels
    vector and non-default step / noise / thresholds."""
    seed_all(ms, 99)
    n, cells = 28, 40
    gc, tf = ms.generateGeneCorrelationMatrix(n, "Random", networkSparsity=0.8)
    gc = gc / 1.2
    tree, projM, constM = ms.formCellTypeTree(n, cells, [-1], [40], [0.2], tf, [0])
    mu = np.linspace(-0.003, 0.003, n)     # small for the same reason (MS:535)
    states = ms.generateInitialStates(n, cells, sameInitialCell=False, geneMeanLevels=mu)
    init = states.copy()
    types = np.zeros(cells, dtype="int64"); iters = np.zeros(cells, dtype="int64")
    ms.findAllNextStates(gc, tree[0], tree, states, types, iters, maxNumIterations=160, convDistAvgNum=8, cellsToSample=9,
                         convergenceThreshold=0.18, noiseDisp=0.03, timeStep=0.02, geneMeanLevels=mu)
    np.savez_compressed(os.path.join(OUT, "mu_vector.npz"), gc=gc, mu=mu, init=init, final=states, types=types, iters=iters,
                        proj=tree[0].proj, const=tree[0].const)


ROUND2_CASES = (case_pipeline7, case_two_roots, case_mu_vector)


def main():
    if not rh.available():
        sys.exit("reference not available; fixtures can only be regenerated in the build container")
    os.makedirs(OUT, exist_ok=True)
    ms = rh.load_reference()
    for fn in (case_networks, case_step, case_cellsteps, case_probe, case_pipeline,
               case_transform_alt, case_poisson_mode):
        fn(ms)
        print("wrote", fn.__name__)
    case_regnet()
    print("wrote case_regnet")
    for fn in ROUND2_CASES:
        fn(ms)
        print("wrote", fn.__name__)


if __name__ == "__main__":
    main()
