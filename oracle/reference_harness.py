"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference MatriSeq module.

Only usable in the build container (``/root/reference`` does not exist on the GPU
box).  Used by ``oracle/make_golden.py`` to generate the committed fixtures under
``tests/golden/`` and by the not-gpu tests that cross-check the restatements in
``oracle/`` against the live reference when it is present.

Recipe: register dummy ``matplotlib`` / ``rpy2`` modules,
put the reference package directory on ``sys.path`` (the module uses the py2
implicit relative import ``from tree import Tree``, MatriSeq.py:7), import
``MatriSeq`` and inject the helper the packaged module forgot to define
(``rangeTransformation``: Matri-seqFinal.py:406-408 is ``2*x-1``).
"""
import importlib
import io
import os
import sys
import tarfile
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("SCRUTINY_REFERENCE_ROOT", "/root/reference")
_PKG_DIR = os.path.join(REFERENCE_ROOT, "RNAscrutiny", "RNAscrutiny")
_SDIST = os.path.join(REFERENCE_ROOT, "RNAscrutiny", "dist", "RNAscrutiny-1.1.tar.gz")

_cached = None
_trrust = None


def available():
    return os.path.isfile(os.path.join(_PKG_DIR, "MatriSeq.py"))


def _stub(name):
    mod = types.ModuleType(name)
    sys.modules[name] = mod
    return mod


def load_trrust_adjacency():
    """The (2895, 2895) 0/1 adjacency MatriSeq.py:97-101 loads from
    ``TRRUST_Network.npy``; absent from the tree but present in the 1.1 sdist."""
    global _trrust
    if _trrust is None:
        with tarfile.open(_SDIST) as tf:
            member = [m for m in tf.getmembers() if m.name.endswith("TRRUST_Network.npy")][0]
            _trrust = np.load(io.BytesIO(tf.extractfile(member).read()))
    return _trrust


def load_reference():
    """Return the reference ``MatriSeq`` module object (imported once)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    for name in ("matplotlib", "matplotlib.pyplot", "rpy2", "rpy2.robjects",
                 "rpy2.robjects.numpy2ri"):
        if name not in sys.modules:
            _stub(name)
    sys.modules["rpy2.robjects"].r = None
    sys.modules["rpy2.robjects.numpy2ri"].numpy2ri = None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if _PKG_DIR not in sys.path:
        sys.path.insert(0, _PKG_DIR)
    ms = importlib.import_module("MatriSeq")
    ms.rangeTransformation = lambda x: 2 * x - 1

    real_load = np.load

    def patched_load(path, *a, **k):
        if isinstance(path, str) and path.endswith("TRRUST_Network.npy"):
            return load_trrust_adjacency().copy()
        return real_load(path, *a, **k)

    ms.np.load = patched_load  # ms.np is numpy itself: patch is global but harmless
    _cached = ms
    return ms


_cached_rni = None


def load_reference_rni():
    """Return the reference ``RegNetInference`` module (stub plotting modules; scipy / sklearn / networkx are
    present in the build container)."""
    global _cached_rni
    if _cached_rni is not None:
        return _cached_rni
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "seaborn"):
        if name not in sys.modules:
            _stub(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
    if _PKG_DIR not in sys.path:
        sys.path.insert(0, _PKG_DIR)
    _cached_rni = importlib.import_module("RegNetInference")
    return _cached_rni
