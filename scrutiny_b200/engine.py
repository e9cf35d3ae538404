"""Thin object wrapper over the C ABI: one ``Engine`` = one ``scr_ctx`` = one GPU.

This is host plumbing only (argument marshalling, error codes -> exceptions); every
computation happens in the CUDA library.
"""
import ctypes as C

import numpy as np

from . import _lib
from .network import dense_to_csr

F32, F64 = 0, 1
COUNT_DTYPES = {"int32": (0, np.int32), "uint16": (1, np.uint16), "uint8": (2, np.uint8)}
E_RESOURCE = -4
STREAM_STEP, STREAM_PROBE, STREAM_COUNTS, STREAM_ALPHA = 0, 1, 2, 16


class ScrutinyError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("scrutiny_b200 error %d: %s" % (code, msg))
        self.code = code


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype)) if a is not None else None


class Engine:
    def __init__(self, device=0, precision="f32"):
        self._lib = _lib.load()
        self._ctx = _lib.c_ctx()
        self.precision = precision
        prec = {"f32": F32, "f64": F64}[precision]
        rc = self._lib.scr_create(C.byref(self._ctx), int(device), prec)
        if rc != 0:
            msg = self._lib.scr_last_error(None)
            raise ScrutinyError(rc, msg.decode() if msg else "scr_create failed")
        self.n = None
        self.cells = 0

    # -- plumbing ---------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            msg = self._lib.scr_last_error(self._ctx)
            raise ScrutinyError(rc, msg.decode() if msg else "")

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.scr_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- network ----------------------------------------------------------------------------
    def set_network(self, gc=None, csr=None, n=None):
        """``gc``: dense (n, n) float64 as the reference passes it (already divided by alpha),
        or ``csr=(rowptr, col, val)``."""
        if csr is None:
            if not hasattr(gc, "tocsr"):            # scipy sparse matrices go straight to CSR
                gc = np.asarray(gc)
            n = gc.shape[0]
            csr = dense_to_csr(gc)
        rowptr, col, val = csr
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        col = np.ascontiguousarray(col, dtype=np.int32)
        val = _f64(val)
        n = int(n if n is not None else len(rowptr) - 1)
        self._check(self._lib.scr_set_network_csr(self._ctx, n, len(val), _ptr(rowptr, C.c_int32),
                                                  _ptr(col, C.c_int32), _ptr(val, C.c_double)))
        self.n = n
        self._csr = (rowptr, col, val)

    def _mean_levels(self, mu, const):
        """geneMeanLevels as the reference accepts it (a scalar, or a per-gene vector broadcast in ``x - geneMeanLevels``,
        MS:530) -> (scalar mu for the ABI, const).  A non-uniform vector becomes the drive bias -gc mu plus const + mu."""
        if np.ndim(mu) == 0:
            return float(mu), const
        mu = np.asarray(mu, dtype=np.float64).ravel()
        if mu.shape != (self.n,):
            raise ValueError("geneMeanLevels must be a scalar or a vector of length n = %d, got shape %s" % (self.n, mu.shape))
        if np.all(mu == mu[0]):
            return float(mu[0]), const
        rowptr, col, val = self._csr
        rows = np.repeat(np.arange(self.n), np.diff(rowptr))
        bias = -np.bincount(rows, weights=val * mu[col], minlength=self.n)
        self._check(self._lib.scr_set_drive_bias(self._ctx, _ptr(_f64(bias), C.c_double)))
        self._bias_set = True
        return 0.0, const + mu

    def _clear_bias(self):
        if getattr(self, "_bias_set", False):
            self._check(self._lib.scr_set_drive_bias(self._ctx, None))
            self._bias_set = False

    def set_genes(self, n):
        """Read-out context: n genes, no network (only states, gene sums and counts work afterwards)."""
        self._check(self._lib.scr_set_genes(self._ctx, int(n)))
        self.n = int(n)

    def host_matvec(self, x):
        x = _f64(x)
        y = np.zeros(self.n)
        self._check(self._lib.scr_debug_host_matvec(self._ctx, _ptr(x, C.c_double), _ptr(y, C.c_double)))
        return y

    def block_records(self):
        """(records (blocks, 4) int32 = first gene / first slot / end slot / flags, list bounds per warp of a group)."""
        rec = np.zeros((256, 4), dtype=np.int32)
        bounds = np.zeros(17, dtype=np.int32)
        nb = self._lib.scr_debug_block_records(self._ctx, _ptr(rec, C.c_int32), 256, _ptr(bounds, C.c_int32), 17)
        if nb < 0:
            self._check(nb)
        wq = self.layout()["warps_per_group"]
        return rec[:nb].copy(), bounds[:wq + 1].copy()

    def gene_order(self):
        """perm[i] = original gene at internal position i (-1 = padding)."""
        perm = np.zeros(self.layout()["n_pad"], dtype=np.int32)
        self._check(self._lib.scr_debug_gene_order(self._ctx, _ptr(perm, C.c_int32), len(perm)))
        return perm

    def noise_stream_version(self):
        return int(self._lib.scr_noise_stream_version())

    def layout(self):
        out = np.zeros(18, dtype=np.int64)
        self._check(self._lib.scr_get_layout(self._ctx, _ptr(out, C.c_int64), 18))
        keys = ("n_pad", "pingpong_genes", "tile_slots", "chunks", "chunk_len", "row_cap",
                "warps_per_group", "groups_per_cta", "smem_bytes", "operator_in_smem",
                "cells_per_group", "dense_stepper", "operator_prefix_bytes", "operator_bytes",
                "gather_wavefronts_model", "gather_wavefronts_floor", "state_in_global",
                "dense_encoding")                                   # 0 = sparse stepper, 1 = fp16 split, 2 = 3xTF32
        return dict(zip(keys, (int(v) for v in out)))

    # -- states -----------------------------------------------------------------------------
    def alloc_cells(self, cells, global_offset=0):
        self._check(self._lib.scr_alloc_cells(self._ctx, int(cells), int(global_offset)))
        self.cells = int(cells)

    def set_states(self, S, cell0=0):
        """S: (n, count) float64 view (row stride may exceed count)."""
        S = np.asarray(S, dtype=np.float64)
        if S.strides[1] != 8:
            S = np.ascontiguousarray(S)
        self._check(self._lib.scr_set_states(self._ctx, int(cell0), S.shape[1], _ptr(S, C.c_double),
                                             S.strides[0] // 8))

    def set_states_broadcast(self, x0):
        x0 = _f64(x0)
        self._check(self._lib.scr_set_states_broadcast(self._ctx, _ptr(x0, C.c_double)))

    def get_states(self, cell0=0, count=None, out=None):
        count = self.cells - cell0 if count is None else count
        if out is None:
            out = np.empty((self.n, count))
        assert out.dtype == np.float64 and out.strides[1] == 8
        self._check(self._lib.scr_get_states(self._ctx, int(cell0), int(count), _ptr(out, C.c_double),
                                             out.strides[0] // 8))
        return out

    # -- hot path ---------------------------------------------------------------------------
    def run_phase(self, cell_ids, steps, proj, const, step_base=None, dt=0.01, sigma=0.05, mu=0.0,
                  seed=0, noise=None):
        cell_ids = np.ascontiguousarray(cell_ids, dtype=np.int64)
        steps = np.ascontiguousarray(steps, dtype=np.int32)
        base = None if step_base is None else np.ascontiguousarray(step_base, dtype=np.int64)
        proj, const = _f64(proj), _f64(const)
        mu, const = self._mean_levels(mu, const)
        const = _f64(const)
        nsteps = 0
        if noise is not None:
            noise = _f64(noise)
            assert noise.ndim == 3 and noise.shape[0] == len(cell_ids) and noise.shape[2] == self.n
            nsteps = noise.shape[1]
        try:
            self._check(self._lib.scr_run_phase(self._ctx, len(cell_ids), _ptr(cell_ids, C.c_int64),
                                                _ptr(steps, C.c_int32), _ptr(base, C.c_int64),
                                                _ptr(proj, C.c_double), _ptr(const, C.c_double), dt, sigma, mu,
                                                int(seed) & (2 ** 64 - 1), _ptr(noise, C.c_double), nsteps))
        finally:
            self._clear_bias()

    def probe(self, sample_ids, proj, const, dt=0.01, sigma=0.05, mu=0.0, thresh=0.1, max_iter=500,
              avg_num=20, seed=0, stream=STREAM_PROBE, w_div=None, noise=None, want_trace=False):
        ids = np.ascontiguousarray(sample_ids, dtype=np.int64)
        k = len(ids)
        proj, const = _f64(proj), _f64(const)
        mu, const = self._mean_levels(mu, const)
        const = _f64(const)
        wd = None if w_div is None else _f64(w_div)
        nz = None
        if noise is not None:
            nz = _f64(noise)
            assert nz.shape == (k, max_iter, self.n)
        iters = np.zeros(k, dtype=np.int32)
        trace = np.zeros((k, max_iter)) if want_trace else None
        try:
            self._check(self._lib.scr_probe(self._ctx, k, _ptr(ids, C.c_int64), _ptr(wd, C.c_double),
                                            _ptr(proj, C.c_double), _ptr(const, C.c_double), dt, sigma, mu, thresh,
                                            int(max_iter), int(avg_num), int(seed) & (2 ** 64 - 1), int(stream),
                                            _ptr(nz, C.c_double), _ptr(iters, C.c_int32), _ptr(trace, C.c_double)))
        finally:
            self._clear_bias()
        return (iters, trace) if want_trace else iters

    def probe_multi(self, samples, projs, consts, seeds, dt=0.01, sigma=0.05, mu=0.0, thresh=0.1, max_iter=500,
                    avg_num=20, stream=STREAM_PROBE):
        """``probe`` for several lineage nodes at once (scr_probe_multi): ``samples`` is a list of id arrays, ``projs`` /
        ``consts`` the nodes' vectors, ``seeds`` one per node.  Returns the list of iteration arrays -- what separate
        ``probe`` calls return; the launches of sibling nodes run side by side on the GPU."""
        if np.ndim(mu) != 0:                                   # per-gene mean levels: the drive bias is a context-wide setting
            return [self.probe(s, p, c, dt=dt, sigma=sigma, mu=mu, thresh=thresh, max_iter=max_iter, avg_num=avg_num,
                               seed=sd, stream=stream) for s, p, c, sd in zip(samples, projs, consts, seeds)]
        ids = [np.ascontiguousarray(s, dtype=np.int64) for s in samples]
        k = np.array([len(s) for s in ids], dtype=np.int32)
        allids = np.ascontiguousarray(np.concatenate(ids))
        P = np.ascontiguousarray(np.stack([_f64(p) for p in projs]))
        Cn = np.ascontiguousarray(np.stack([_f64(c) for c in consts]))
        sd = np.array([int(v) & (2 ** 64 - 1) for v in seeds], dtype=np.uint64)
        iters = np.zeros(len(allids), dtype=np.int32)
        self._check(self._lib.scr_probe_multi(self._ctx, len(ids), _ptr(k, C.c_int32), _ptr(allids, C.c_int64),
                                              _ptr(P, C.c_double), _ptr(Cn, C.c_double), dt, sigma, float(mu), thresh,
                                              int(max_iter), int(avg_num), sd.ctypes.data_as(C.POINTER(C.c_uint64)),
                                              int(stream), _ptr(iters, C.c_int32)))
        return np.split(iters, np.cumsum(k)[:-1])

    def gene_sums(self):
        out = np.zeros(self.n)
        self._check(self._lib.scr_gene_sums(self._ctx, _ptr(out, C.c_double)))
        return out

    def counts(self, shift, size_factors, exp_scale=1.0, seed=0, uniforms=None, out=None, out_device_ptr=None,
               want_lambda=False, cell_begin=0, cell_count=None):
        """Returns int32 (cells, n) (host) unless ``out_device_ptr`` (int) is given.  ``cell_begin`` / ``cell_count``
        restrict the call to a block of local cells (scr_counts_range): ``size_factors``, ``uniforms``, ``out`` /
        ``out_device_ptr`` then describe that block's rows only."""
        rows = self.cells - int(cell_begin) if cell_count is None else int(cell_count)
        shift, size_factors = _f64(shift), _f64(size_factors)
        assert len(size_factors) == rows and len(shift) == self.n
        uni, draws = None, 0
        if uniforms is not None:
            uni = _f64(uniforms)
            assert uni.shape[:2] == (rows, self.n)
            draws = uni.shape[2]
        lam = np.zeros((rows, self.n)) if want_lambda else None
        if out_device_ptr is not None:
            ptr, on_dev = C.c_void_p(int(out_device_ptr)), 1
        else:
            if out is None:
                out = np.empty((rows, self.n), dtype=np.int32)
            assert out.dtype == np.int32 and out.flags.c_contiguous and out.shape == (rows, self.n)
            ptr, on_dev = C.c_void_p(out.ctypes.data), 0
        self._check(self._lib.scr_counts_range(self._ctx, int(cell_begin), rows, _ptr(shift, C.c_double), float(exp_scale),
                                               _ptr(size_factors, C.c_double), int(seed) & (2 ** 64 - 1),
                                               _ptr(uni, C.c_double), int(draws), ptr, on_dev, _ptr(lam, C.c_double)))
        return (out, lam) if want_lambda else out

    def counts_ext(self, shift, size_factors, nb_dispersion=0.0, dropout=0.0, exp_scale=1.0, seed=0, uniforms=None,
                   want_lambda=False, cell_begin=0, cell_count=None):
        """Read-out with the default-off extensions (scr_counts_ext): negative-binomial draws with dispersion
        ``nb_dispersion`` (variance lambda + phi lambda^2) and Bernoulli ``dropout``.  int32 (rows, n) host result."""
        rows = self.cells - int(cell_begin) if cell_count is None else int(cell_count)
        shift, size_factors = _f64(shift), _f64(size_factors)
        assert len(size_factors) == rows and len(shift) == self.n
        uni, draws = None, 0
        if uniforms is not None:
            uni = _f64(uniforms)
            assert uni.shape[:2] == (rows, self.n)
            draws = uni.shape[2]
        lam = np.zeros((rows, self.n)) if want_lambda else None
        out = np.empty((rows, self.n), dtype=np.int32)
        self._check(self._lib.scr_counts_ext(self._ctx, int(cell_begin), rows, _ptr(shift, C.c_double), float(exp_scale),
                                             _ptr(size_factors, C.c_double), int(seed) & (2 ** 64 - 1), float(nb_dispersion),
                                             float(dropout), _ptr(uni, C.c_double), int(draws), 0, C.c_void_p(out.ctypes.data), 0,
                                             _ptr(lam, C.c_double), None, 0, None))
        return (out, lam) if want_lambda else out

    def counts_compact(self, shift, size_factors, exp_scale=1.0, seed=0, dtype="uint8", out=None, out_device_ptr=None,
                       cell_begin=0, cell_count=None, overflow_capacity=None):
        """Free-running read-out stored as ``dtype`` ("uint8" / "uint16" / "int32"), scr_counts_compact.  Returns
        ``(out, overflow)``: ``out`` (rows, n) host array (None with ``out_device_ptr``), ``overflow`` (k, 3) int64 rows
        (row of this call, gene, count) for the entries stored as the escape code (dtype max), sorted by (row, gene)."""
        code, np_dtype = COUNT_DTYPES[dtype]
        rows = self.cells - int(cell_begin) if cell_count is None else int(cell_count)
        shift, size_factors = _f64(shift), _f64(size_factors)
        assert len(size_factors) == rows and len(shift) == self.n
        if out_device_ptr is not None:
            ptr, on_dev = C.c_void_p(int(out_device_ptr)), 1
        else:
            if out is None:
                out = np.empty((rows, self.n), dtype=np_dtype)
            assert out.dtype == np_dtype and out.flags.c_contiguous and out.shape == (rows, self.n)
            ptr, on_dev = C.c_void_p(out.ctypes.data), 0
        cap = int(overflow_capacity) if overflow_capacity is not None else max(4096, rows * self.n // 4096)
        while True:
            ovf = np.zeros((max(cap, 1), 3), dtype=np.int64)
            count = C.c_int64(0)
            rc = self._lib.scr_counts_compact(self._ctx, int(cell_begin), rows, _ptr(shift, C.c_double), float(exp_scale),
                                              _ptr(size_factors, C.c_double), int(seed) & (2 ** 64 - 1), code, ptr, on_dev,
                                              _ptr(ovf, C.c_int64), cap, C.byref(count))
            if rc == E_RESOURCE and count.value > cap:      # list too small: the draw is deterministic, sample again
                cap = int(count.value)
                continue
            self._check(rc)
            return out, ovf[:count.value].copy()

    # -- RegNetInference front end ------------------------------------------------------------
    def rank_transform(self, S):
        """In-place convertToS (RegNetInference.py:69-78) of a (genes, cells) float64 array whose rows may be
        strided (ld = row stride)."""
        assert S.dtype == np.float64 and S.ndim == 2 and S.strides[1] == 8 and S.strides[0] % 8 == 0
        n, cells = S.shape
        self._check(self._lib.scr_rank_transform(self._ctx, n, cells, _ptr(S, C.c_double), S.strides[0] // 8))
        return S

    def rank_transform_stats(self):
        ms, a, b = C.c_double(0.0), C.c_int64(0), C.c_int64(0)
        self._check(self._lib.scr_rank_transform_stats(self._ctx, C.byref(ms), C.byref(a), C.byref(b)))
        return {"kernel_ms": ms.value, "counting_cells": a.value, "sorted_cells": b.value}

    def cell_pairs(self, S, pseudotimes, cell_types, thresh):
        """(pairs, 2) int64 array in the reference's order (RegNetInference.py:397-416)."""
        assert S.dtype == np.float64 and S.ndim == 2 and S.strides[1] == 8 and S.strides[0] % 8 == 0
        n, cells = S.shape
        pt = _f64(pseudotimes)
        ty = np.ascontiguousarray(cell_types, dtype=np.int64)
        assert pt.shape == (cells,) and ty.shape == (cells,)
        count = C.c_int64(0)
        args = (self._ctx, n, cells, _ptr(S, C.c_double), S.strides[0] // 8, _ptr(pt, C.c_double),
                _ptr(ty, C.c_int64), float(thresh))
        self._check(self._lib.scr_cell_pairs(*args, None, 0, C.byref(count)))
        pairs = np.empty((count.value, 2), dtype=np.int64)
        if count.value:
            self._check(self._lib.scr_cell_pairs(*args, _ptr(pairs, C.c_int64), count.value, C.byref(count)))
            assert count.value == len(pairs)
        return pairs

    # -- debug ------------------------------------------------------------------------------
    def debug_philox(self, ctr4, k0, k1):
        ctr = np.ascontiguousarray(ctr4, dtype=np.uint32).reshape(-1, 4)
        out = np.zeros_like(ctr)
        self._check(self._lib.scr_debug_philox(self._ctx, len(ctr), _ptr(ctr, C.c_uint32), int(k0), int(k1),
                                               _ptr(out, C.c_uint32)))
        return out

    def debug_noise(self, global_cell, step, sigma=0.05, seed=0, stream=STREAM_STEP):
        out = np.zeros(self.n)
        self._check(self._lib.scr_debug_noise(self._ctx, int(global_cell), int(step), float(sigma),
                                              int(seed) & (2 ** 64 - 1), int(stream), _ptr(out, C.c_double)))
        return out

    def launch_count(self):
        return int(self._lib.scr_launch_count(self._ctx))

    def path_launches(self):
        return {"sparse_persistent": int(self._lib.scr_debug_path_launches(self._ctx, 0)),
                "dense_gemm": int(self._lib.scr_debug_path_launches(self._ctx, 1))}

    def step_kernel_stats(self, reset=False):
        ms = C.c_double(0)
        n = C.c_int64(0)
        self._check(self._lib.scr_step_kernel_stats(self._ctx, C.byref(ms), C.byref(n), int(reset)))
        return ms.value, n.value

    def probe_kernel_stats(self, reset=False):
        ms = C.c_double(0)
        n = C.c_int64(0)
        self._check(self._lib.scr_probe_kernel_stats(self._ctx, C.byref(ms), C.byref(n), int(reset)))
        return ms.value, n.value
