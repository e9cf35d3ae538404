"""ctypes binding of ``libscrutiny_b200.so`` (the C ABI declared in include/scrutiny_b200.h).

The library is built in-tree (``scrutiny_b200/csrc/Makefile`` or ``__graft_entry__.build()``).
There is no fallback of any kind: a missing library raises immediately.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SCRUTINY_B200_LIB: developer override used by A/B kernel experiments (scripts/ab_build.sh)
LIB_PATH = os.environ.get("SCRUTINY_B200_LIB") or os.path.join(_HERE, "libscrutiny_b200.so")

c_ctx = C.c_void_p
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u32p = C.POINTER(C.c_uint32)
_f64p = C.POINTER(C.c_double)

# name -> (restype, argtypes); mirrors include/scrutiny_b200.h one to one
SIGNATURES = {
    "scr_abi_version": (C.c_int, []),
    "scr_noise_stream_version": (C.c_int, []),
    "scr_create": (C.c_int, [C.POINTER(c_ctx), C.c_int, C.c_int]),
    "scr_destroy": (None, [c_ctx]),
    "scr_last_error": (C.c_char_p, [c_ctx]),
    "scr_sync": (C.c_int, [c_ctx]),
    "scr_set_network_csr": (C.c_int, [c_ctx, C.c_int32, C.c_int64, _i32p, _i32p, _f64p]),
    "scr_set_network_dense": (C.c_int, [c_ctx, C.c_int32, _f64p]),
    "scr_set_drive_bias": (C.c_int, [c_ctx, _f64p]),
    "scr_debug_host_matvec": (C.c_int, [c_ctx, _f64p, _f64p]),
    "scr_debug_block_records": (C.c_int, [c_ctx, _i32p, C.c_int32, _i32p, C.c_int32]),
    "scr_debug_gene_order": (C.c_int, [c_ctx, _i32p, C.c_int32]),
    "scr_get_layout": (C.c_int, [c_ctx, _i64p, C.c_int]),
    "scr_alloc_cells": (C.c_int, [c_ctx, C.c_int64, C.c_int64]),
    "scr_set_states": (C.c_int, [c_ctx, C.c_int64, C.c_int64, _f64p, C.c_int64]),
    "scr_set_states_broadcast": (C.c_int, [c_ctx, _f64p]),
    "scr_get_states": (C.c_int, [c_ctx, C.c_int64, C.c_int64, _f64p, C.c_int64]),
    "scr_run_phase": (C.c_int, [c_ctx, C.c_int64, _i64p, _i32p, _i64p, _f64p, _f64p, C.c_double,
                                C.c_double, C.c_double, C.c_uint64, _f64p, C.c_int64]),
    "scr_probe": (C.c_int, [c_ctx, C.c_int32, _i64p, _f64p, _f64p, _f64p, C.c_double, C.c_double,
                            C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_uint64, C.c_uint32,
                            _f64p, _i32p, _f64p]),
    "scr_probe_multi": (C.c_int, [c_ctx, C.c_int32, _i32p, _i64p, _f64p, _f64p, C.c_double, C.c_double, C.c_double,
                                  C.c_double, C.c_int32, C.c_int32, C.POINTER(C.c_uint64), C.c_uint32, _i32p]),
    "scr_gene_sums": (C.c_int, [c_ctx, _f64p]),
    "scr_counts": (C.c_int, [c_ctx, _f64p, C.c_double, _f64p, C.c_uint64, _f64p, C.c_int32,
                             C.c_void_p, C.c_int, _f64p]),
    "scr_counts_range": (C.c_int, [c_ctx, C.c_int64, C.c_int64, _f64p, C.c_double, _f64p, C.c_uint64, _f64p, C.c_int32,
                                   C.c_void_p, C.c_int, _f64p]),
    "scr_counts_compact": (C.c_int, [c_ctx, C.c_int64, C.c_int64, _f64p, C.c_double, _f64p, C.c_uint64, C.c_int, C.c_void_p,
                                     C.c_int, _i64p, C.c_int64, _i64p]),
    "scr_counts_ext": (C.c_int, [c_ctx, C.c_int64, C.c_int64, _f64p, C.c_double, _f64p, C.c_uint64, C.c_double, C.c_double,
                                 _f64p, C.c_int32, C.c_int, C.c_void_p, C.c_int, _f64p, _i64p, C.c_int64, _i64p]),
    "scr_set_genes": (C.c_int, [c_ctx, C.c_int32]),
    "scr_debug_philox": (C.c_int, [c_ctx, C.c_int64, _u32p, C.c_uint32, C.c_uint32, _u32p]),
    "scr_debug_noise": (C.c_int, [c_ctx, C.c_int64, C.c_uint32, C.c_double, C.c_uint64, C.c_uint32, _f64p]),
    "scr_launch_count": (C.c_int64, [c_ctx]),
    "scr_debug_path_launches": (C.c_int64, [c_ctx, C.c_int]),
    "scr_last_kernel_ms": (C.c_double, [c_ctx]),
    "scr_step_kernel_stats": (C.c_int, [c_ctx, _f64p, _i64p, C.c_int]),
    "scr_probe_kernel_stats": (C.c_int, [c_ctx, _f64p, _i64p, C.c_int]),
    "scr_rank_transform": (C.c_int, [c_ctx, C.c_int64, C.c_int64, _f64p, C.c_int64]),
    "scr_rank_transform_stats": (C.c_int, [c_ctx, _f64p, _i64p, _i64p]),
    "scr_cell_pairs": (C.c_int, [c_ctx, C.c_int64, C.c_int64, _f64p, C.c_int64, _f64p, _i64p, C.c_double,
                                 _i64p, C.c_int64, _i64p]),
}

_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                "scrutiny_b200: %s is missing -- build it with `make -C scrutiny_b200/csrc` "
                "(or __graft_entry__.build()). There is no CPU fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
