"""ctypes binding of ``libpeakachu_b200.so`` (include/peakachu_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is
visible, the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PEAKACHU_B200_LIB points at another build of the same library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("PEAKACHU_B200_LIB") or os.path.join(_HERE, "libpeakachu_b200.so")

PK_MEM_HOST, PK_MEM_DEVICE = 0, 1
PK_PIXELS_SORTED = 0x100
PK_ENC_COO, PK_ENC_CSR32, PK_ENC_CSR16, PK_ENC_ROWS = 0, 1, 2, 3

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)
c_f32p = C.POINTER(C.c_float)
c_u8p = C.POINTER(C.c_uint8)



class Unit(C.Structure):
    """pk_unit (include/peakachu_b200.h)."""
    _fields_ = [("tag", C.c_int64), ("n_bins", C.c_int32), ("row_begin", C.c_int32), ("row_end", C.c_int32),
                ("encoding", C.c_int32), ("a", C.c_void_p), ("b", C.c_void_p), ("c", C.c_void_p),
                ("size", C.c_int64), ("weights", C.c_void_p), ("poisson_weights", C.c_void_p), ("min_prob", C.c_double)]


class UnitResult(C.Structure):
    """pk_unit_result (include/peakachu_b200.h)."""
    _fields_ = [("tag", C.c_int64), ("n_bins", C.c_int32), ("row_begin", C.c_int32), ("row_end", C.c_int32),
                ("whole", C.c_int32), ("n_records", C.c_int64), ("n_candidates", C.c_int64),
                ("n_windows", C.c_int64), ("n_batches", C.c_int64), ("x", C.c_void_p), ("y", C.c_void_p),
                ("batch", C.c_void_p), ("prob", C.c_void_p), ("value", C.c_void_p), ("batch_windows", C.c_void_p)]


# name -> (restype, argtypes); every symbol declared in include/peakachu_b200.h
SIGNATURES = {
    "pk_last_error": (C.c_char_p, []),
    "pk_abi_version": (C.c_int, []),
    "pk_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "pk_device_pci_bus_id": (C.c_int, [C.c_int, C.c_char_p, C.c_int]),
    "pk_forest_create": (C.c_int, [C.c_int, C.c_int32, C.c_int32, c_i64p, c_i32p, c_f64p, c_i32p, c_i32p,
                                   c_u8p, c_f64p, C.POINTER(C.c_void_p)]),
    "pk_forest_destroy": (C.c_int, [C.c_void_p]),
    "pk_forest_info": (C.c_int, [C.c_void_p, c_i32p, c_i32p, c_i64p]),
    "pk_forest_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pk_chrom_create": (C.c_int, [C.c_int, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int, C.c_void_p,
                                  C.POINTER(C.c_void_p)]),
    "pk_chrom_destroy": (C.c_int, [C.c_void_p]),
    "pk_chrom_bounds": (C.c_int, [C.c_void_p, c_i32p, c_i32p, c_i32p]),
    "pk_chrom_upload_pixels": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]),
    "pk_chrom_upload_csr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]),
    "pk_chrom_upload_csr16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]),
    "pk_chrom_set_poisson_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "pk_chrom_upload_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]),
    "pk_engine_create": (C.c_int, [C.c_int, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int, C.POINTER(C.c_void_p)]),
    "pk_engine_destroy": (C.c_int, [C.c_void_p]),
    "pk_engine_submit": (C.c_int, [C.c_void_p, C.POINTER(Unit)]),
    "pk_engine_collect": (C.c_int, [C.c_void_p, C.POINTER(UnitResult), C.c_int64, c_i64p]),
    "pk_engine_reset": (C.c_int, [C.c_void_p]),
    "pk_release_memory": (C.c_int, []),
    "pk_format_bedpe": (C.c_int, [C.c_char_p, C.c_int64, c_i32p, c_i32p, c_f64p, c_f64p, C.c_int64, C.c_char_p,
                                  C.c_int64, c_i64p]),
    "pk_stream_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "pk_h5_decode_chunks": (C.c_int, [C.c_void_p, C.c_int64, c_i64p, c_i64p, c_i64p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_int32]),
    "pk_rows_pack": (C.c_int, [c_i64p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32,
                               C.c_void_p, C.c_void_p, C.c_int64, c_i64p, C.c_int32]),
    "pk_stream_destroy": (C.c_int, [C.c_int, C.c_void_p]),
    "pk_stream_create_priority": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "pk_chrom_set_score_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pk_selftest_divide": (C.c_int, [C.c_int, C.c_int64, C.c_uint64, c_i64p]),
    "pk_chrom_depth": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int64)]),
    "pk_chrom_diag_sums": (C.c_int, [C.c_void_p, c_f64p, c_i64p]),
    "pk_chrom_fit_expected": (C.c_int, [C.c_void_p]),
    "pk_chrom_set_expected": (C.c_int, [C.c_void_p, c_f64p, c_f64p]),
    "pk_chrom_get_expected": (C.c_int, [C.c_void_p, c_f64p]),
    "pk_chrom_find_candidates": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_i64p]),
    "pk_chrom_candidates": (C.c_int, [C.c_void_p, c_i32p, c_i32p, C.c_int64, c_i64p]),
    "pk_chrom_features": (C.c_int, [C.c_void_p, c_u8p, c_f32p, c_f64p, C.c_int64]),
    "pk_chrom_fused_features": (C.c_int, [C.c_void_p, C.c_void_p, c_u8p, c_f32p, C.c_int64]),
    "pk_chrom_features_at": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, c_u8p, c_f32p, c_f64p]),
    "pk_chrom_score": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double]),
    "pk_chrom_result_count": (C.c_int, [C.c_void_p, c_i64p, c_i64p, c_i64p]),
    "pk_chrom_batch_windows": (C.c_int, [C.c_void_p, c_i64p, C.c_int64, c_i64p]),
    "pk_chrom_fetch_results": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int64, C.c_int]),
    "pk_poisson_critical_mu": (C.c_int, [C.c_int32, c_f64p]),
    "pk_fit_expected": (C.c_int, [c_f64p, c_i64p, C.c_int32, c_f64p]),
    "pk_set_tuning": (C.c_int, [C.c_char_p, C.c_int]),
    "pk_chrom_stage_ms": (C.c_int, [C.c_void_p, c_f32p]),
}


class PKError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("peakachu_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    """Load the CUDA library once; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C peakachu_b200/csrc`). There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise PKError(rc, lib().pk_last_error().decode("utf-8", "replace"))


def require_device():
    n = C.c_int(0)
    check(lib().pk_device_count(C.byref(n)))
    return n.value


def ptr(a, typ=None):
    """ctypes pointer to a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(typ) if typ is not None else C.c_void_p(a.ctypes.data)


def as_c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


# ---------------------------------------------------------------------------
# The expected-curve fit on the device (k_fit_expected) restates, operation for operation, what the
# reference gets from its environment: scipy.optimize.isotonic_regression's PAVA, scikit-learn's
# IsotonicRegression knot trimming and numpy.interp (utils.py:173-176). That arithmetic is pinned to the
# versions the golden fixtures were made with; other versions of those packages may round differently.
# ---------------------------------------------------------------------------
PINNED_VERSIONS = {"sklearn": "1.9", "scipy": "1.18", "numpy": "2.3"}
_expected_mode = None


def _major_minor(v):
    return ".".join(str(v).split(".")[:2])


def installed_versions():
    out = {}
    for mod in PINNED_VERSIONS:
        try:
            out[mod] = _major_minor(__import__(mod).__version__)
        except Exception:
            out[mod] = None
    return out


def expected_mode():
    """Where the expected curve is fitted: "device" (the library's own fit) or "host" (the installed
    scikit-learn, handed over with pk_chrom_set_expected). PEAKACHU_B200_EXPECTED = device | host | auto;
    auto (the default) takes the device and warns once when the environment's scikit-learn / scipy / numpy
    are not the versions the device arithmetic is pinned to."""
    global _expected_mode
    if _expected_mode is None:
        want = os.environ.get("PEAKACHU_B200_EXPECTED", "auto").lower()
        if want not in ("auto", "device", "host"):
            raise ValueError("PEAKACHU_B200_EXPECTED must be auto, device or host")
        if want == "auto":
            have = installed_versions()
            off = {m: v for m, v in have.items() if v is not None and v != PINNED_VERSIONS[m]}
            if off:
                import warnings
                warnings.warn("peakachu_b200 fits the expected curve on the GPU with the arithmetic of %s; this environment has %s, "
                              "whose IsotonicRegression may round the last bits differently. Set PEAKACHU_B200_EXPECTED=host to fit "
                              "the curve with the installed scikit-learn instead (slower: one host round trip per chromosome)."
                              % (", ".join("%s %s" % kv for kv in PINNED_VERSIONS.items()),
                                 ", ".join("%s %s" % kv for kv in off.items())), RuntimeWarning, stacklevel=2)
            want = "device"
        _expected_mode = want
    return _expected_mode
