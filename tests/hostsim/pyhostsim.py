"""ctypes binding of tests/hostsim/libhostsim.so -- TEST HARNESS ONLY (see hostsim.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

from secphase_b200.flatbatch import CFlatBatch

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
LIB = os.path.join(_HERE, "libhostsim.so")
SRC = os.path.join(_HERE, "hostsim.cpp")
CSRC = os.path.join(_ROOT, "secphase_b200", "csrc")


class SpParams(C.Structure):
    _fields_ = [
        ("baq_flag", C.c_int32), ("consensus", C.c_int32), ("indel_threshold", C.c_int32),
        ("min_q", C.c_int32), ("min_score", C.c_int32), ("set_q", C.c_int32), ("flank_margin", C.c_int32),
        ("prim_margin_score", C.c_double), ("prim_margin_random", C.c_double),
        ("conf_d", C.c_double), ("conf_e", C.c_double), ("conf_b", C.c_double),
    ]


def build(force=False):
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    newest = max(os.path.getmtime(p) for p in deps)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
                               "-o", LIB, SRC])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.hs_out_create.restype = C.c_void_p
        L.hs_out_destroy.argtypes = [C.c_void_p]
        L.hs_run.argtypes = [C.POINTER(CFlatBatch), C.POINTER(SpParams), C.c_void_p, C.c_void_p, C.c_int,
                             C.c_int, C.c_uint, C.c_void_p]
        L.hs_run.restype = C.c_int
        L.hs_run2.argtypes = [C.POINTER(CFlatBatch), C.POINTER(SpParams), C.c_void_p, C.c_void_p, C.c_int,
                              C.c_int, C.c_uint, C.c_int, C.c_void_p]
        L.hs_run2.restype = C.c_int
        L.hs_run3.argtypes = [C.POINTER(CFlatBatch), C.POINTER(SpParams), C.c_void_p, C.c_void_p, C.c_int,
                              C.c_int, C.c_uint, C.c_int, C.c_int, C.c_void_p]
        L.hs_run3.restype = C.c_int
        L.hs_fast_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.hs_hmmf.argtypes = [C.POINTER(SpParams), C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                              C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.hs_hmmf.restype = C.c_int
        L.hs_plan_sig.argtypes = [C.POINTER(CFlatBatch), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.hs_plan_sig.restype = C.c_uint64
        L.hs_qual.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.hs_qual.restype = C.POINTER(C.c_uint8)
        for name, rt in [("group", C.c_int32), ("score", C.c_double), ("extent", C.c_int32),
                         ("blocks", C.c_int32), ("block_off", C.c_int64), ("items", C.c_int32), ("rows", C.c_int32),
                         ("rows_pmax", C.c_double)]:
            f = getattr(L, "hs_" + name)
            f.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
            f.restype = C.POINTER(rt)
        L.hs_markers.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
        L.hs_markers.restype = C.POINTER(C.c_int32)
        L.hs_marker_off.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
        L.hs_marker_off.restype = C.POINTER(C.c_int64)
        L.hs_cells.argtypes = [C.c_void_p]
        L.hs_cells.restype = C.c_int64
        L.hs_err.argtypes = [C.c_void_p]
        L.hs_err.restype = C.c_int32
        L.hs_hmm.argtypes = [C.POINTER(SpParams), C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                             C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hs_hmm.restype = C.c_int
        L.hs_hmm2.argtypes = [C.POINTER(SpParams), C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                              C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.hs_hmm2.restype = C.c_int
        L.hs_rng_create.argtypes = [C.c_uint]
        L.hs_rng_create.restype = C.c_void_p
        L.hs_rng_next.argtypes = [C.c_void_p]
        L.hs_rng_next.restype = C.c_int
        L.hs_rng_destroy.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _take(ptr, n, width, dt):
    if n == 0:
        return np.zeros((0, width) if width > 1 else (0,), dtype=dt)
    a = np.ctypeslib.as_array(ptr, shape=(int(n),)).astype(dt, copy=True)
    return a.reshape(-1, width) if width > 1 else a


def params_from_oracle(op):
    return SpParams(**{k: getattr(op, k) for k, _ in SpParams._fields_})


def run(batch, params, ref_codes, contig_off, safe_caps=False, seed=1, full_baq=False, hmm_mode=0, group_lanes=False):
    """hmm_mode: 0 strict kernels; 1 fast kernel + guard-band re-run (the product's default); 2 = 1 and every
    fast instance cross-checked against the strict kernel (r["fast"] statistics, err bit 0x200 on a miss).
    group_lanes: stages K2 + K3 in their lane-per-alignment form (sp_group_warp.cuh) on the emulated warp."""
    L = lib()
    out = L.hs_out_create()
    L.hs_set_group_lanes(1 if group_lanes else 0)
    try:
        cb = batch.as_c()
        ref_codes = np.ascontiguousarray(ref_codes, np.uint8)
        contig_off = np.ascontiguousarray(contig_off, np.int64)
        rc = L.hs_run3(C.byref(cb), C.byref(params), ref_codes.ctypes.data, contig_off.ctypes.data,
                       len(contig_off) - 1, 1 if safe_caps else 0, seed, 1 if full_baq else 0, hmm_mode, out)
        if rc != 0:
            raise RuntimeError(f"hs_run failed: {rc}")
        n = C.c_int64()
        r = {"kind": "hostsim"}
        r["groups"] = _take(L.hs_group(out, C.byref(n)), n.value, 10, np.int32)
        r["scores"] = _take(L.hs_score(out, C.byref(n)), n.value, 1, np.float64)
        r["extents"] = _take(L.hs_extent(out, C.byref(n)), n.value, 4, np.int32)
        r["blocks"] = _take(L.hs_blocks(out, C.byref(n)), n.value, 6, np.int32)
        r["block_off"] = _take(L.hs_block_off(out, C.byref(n)), n.value, 1, np.int64)
        r["items"] = _take(L.hs_items(out, C.byref(n)), n.value, 8, np.int32)
        r["rows"] = _take(L.hs_rows(out, C.byref(n)), n.value, 4, np.int32)
        r["rows_pmax"] = _take(L.hs_rows_pmax(out, C.byref(n)), n.value, 1, np.float64)
        for st, nm in enumerate(("markers_pre", "markers_baq", "markers_final")):
            r[nm] = _take(L.hs_markers(out, st, C.byref(n)), n.value, 6, np.int32)
            r[nm + "_off"] = _take(L.hs_marker_off(out, st, C.byref(n)), n.value, 1, np.int64)
        if full_baq:
            r["qual"] = _take(L.hs_qual(out, C.byref(n)), n.value, 1, np.uint8)
        st = (C.c_int64 * 5)()
        drift = (C.c_double * 2)()
        L.hs_fast_stats(out, st, drift)
        r["fast"] = dict(zip(("instances", "rerun", "near_threshold", "near_tie", "numeric"), [int(x) for x in st]))
        r["fast"]["max_rel_drift"] = drift[0]
        r["fast"]["max_abs_drift_ulp"] = drift[1]
        r["cells"] = int(L.hs_cells(out))
        r["err"] = int(L.hs_err(out))
        return r
    finally:
        L.hs_out_destroy(out)


def hmm2(params, ref, query, par_bw, rows_t, unrolled=True):
    """The shared-memory-band kernel body (sp_hmm2.cuh) on the host; None if the band is too wide for it.
    unrolled: use the fully unrolled row bodies where the band class has them (a full warp's path)."""
    L = lib()
    ref = np.ascontiguousarray(ref, np.uint8)
    query = np.ascontiguousarray(query, np.uint8)
    rows_t = np.ascontiguousarray(rows_t, np.int32)
    n = len(rows_t)
    state = np.zeros(n, np.int32)
    q = np.zeros(n, np.uint8)
    pmax = np.zeros(n, np.float64)
    rc = L.hs_hmm2(C.byref(params), ref.ctypes.data, len(ref), query.ctypes.data, len(query), par_bw,
                   rows_t.ctypes.data, n, state.ctypes.data, q.ctypes.data, pmax.ctypes.data, 1 if unrolled else 0)
    if rc != 0:
        return None
    return dict(state=state, q=q, pmax=pmax)


def hmmf(params, ref, query, par_bw, rows_t, extra_bw=0):
    """The fast-arithmetic kernel body (sp_hmmf.cuh) on the host; None if the band is too wide for it.
    extra_bw > 0: run in a virtual band that much wider than the instance's own (a lane of a mixed warp)."""
    L = lib()
    ref = np.ascontiguousarray(ref, np.uint8)
    query = np.ascontiguousarray(query, np.uint8)
    rows_t = np.ascontiguousarray(rows_t, np.int32)
    n = len(rows_t)
    state = np.zeros(n, np.int32)
    q = np.zeros(n, np.uint8)
    pmax = np.zeros(n, np.float64)
    fl = L.hs_hmmf(C.byref(params), ref.ctypes.data, len(ref), query.ctypes.data, len(query), par_bw,
                   rows_t.ctypes.data, n, state.ctypes.data, q.ctypes.data, pmax.ctypes.data, extra_bw)
    if fl < 0:
        return None
    return dict(state=state, q=q, pmax=pmax, flags=fl)


def hmm(params, ref, query, par_bw, rows_t, want_s=False):
    L = lib()
    ref = np.ascontiguousarray(ref, np.uint8)
    query = np.ascontiguousarray(query, np.uint8)
    rows_t = np.ascontiguousarray(rows_t, np.int32)
    n = len(rows_t)
    state = np.zeros(n, np.int32)
    q = np.zeros(n, np.uint8)
    pmax = np.zeros(n, np.float64)
    s = np.zeros(len(query) + 2, np.float64)
    rc = L.hs_hmm(C.byref(params), ref.ctypes.data, len(ref), query.ctypes.data, len(query), par_bw,
                  rows_t.ctypes.data, n, state.ctypes.data, q.ctypes.data, pmax.ctypes.data, s.ctypes.data)
    assert rc == 0
    out = dict(state=state, q=q, pmax=pmax)
    if want_s:
        out["s"] = s
    return out


def plan_signature(batch, indel_threshold, n_threads, safe_caps=False):
    """(rc, hash of every offset table of the batch plan) -- sp_make_plan with n_threads scan threads."""
    cb = batch.as_c()
    rc = C.c_int()
    h = lib().hs_plan_sig(C.byref(cb), indel_threshold, 1 if safe_caps else 0, n_threads, C.byref(rc))
    return rc.value, int(h)


def walk_warp_check(batch, params):
    """Warp-cooperative walker (32 emulated lanes) vs the serial walker over every alignment of the batch:
    dict(alignments, handled, different, first, what)."""
    L = lib()
    L.hs_walk_warp_check.argtypes = [C.POINTER(CFlatBatch), C.POINTER(SpParams), C.POINTER(C.c_int64)]
    L.hs_walk_warp_check.restype = C.c_int
    cb = batch.as_c()
    out = (C.c_int64 * 5)()
    rc = L.hs_walk_warp_check(C.byref(cb), C.byref(params), out)
    if rc != 0:
        raise RuntimeError(f"hs_walk_warp_check rc={rc}")
    return dict(alignments=out[0], handled=out[1], different=out[2], first=out[3], what=out[4])
