"""ctypes binding of the CPU checkers (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (secphase_b200/) never does.

kind="reference": oracle/_ref/libsecphase_ref.so -- the reference's OWN marker-path sources (cigar_it.c,
                  ptAlignment.c, ptMarker.c, ptBlock.c, common.c, compiled unmodified against oracle/shim)
                  + the restated probaln_glocal (oracle/probaln_port.c).  Built where /root/reference is
                  mounted; the .so travels with the repo snapshot to the GPU box.
kind="port":      oracle/liboracle_port.so -- reserved for a plain-C restatement of the whole marker path
                  exporting the same entry points (oracle/secphase_port.c, not written: the marker path is
                  checked against the reference's own code instead, and against the golden fixtures that
                  code wrote, tests/golden/, where the library is absent).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from secphase_b200.flatbatch import CFlatBatch

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libsecphase_ref.so")
PORT_LIB = os.path.join(_HERE, "liboracle_port.so")
HMM_LIB = os.path.join(_HERE, "liboracle_hmm.so")

GROUP_W, MARKER_W, BLOCK_W, HMM_W = 10, 6, 6, 8


class OracleParams(C.Structure):
    _fields_ = [
        ("baq_flag", C.c_int32), ("consensus", C.c_int32), ("indel_threshold", C.c_int32),
        ("min_q", C.c_int32), ("min_score", C.c_int32), ("set_q", C.c_int32), ("flank_margin", C.c_int32),
        ("prim_margin_score", C.c_double), ("prim_margin_random", C.c_double),
        ("conf_d", C.c_double), ("conf_e", C.c_double), ("conf_b", C.c_double),
    ]


class OracleRefSeq(C.Structure):
    _fields_ = [("n_contigs", C.c_int32), ("names", C.POINTER(C.c_char_p)),
                ("seqs", C.POINTER(C.c_void_p)), ("lens", C.POINTER(C.c_int64))]


def build(reference_root="/root/reference"):
    """Compile the checkers (building the checker is not using it)."""
    targets = ["port"]
    if os.path.exists(os.path.join(reference_root, "programs", "submodules", "ptMarker", "ptMarker.c")):
        targets.append("_ref")
    subprocess.check_call(["make", "-s", "-C", _HERE, f"REFERENCE={reference_root}"] + targets)


def available_kinds():
    k = []
    if os.path.exists(REF_LIB):
        k.append("reference")
    if os.path.exists(PORT_LIB):
        k.append("port")
    return k


_libs = {}


def _load(kind):
    if kind in _libs:
        return _libs[kind]
    path = REF_LIB if kind == "reference" else PORT_LIB
    if not os.path.exists(path):
        raise FileNotFoundError(f"oracle library for kind={kind!r} not built: {path}")
    L = C.CDLL(path)
    L.oracle_out_create.argtypes = [C.c_int]
    L.oracle_out_create.restype = C.c_void_p
    L.oracle_out_destroy.argtypes = [C.c_void_p]
    L.oracle_run.argtypes = [C.POINTER(CFlatBatch), C.POINTER(OracleParams), C.POINTER(OracleRefSeq), C.c_void_p]
    L.oracle_run.restype = C.c_int
    L.oracle_srand.argtypes = [C.c_uint]
    L.oracle_kind.restype = C.c_char_p
    for name, rt in [("groups", C.c_int32), ("scores", C.c_double), ("extents", C.c_int32),
                     ("blocks", C.c_int32), ("block_off", C.c_int64), ("hmm", C.c_int32),
                     ("hmm_state", C.c_int32), ("hmm_q", C.c_uint8), ("qual", C.c_uint8)]:
        f = getattr(L, "oracle_out_" + name)
        f.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        f.restype = C.POINTER(rt)
    if kind == "reference":
        L.oracle_out_enable_outputs.argtypes = [C.c_void_p]
        L.oracle_out_save.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.oracle_out_save.restype = C.c_int
        L.oracle_merge_blocks.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64]
        L.oracle_merge_blocks.restype = C.c_int64
        L.oracle_select.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(OracleParams), C.c_void_p]
        L.oracle_select.restype = C.c_int
    L.oracle_out_markers.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
    L.oracle_out_markers.restype = C.POINTER(C.c_int32)
    L.oracle_out_marker_off.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
    L.oracle_out_marker_off.restype = C.POINTER(C.c_int64)
    _libs[kind] = L
    return L


def _take(ptr, n, width, dt):
    total = int(n) * width
    if total == 0:
        return np.zeros((0, width) if width > 1 else (0,), dtype=dt)
    a = np.ctypeslib.as_array(ptr, shape=(total,)).astype(dt, copy=True)
    return a.reshape(-1, width) if width > 1 else a


def make_refseq(names, ptrs, lens):
    """names: list[str]; ptrs: list[int address of ASCII contig]; lens: list[int]."""
    n = len(names)
    r = OracleRefSeq()
    r.n_contigs = n
    r._names = (C.c_char_p * n)(*[s.encode() for s in names])
    r._seqs = (C.c_void_p * n)(*ptrs)
    r._lens = (C.c_int64 * n)(*lens)
    r.names = C.cast(r._names, C.POINTER(C.c_char_p))
    r.seqs = C.cast(r._seqs, C.POINTER(C.c_void_p))
    r.lens = C.cast(r._lens, C.POINTER(C.c_int64))
    return r


def preset_params(preset="hifi", **over):
    """secphase.c:477-504 (--hifi / --ont) on top of the defaults at secphase.c:420-449."""
    p = dict(baq_flag=1, consensus=1, indel_threshold=10, min_q=10, min_score=-10, set_q=40, flank_margin=500,
             prim_margin_score=40.0, prim_margin_random=0.0, conf_d=1e-4, conf_e=0.1, conf_b=20.0)
    if preset == "ont":
        p.update(indel_threshold=20, conf_d=1e-3, set_q=20, prim_margin_score=20.0)
    elif preset != "hifi":
        raise KeyError(preset)
    p.update(over)
    return OracleParams(**p)


def merge_blocks(rows, mode):
    """The reference's own ptBlock_merge_blocks (mode 0) / ptBlock_merge_blocks_v2 (mode 1) on rows of
    (start, end, count); count < 0 = no count data.  Needs kind="reference" (oracle/_ref)."""
    L = _load("reference")
    rows = np.ascontiguousarray(rows, np.int32).reshape(-1, 3)
    cap = 4 * len(rows) + 8
    out = np.zeros((cap, 3), np.int32)
    k = L.oracle_merge_blocks(rows.ctypes.data, len(rows), mode, out.ctypes.data, cap)
    assert k <= cap
    return out[:k]


def select(grp_aln_off, flag, scores, params, seed=1):
    """The reference's own get_best_record_index over already-scored groups, in group order, with one
    rand() stream seeded like a fresh process (srand(1) == never seeded): the selection a single-worker
    run of the reference makes.  Needs kind="reference"."""
    L = _load("reference")
    gao = np.ascontiguousarray(grp_aln_off, np.int32)
    fl = np.ascontiguousarray(flag, np.int32)
    sc = np.ascontiguousarray(scores, np.float64)
    best = np.zeros(len(gao) - 1, np.int32)
    if seed is not None:
        L.oracle_srand(seed)
    rc = L.oracle_select(len(gao) - 1, gao.ctypes.data, fl.ctypes.data, sc.ctypes.data, C.byref(params), best.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle_select failed: {rc}")
    return best


def run(batch, params, refseq, kind=None, keep_hmm=False, seed=1, outputs=None):
    """Run every read group of `batch` (secphase_b200.flatbatch.FlatBatch; or a list of them, processed in
    order with one rand() stream) through the CPU checker.  outputs=(dir, prefix) additionally
    writes out.log and the two marker BED files the way secphase.c does (reference kind only)."""
    if kind is None:
        kind = available_kinds()[0]
    L = _load(kind)
    out = L.oracle_out_create(1 if keep_hmm else 0)
    try:
        if seed is not None:
            L.oracle_srand(seed)
        if outputs is not None:
            if kind != "reference":
                raise ValueError("text outputs need the reference kind")
            L.oracle_out_enable_outputs(out)
        for b1 in (batch if isinstance(batch, (list, tuple)) else [batch]):
            cb = b1.as_c()
            rc = L.oracle_run(C.byref(cb), C.byref(params), C.byref(refseq), out)
            if rc != 0:
                raise RuntimeError(f"oracle_run failed: {rc}")
        n = C.c_int64()
        res = {"kind": kind}
        if outputs is not None:
            tot = (C.c_int32 * 4)()
            nm = C.c_int32()
            rc = L.oracle_out_save(out, os.fsencode(outputs[0]), outputs[1].encode(), tot, C.byref(nm))
            if rc != 0:
                raise RuntimeError(f"oracle_out_save failed: {rc}")
            res["totals"] = list(tot)
            res["reads_modified_by_marker"] = nm.value
        res["groups"] = _take(L.oracle_out_groups(out, C.byref(n)), n.value, GROUP_W, np.int32)
        res["scores"] = _take(L.oracle_out_scores(out, C.byref(n)), n.value, 1, np.float64)
        res["extents"] = _take(L.oracle_out_extents(out, C.byref(n)), n.value, 4, np.int32)
        res["blocks"] = _take(L.oracle_out_blocks(out, C.byref(n)), n.value, BLOCK_W, np.int32)
        res["block_off"] = _take(L.oracle_out_block_off(out, C.byref(n)), n.value, 1, np.int64)
        res["hmm"] = _take(L.oracle_out_hmm(out, C.byref(n)), n.value, HMM_W, np.int32)
        res["hmm_state"] = _take(L.oracle_out_hmm_state(out, C.byref(n)), n.value, 1, np.int32)
        res["hmm_q"] = _take(L.oracle_out_hmm_q(out, C.byref(n)), n.value, 1, np.uint8)
        res["qual"] = _take(L.oracle_out_qual(out, C.byref(n)), n.value, 1, np.uint8)
        for st, nm in enumerate(("markers_pre", "markers_baq", "markers_final")):
            res[nm] = _take(L.oracle_out_markers(out, st, C.byref(n)), n.value, MARKER_W, np.int32)
            res[nm + "_off"] = _take(L.oracle_out_marker_off(out, st, C.byref(n)), n.value, 1, np.int64)
        return res
    finally:
        L.oracle_out_destroy(out)


# ---- HMM-only checker -------------------------------------------------------------
_hmm = None


def hmm_lib():
    global _hmm
    if _hmm is None:
        if not os.path.exists(HMM_LIB):
            build()
        _hmm = C.CDLL(HMM_LIB)
        _hmm.oracle_probaln_glocal_ex.argtypes = [
            C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_int,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _hmm.oracle_probaln_glocal_ex.restype = C.c_int
        _hmm.oracle_probaln_cells.argtypes = [C.c_int, C.c_int, C.c_int]
        _hmm.oracle_probaln_cells.restype = C.c_long
    return _hmm


def probaln(ref, query, iqual, d, e, bw, want_s=False):
    """ref/query: uint8 codes 0..4; iqual: uint8 per query base.  Returns dict(state,q,pmax,pb[,s],Pr)."""
    ref = np.ascontiguousarray(ref, np.uint8)
    query = np.ascontiguousarray(query, np.uint8)
    iqual = np.ascontiguousarray(iqual, np.uint8)
    lq = len(query)
    state = np.zeros(lq, np.int32)
    q = np.zeros(lq, np.uint8)
    pmax = np.zeros(lq, np.float64)
    s = np.zeros(lq + 2, np.float64)
    pb = C.c_double()
    pr = hmm_lib().oracle_probaln_glocal_ex(ref.ctypes.data, len(ref), query.ctypes.data, lq, iqual.ctypes.data,
                                            d, e, bw, state.ctypes.data, q.ctypes.data, s.ctypes.data,
                                            pmax.ctypes.data, C.byref(pb))
    out = dict(state=state, q=q, pmax=pmax, pb=pb.value, Pr=pr)
    if want_s:
        out["s"] = s
    return out


def probaln_cells(l_ref, l_query, bw):
    return int(hmm_lib().oracle_probaln_cells(l_ref, l_query, bw))
