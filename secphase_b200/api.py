"""ctypes host layer over libsecphase_b200.so.

Mirrors the reference's per-job interface: `SpParams` carries the scalars of work_arg_t
(programs/submodules/tpool/tpool.h:26-55) with the same names, `params_for("hifi"|"ont")`
applies the presets of programs/src/secphase.c:477-504, and `Secphase.run(batch)` is the batch
form of runOneThread's marker branch (secphase.c:156-219): it returns, per read group, what that
job leaves behind -- alignment scores, selected index, alignment extents, final markers.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SECPHASE_B200_LIB") or os.path.join(_HERE, "lib", "libsecphase_b200.so")

GROUP_W, MARKER_W, BLOCK_W, HMM_W = 10, 6, 6, 8
N_SLOTS = 3


class SecphaseError(RuntimeError):
    pass


class SpParams(C.Structure):
    _fields_ = [
        ("baq_flag", C.c_int32), ("consensus", C.c_int32), ("indel_threshold", C.c_int32),
        ("min_q", C.c_int32), ("min_score", C.c_int32), ("set_q", C.c_int32), ("flank_margin", C.c_int32),
        ("prim_margin_score", C.c_double), ("prim_margin_random", C.c_double),
        ("conf_d", C.c_double), ("conf_e", C.c_double), ("conf_b", C.c_double),
    ]


class _CFlatBatch(C.Structure):
    _I32P, _I64P = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    _fields_ = [
        ("n_groups", C.c_int32), ("n_alns", C.c_int32),
        ("grp_aln_off", _I32P), ("qname_off", _I64P), ("qname_pool", C.POINTER(C.c_uint8)),
        ("flag", _I32P), ("tid", _I32P), ("pos", _I32P), ("l_qseq", _I32P), ("n_cigar", _I32P), ("tag_kind", _I32P),
        ("cigar_off", _I64P), ("tag_off", _I64P), ("seq_off", _I64P), ("qual_off", _I64P),
        ("cigar_pool", C.POINTER(C.c_uint32)), ("tag_pool", C.POINTER(C.c_uint8)),
        ("seq_pool", C.POINTER(C.c_uint8)), ("qual_pool", C.POINTER(C.c_uint8)),
    ]


class _CResult(C.Structure):
    _fields_ = [
        ("n_groups", C.c_int32), ("n_alns", C.c_int32),
        ("group", C.POINTER(C.c_int32)), ("score", C.POINTER(C.c_double)), ("extent", C.POINTER(C.c_int32)),
        ("marker_off", C.POINTER(C.c_int64)), ("marker", C.POINTER(C.c_int32)),
        ("hmm_instances", C.c_int64), ("hmm_cells", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("ms_total", C.c_float), ("ms_hmm", C.c_float), ("ms_stage", C.c_float * 8), ("gpu_launches", C.c_int32),
        ("baq_qual", C.POINTER(C.c_uint8)), ("baq_qual_bytes", C.c_int64),
        ("hmm_mode", C.c_int32), ("pad0", C.c_int32), ("hmm_strict_reruns", C.c_int64),
    ]


_lib = None


def load_library(path=None):
    """Load libsecphase_b200.so; raises SecphaseError (never falls back) when it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise SecphaseError(
            f"{path} not found: build it with `python -m secphase_b200.build` (nvcc, sm_100a). "
            "secphase_b200 has no CPU fallback.")
    L = C.CDLL(path)
    L.sp_last_error.restype = C.c_char_p
    L.sp_version.restype = C.c_char_p
    L.sp_params_default.argtypes = [C.POINTER(SpParams)]
    L.sp_params_preset.argtypes = [C.POINTER(SpParams), C.c_char_p]
    L.sp_params_preset.restype = C.c_int
    L.sp_create.argtypes = [C.POINTER(SpParams), C.c_int]
    L.sp_create.restype = C.c_void_p
    L.sp_destroy.argtypes = [C.c_void_p]
    L.sp_set_reference_ascii.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.sp_set_reference_ascii.restype = C.c_int
    L.sp_set_reference_codes.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    L.sp_set_reference_codes.restype = C.c_int
    L.sp_submit.argtypes = [C.c_void_p, C.POINTER(_CFlatBatch), C.c_int]
    L.sp_submit.restype = C.c_int
    L.sp_wait.argtypes = [C.c_void_p, C.c_int, C.POINTER(_CResult)]
    L.sp_wait.restype = C.c_int
    L.sp_upload.argtypes = [C.c_void_p, C.POINTER(_CFlatBatch), C.c_int]
    L.sp_upload.restype = C.c_int
    L.sp_run_resident.argtypes = [C.c_void_p, C.c_int]
    L.sp_run_resident.restype = C.c_int
    L.sp_debug_table.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_int32)),
                                 C.POINTER(C.POINTER(C.c_int64))]
    L.sp_debug_table.restype = C.c_int64
    L.sp_hmm_batch.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 12 + [C.POINTER(C.c_float)]
    L.sp_hmm_batch.restype = C.c_int
    L.sp_fp64_peak.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_float)]
    L.sp_fp64_peak.restype = C.c_int
    L.sp_host_alloc.argtypes = [C.c_size_t]
    L.sp_host_alloc.restype = C.c_void_p
    L.sp_host_free.argtypes = [C.c_void_p]
    L.sp_sm_partition.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.sp_sm_partition.restype = C.c_int
    L.sp_set_write_qual.argtypes = [C.c_void_p, C.c_int]
    L.sp_set_write_qual.restype = C.c_int
    L.sp_set_hmm_mode.argtypes = [C.c_void_p, C.c_int]
    L.sp_set_hmm_mode.restype = C.c_int
    L.sp_get_hmm_mode.argtypes = [C.c_void_p]
    L.sp_get_hmm_mode.restype = C.c_int
    L.sp_poll.argtypes = [C.c_void_p, C.c_int]
    L.sp_poll.restype = C.c_int
    L.sp_mark.argtypes = [C.c_void_p]
    L.sp_mark.restype = C.c_int
    L.sp_elapsed_since_mark.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
    L.sp_elapsed_since_mark.restype = C.c_int
    L.sp_rng_seed.argtypes = [C.c_void_p, C.c_uint]
    L.sp_rng_next.argtypes = [C.c_void_p]
    L.sp_rng_next.restype = C.c_int
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "sp_params_default", "sp_params_preset", "sp_create", "sp_destroy", "sp_last_error", "sp_version",
    "sp_set_reference_ascii", "sp_set_reference_codes", "sp_submit", "sp_wait", "sp_poll", "sp_sm_partition", "sp_mark", "sp_elapsed_since_mark", "sp_upload", "sp_run_resident",
    "sp_set_write_qual", "sp_set_hmm_mode", "sp_get_hmm_mode", "sp_debug_table", "sp_hmm_batch", "sp_fp64_peak", "sp_rng_seed", "sp_rng_next", "sp_host_alloc", "sp_host_free",
]


class PinnedArray:
    """numpy view over page-locked memory from sp_host_alloc (freed when the object dies)."""

    def __init__(self, like):
        L = load_library()
        like = np.ascontiguousarray(like)
        self._L = L
        self.nbytes = max(int(like.nbytes), 1)
        self.ptr = L.sp_host_alloc(self.nbytes)
        if not self.ptr:
            raise SecphaseError(L.sp_last_error().decode())
        buf = (C.c_uint8 * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=like.dtype, count=like.size)
        self.array[...] = like.ravel()

    def __del__(self):
        try:
            if self.ptr:
                self._L.sp_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def pin_batch(batch):
    """Copy of a FlatBatch whose four pools live in page-locked memory (what a reader thread that
    decodes straight into sp_host_alloc'ed buffers would hand to sp_submit)."""
    import copy
    out = copy.copy(batch)
    out._pins = []
    for name in ("cigar_pool", "tag_pool", "seq_pool", "qual_pool"):
        pa = PinnedArray(getattr(batch, name))
        out._pins.append(pa)
        setattr(out, name, pa.array)
    return out


def params_for(preset=None, **overrides):
    """secphase.c:420-449 defaults, then the --hifi/--ont preset (477-504), then explicit overrides
    (the reference's option parsing is order-sensitive in the same way: later flags win)."""
    L = load_library()
    p = SpParams()
    L.sp_params_default(C.byref(p))
    if preset is not None:
        if L.sp_params_preset(C.byref(p), preset.encode()) != 0:
            raise SecphaseError(L.sp_last_error().decode())
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def _c_batch(batch):
    """secphase_b200.flatbatch.FlatBatch (or anything with the same numpy attributes) -> _CFlatBatch."""
    s = _CFlatBatch()
    s.n_groups = batch.n_groups
    s.n_alns = batch.n_alns
    for name, ftype in _CFlatBatch._fields_[2:]:
        a = getattr(batch, name)
        setattr(s, name, C.cast(a.ctypes.data, ftype))
    s._keep = batch
    return s


def _take(ptr, n, width, dt):
    total = int(n) * width
    if total == 0:
        return np.zeros((0, width) if width > 1 else (0,), dtype=dt)
    a = np.ctypeslib.as_array(ptr, shape=(total,)).astype(dt, copy=True)
    return a.reshape(-1, width) if width > 1 else a


class Secphase:
    """One context per GPU (include/secphase_b200.h).  Not thread-safe; drive from one thread."""

    def __init__(self, preset="hifi", device=0, params=None, **overrides):
        self._L = load_library()
        self.params = params if params is not None else params_for(preset, **overrides)
        self._h = self._L.sp_create(C.byref(self.params), device)
        if not self._h:
            raise SecphaseError(self._L.sp_last_error().decode())
        self._keep = {}

    # -- plumbing ---------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise SecphaseError(f"libsecphase_b200 error {rc}: {self._L.sp_last_error().decode()}")

    def close(self):
        if getattr(self, "_h", None):
            self._L.sp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- reference ----------------------------------------------------------------------------
    def set_reference_codes(self, codes, contig_off):
        codes = np.ascontiguousarray(codes, np.uint8)
        contig_off = np.ascontiguousarray(contig_off, np.int64)
        self._ck(self._L.sp_set_reference_codes(self._h, len(contig_off) - 1, codes.ctypes.data, contig_off.ctypes.data))

    def set_reference_ascii(self, ptrs, lens):
        n = len(lens)
        arr = (C.c_void_p * n)(*ptrs)
        ln = (C.c_int64 * n)(*lens)
        self._ck(self._L.sp_set_reference_ascii(self._h, n, arr, ln))

    # -- batches ------------------------------------------------------------------------------
    def submit(self, batch, slot=0):
        cb = _c_batch(batch)
        self._keep[slot] = cb
        self._ck(self._L.sp_submit(self._h, C.byref(cb), slot))

    def upload(self, batch, slot=0):
        cb = _c_batch(batch)
        self._keep[slot] = cb
        self._ck(self._L.sp_upload(self._h, C.byref(cb), slot))

    def run_resident(self, slot=0):
        self._ck(self._L.sp_run_resident(self._h, slot))

    def wait(self, slot=0, copy=True):
        r = _CResult()
        self._ck(self._L.sp_wait(self._h, slot, C.byref(r)))
        out = {
            "n_groups": r.n_groups, "n_alns": r.n_alns,
            "hmm_instances": r.hmm_instances, "hmm_cells": r.hmm_cells,
            "h2d_bytes": r.h2d_bytes, "d2h_bytes": r.d2h_bytes,
            "ms_total": r.ms_total, "ms_hmm": r.ms_hmm, "ms_stage": list(r.ms_stage),
            "gpu_launches": r.gpu_launches,
            "hmm_mode": "fast" if r.hmm_mode == 1 else "strict", "hmm_rerun": int(r.hmm_strict_reruns),
        }
        if copy:
            out["groups"] = _take(r.group, r.n_groups, GROUP_W, np.int32)
            out["scores"] = _take(r.score, r.n_alns, 1, np.float64)
            out["extents"] = _take(r.extent, r.n_alns, 4, np.int32)
            out["markers_final_off"] = _take(r.marker_off, r.n_groups + 1, 1, np.int64)
            nm = int(out["markers_final_off"][-1]) if r.n_groups >= 0 else 0
            out["markers_final"] = _take(r.marker, nm, MARKER_W, np.int32)
            if r.baq_qual:
                out["baq_qual"] = _take(r.baq_qual, r.baq_qual_bytes, 1, np.uint8)
        return out

    def run(self, batch, slot=0):
        """Score every read group of `batch`; blocking convenience wrapper."""
        self.submit(batch, slot)
        return self.wait(slot)

    def set_write_qual(self, on=True):
        """-w/--writeBam mode (secphase.c:182-189): wait() then also returns `baq_qual`, the records'
        quality arrays after calc_update_baq_all, laid out like the batch's qual_pool."""
        self._ck(self._L.sp_set_write_qual(self._h, 1 if on else 0))

    def set_hmm_mode(self, mode):
        """"strict": the reference's rounding order; "fast" (default): FMA / scale-free arithmetic with a guard
        band and a strict re-run of the flagged instances (include/secphase_b200.h, sp_set_hmm_mode)."""
        self._ck(self._L.sp_set_hmm_mode(self._h, {"strict": 0, "fast": 1}[mode]))

    def hmm_mode(self):
        return "fast" if self._L.sp_get_hmm_mode(self._h) == 1 else "strict"

    def enable_debug_tables(self):
        self._L.sp_debug_table(self._h, 0, -1, None, None)

    def debug_force_block_cap(self, cap):
        """Tests: clamp the first plan's per-group block workspace to `cap` entries (0 = off), which makes the
        kernels flag SP_GERR_BLOCK_CAP and sp_wait re-run the batch with the provable bounds."""
        self._L.sp_debug_table(self._h, 0, -100 - int(cap), None, None)

    def cap_retries(self):
        return int(self._L.sp_debug_table(self._h, 0, -2, None, None))

    def walk_fallbacks(self, slot=0):
        """Alignments of the slot's last batch that the warp-cooperative walker handed to the serial one
        (-1 when SECPHASE_B200_WALK=serial selected the serial walker for everything)."""
        return int(self._L.sp_debug_table(self._h, slot, -3, None, None))

    def debug_table(self, what, slot=0):
        rows = C.POINTER(C.c_int32)()
        off = C.POINTER(C.c_int64)()
        n = self._L.sp_debug_table(self._h, slot, what, C.byref(rows), C.byref(off))
        if n < 0:
            raise SecphaseError(f"sp_debug_table({what}) -> {n}: {self._L.sp_last_error().decode()}")
        width = {0: MARKER_W, 1: MARKER_W, 2: BLOCK_W, 3: HMM_W, 4: 4}[what]
        table = _take(rows, n, width, np.int32)
        offs = None
        if what in (0, 1, 2) and off:
            cb = self._keep[slot]
            cnt = (cb.n_groups if what in (0, 1) else cb.n_alns) + 1
            offs = _take(off, cnt, 1, np.int64)
        return table, offs

    def run_debug(self, batch, slot=0):
        """run() plus every intermediate table, named like oracle.pyoracle.run()'s result."""
        self.enable_debug_tables()
        res = self.run(batch, slot)
        res["markers_pre"], res["markers_pre_off"] = self.debug_table(0, slot)
        res["markers_baq"], res["markers_baq_off"] = self.debug_table(1, slot)
        res["blocks"], res["block_off"] = self.debug_table(2, slot)
        res["items"], _ = self.debug_table(3, slot)
        res["rows"], _ = self.debug_table(4, slot)
        return res

    # -- the HMM alone ------------------------------------------------------------------------
    def hmm_batch(self, refs, queries, par_bw, rows):
        """refs/queries: lists of uint8 code arrays; par_bw: list of conf.bw; rows: list of ascending
        0-based row arrays.  Returns (state, q, pmax) lists and the kernel time in ms."""
        n = len(refs)
        l_ref = np.array([len(x) for x in refs], np.int32)
        l_q = np.array([len(x) for x in queries], np.int32)
        ref_off = np.zeros(n + 1, np.int64)
        ref_off[1:] = np.cumsum(l_ref)
        q_off = np.zeros(n + 1, np.int64)
        q_off[1:] = np.cumsum(l_q)
        row_off = np.zeros(n + 1, np.int64)
        row_off[1:] = np.cumsum([len(r) for r in rows])
        ref_pool = np.concatenate([np.asarray(x, np.uint8) for x in refs]) if n else np.zeros(1, np.uint8)
        q_pool = np.concatenate([np.asarray(x, np.uint8) for x in queries]) if n else np.zeros(1, np.uint8)
        rows_all = (np.concatenate([np.asarray(r, np.int32) for r in rows]) if n else np.zeros(0, np.int32))
        rows_all = np.ascontiguousarray(rows_all, np.int32)
        if len(rows_all) == 0:
            rows_all = np.zeros(1, np.int32)
        nr = int(row_off[-1])
        state = np.zeros(max(nr, 1), np.int32)
        q = np.zeros(max(nr, 1), np.uint8)
        pmax = np.zeros(max(nr, 1), np.float64)
        bw = np.ascontiguousarray(par_bw, np.int32)
        ms = C.c_float()
        self._ck(self._L.sp_hmm_batch(self._h, n, ref_pool.ctypes.data, ref_off.ctypes.data, l_ref.ctypes.data,
                                      q_pool.ctypes.data, q_off.ctypes.data, l_q.ctypes.data, bw.ctypes.data,
                                      row_off.ctypes.data, rows_all.ctypes.data, state.ctypes.data, q.ctypes.data,
                                      pmax.ctypes.data, C.byref(ms)))
        sp = lambda a: [a[int(row_off[j]):int(row_off[j + 1])] for j in range(n)]  # noqa: E731
        return sp(state), sp(q), sp(pmax), ms.value

    def poll(self, slot=0):
        """True once the slot's batch has finished on the device (wait() will not block on the GPU)."""
        r = self._L.sp_poll(self._h, slot)
        if r < 0:
            self._ck(r)
        return r == 1

    def sm_partition(self):
        """(partitioned, int_sms, hmm_sms): SMs set aside for the integer stages / left to the HMM kernels."""
        a, b = C.c_int32(), C.c_int32()
        r = self._L.sp_sm_partition(self._h, C.byref(a), C.byref(b))
        if r < 0:
            self._ck(r)
        return bool(r), a.value, b.value

    def mark(self):
        """Start of a device-side stopwatch (CUDA event on slot 0's stream); see elapsed_since_mark."""
        self._ck(self._L.sp_mark(self._h))

    def elapsed_since_mark(self, slot=0):
        """CUDA-event milliseconds from mark() to the end of the last batch waited for on `slot`."""
        ms = C.c_float()
        self._ck(self._L.sp_elapsed_since_mark(self._h, slot, C.byref(ms)))
        return ms.value

    def fp64_peak(self, mode=0):
        ops = C.c_double()
        ms = C.c_float()
        self._ck(self._L.sp_fp64_peak(self._h, mode, C.byref(ops), C.byref(ms)))
        return ops.value, ms.value

    def rng_seed(self, seed):
        self._L.sp_rng_seed(self._h, seed)

    def rng_next(self):
        return self._L.sp_rng_next(self._h)
