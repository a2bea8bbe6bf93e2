"""ctypes binding of tools/synth/synth.c (synthetic assembly + read groups; test/bench data)."""
import ctypes as C
import os
import subprocess

import numpy as np

from secphase_b200.flatbatch import CFlatBatch, FlatBatch

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_LIB_PATH = os.path.join(_ROOT, "secphase_b200", "lib", "libsp_synth.so")
_SRC = os.path.join(_ROOT, "tools", "synth", "synth.c")


class SynthCfg(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("n_loci", C.c_int32), ("n_copies", C.c_int32), ("locus_len", C.c_int64),
        ("snv_rate", C.c_double), ("indel_rate", C.c_double), ("long_indel_rate", C.c_double),
        ("hp_frac", C.c_double), ("n_rate", C.c_double), ("rc_contig_prob", C.c_double),
        ("len_mean", C.c_double), ("len_sd", C.c_double), ("len_min", C.c_int32), ("len_max", C.c_int32),
        ("err_sub", C.c_double), ("err_ins", C.c_double), ("err_del", C.c_double),
        ("hp_indel_mult", C.c_double), ("long_err_indel_prob", C.c_double), ("qual_model", C.c_int32),
        ("max_secondaries", C.c_int32), ("wrong_primary_prob", C.c_double), ("clip_prob", C.c_double),
        ("clip_max", C.c_int32), ("hard_clip_prob", C.c_double), ("eqx", C.c_int32), ("use_md", C.c_int32),
    ]


def build_lib(force=False):
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _LIB_PATH, _SRC, "-lm"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_lib())
        _lib.synth_default_cfg.argtypes = [C.POINTER(SynthCfg), C.c_int]
        _lib.synth_create.argtypes = [C.POINTER(SynthCfg)]
        _lib.synth_create.restype = C.c_void_p
        _lib.synth_destroy.argtypes = [C.c_void_p]
        _lib.synth_n_contigs.argtypes = [C.c_void_p]
        _lib.synth_n_contigs.restype = C.c_int32
        _lib.synth_contig_name.argtypes = [C.c_void_p, C.c_int32]
        _lib.synth_contig_name.restype = C.c_char_p
        _lib.synth_contig_seq.argtypes = [C.c_void_p, C.c_int32]
        _lib.synth_contig_seq.restype = C.c_void_p
        _lib.synth_contig_len.argtypes = [C.c_void_p, C.c_int32]
        _lib.synth_contig_len.restype = C.c_int64
        _lib.synth_generate.argtypes = [C.c_void_p, C.c_int64, C.c_int32]
        _lib.synth_generate.restype = C.c_void_p
        _lib.synth_batch_view.argtypes = [C.c_void_p]
        _lib.synth_batch_view.restype = C.POINTER(CFlatBatch)
        _lib.synth_batch_free.argtypes = [C.c_void_p]
    return _lib


PRESETS = {"hifi": 0, "ont": 1, "stress": 2}


def default_cfg(preset="hifi", **overrides):
    cfg = SynthCfg()
    lib().synth_default_cfg(C.byref(cfg), PRESETS[preset])
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise KeyError(k)
        setattr(cfg, k, v)
    return cfg


class Synth:
    """A synthetic assembly; generate(first, n) returns a FlatBatch of n read groups."""

    def __init__(self, cfg):
        self.cfg = cfg
        self._h = lib().synth_create(C.byref(cfg))
        if not self._h:
            raise ValueError("bad synth config")
        self.n_contigs = lib().synth_n_contigs(self._h)
        self.names = [lib().synth_contig_name(self._h, i).decode() for i in range(self.n_contigs)]
        self.lens = [int(lib().synth_contig_len(self._h, i)) for i in range(self.n_contigs)]

    def contig_ascii(self, tid):
        """numpy uint8 view (no copy) of the ASCII contig; valid while self lives."""
        p = lib().synth_contig_seq(self._h, tid)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(self.lens[tid],))

    def contig_ptr(self, tid):
        return lib().synth_contig_seq(self._h, tid)

    def generate(self, first, n):
        h = lib().synth_generate(self._h, first, n)
        try:
            return FlatBatch.from_c(lib().synth_batch_view(h).contents)
        finally:
            lib().synth_batch_free(h)

    def close(self):
        if self._h:
            lib().synth_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
