"""ctypes binding of libsecphase_host.so (include/secphase_host.h): BGZF/BAM ingest into flat
batches, FASTA load, BED block tables and out.log records -- the host side of the `secphase`
drop-in.  No CUDA in here; scoring is secphase_b200.api (libsecphase_b200.so)."""
import ctypes as C
import os

import numpy as np

from .flatbatch import CFlatBatch, FlatBatch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsecphase_host.so")
CLI_PATH = os.path.join(os.path.dirname(_HERE), "bin", "secphase")

_lib = None


class HostError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HostError(f"{LIB_PATH} is not built; run `python -m secphase_b200.build`")
    L = C.CDLL(LIB_PATH)
    L.sph_last_error.restype = C.c_char_p
    L.sph_bam_open.argtypes = [C.c_char_p, C.c_int]
    L.sph_bam_open.restype = C.c_void_p
    L.sph_bam_close.argtypes = [C.c_void_p]
    L.sph_bam_n_targets.argtypes = [C.c_void_p]
    L.sph_bam_n_targets.restype = C.c_int32
    L.sph_bam_target_name.argtypes = [C.c_void_p, C.c_int32]
    L.sph_bam_target_name.restype = C.c_char_p
    L.sph_bam_target_len.argtypes = [C.c_void_p, C.c_int32]
    L.sph_bam_target_len.restype = C.c_int64
    L.sph_bam_set_contig_limits.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    L.sph_batch_create.argtypes = [C.c_void_p, C.c_void_p]
    L.sph_batch_create.restype = C.c_void_p
    L.sph_batch_destroy.argtypes = [C.c_void_p]
    L.sph_batch_view.argtypes = [C.c_void_p]
    L.sph_batch_view.restype = C.POINTER(CFlatBatch)
    L.sph_batch_record_index.argtypes = [C.c_void_p]
    L.sph_batch_record_index.restype = C.POINTER(C.c_int64)
    L.sph_bam_next_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64]
    L.sph_bam_next_batch.restype = C.c_int32
    L.sph_bam_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.sph_bam_skipped_groups.argtypes = [C.c_void_p]
    L.sph_bam_skipped_groups.restype = C.c_int64
    L.sph_bam_write.argtypes = [C.c_char_p, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_int64),
                                C.POINTER(CFlatBatch), C.c_int, C.c_int]
    L.sph_bamw_open.argtypes = [C.c_char_p, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.c_int, C.c_int]
    L.sph_bamw_open.restype = C.c_void_p
    L.sph_bamw_add.argtypes = [C.c_void_p, C.POINTER(CFlatBatch)]
    L.sph_bamw_close.argtypes = [C.c_void_p]
    L.sph_fasta_load.argtypes = [C.c_char_p, C.c_int]
    L.sph_fasta_load.restype = C.c_void_p
    L.sph_fasta_free.argtypes = [C.c_void_p]
    L.sph_fasta_n.argtypes = [C.c_void_p]
    L.sph_fasta_n.restype = C.c_int32
    L.sph_fasta_name.argtypes = [C.c_void_p, C.c_int32]
    L.sph_fasta_name.restype = C.c_char_p
    L.sph_fasta_len.argtypes = [C.c_void_p, C.c_int32]
    L.sph_fasta_len.restype = C.c_int64
    L.sph_fasta_codes.argtypes = [C.c_void_p]
    L.sph_fasta_codes.restype = C.POINTER(C.c_uint8)
    L.sph_fasta_offsets.argtypes = [C.c_void_p]
    L.sph_fasta_offsets.restype = C.POINTER(C.c_int64)
    L.sph_fasta_write.argtypes = [C.c_char_p, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_void_p),
                                  C.POINTER(C.c_int64), C.c_int]
    L.sph_blocks_create.argtypes = [C.c_int]
    L.sph_blocks_create.restype = C.c_void_p
    L.sph_blocks_destroy.argtypes = [C.c_void_p]
    L.sph_blocks_add.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, C.c_int32]
    L.sph_blocks_add_count.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, C.c_int32, C.c_int32]
    L.sph_blocks_merge_v2.argtypes = [C.c_void_p]
    L.sph_blocks_merge.argtypes = [C.c_void_p]
    L.sph_blocks_total_length.argtypes = [C.c_void_p]
    L.sph_blocks_total_length.restype = C.c_int64
    L.sph_blocks_total_number.argtypes = [C.c_void_p]
    L.sph_blocks_total_number.restype = C.c_int64
    L.sph_blocks_save_bed.argtypes = [C.c_void_p, C.c_char_p]
    L.sph_blocks_export.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int64]
    L.sph_blocks_export.restype = C.c_int64
    L.sph_format_marker_record.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int32, C.c_int32,
                                           C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_char_p),
                                           C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32]
    L.sph_format_marker_record.restype = C.c_int64
    L.sph_batch_keep_records.argtypes = [C.c_void_p, C.c_int]
    L.sph_batch_records.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_int64))]
    L.sph_batch_records.restype = C.POINTER(C.c_uint8)
    L.sph_bam_header_text.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    L.sph_bam_header_text.restype = C.c_void_p
    L.sph_format_sam_record.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32,
                                        C.POINTER(C.c_char_p), C.c_void_p]
    L.sph_format_sam_record.restype = C.c_int64
    L.sph_samw_open.argtypes = [C.c_char_p, C.c_void_p]
    L.sph_samw_open.restype = C.c_void_p
    L.sph_samw_open_mt.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
    L.sph_samw_open_mt.restype = C.c_void_p
    L.sph_samw_write_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.sph_samw_close.argtypes = [C.c_void_p]
    _lib = L
    return L


def _err():
    return (lib().sph_last_error() or b"").decode(errors="replace")


def _names_lens(names, lens):
    n = len(names)
    cn = (C.c_char_p * n)(*[s.encode() for s in names])
    cl = (C.c_int64 * n)(*[int(x) for x in lens])
    return n, cn, cl


def write_bam(path, names, lens, batches, level=1, threads=4):
    """Writes one FlatBatch or a list of them as a queryname-grouped BAM (test/benchmark data)."""
    if isinstance(batches, FlatBatch):
        batches = [batches]
    n, cn, cl = _names_lens(names, lens)
    w = lib().sph_bamw_open(os.fsencode(path), n, cn, cl, level, threads)
    if not w:
        raise HostError(_err())
    try:
        for b in batches:
            cb = b.as_c()
            if lib().sph_bamw_add(w, C.byref(cb)) != 0:
                raise HostError(_err())
    finally:
        if lib().sph_bamw_close(w) != 0:
            raise HostError(_err())


def write_fasta(path, names, seq_ptrs, lens, line_width=60):
    """seq_ptrs: addresses (int) or bytes objects of the ASCII contigs.  Also writes <path>.fai."""
    n = len(names)
    keep = []
    ptrs = []
    for s in seq_ptrs:
        if isinstance(s, (bytes, bytearray)):
            buf = C.create_string_buffer(bytes(s), len(s))
            keep.append(buf)
            ptrs.append(C.cast(buf, C.c_void_p).value)
        else:
            ptrs.append(int(s))
    cn = (C.c_char_p * n)(*[x.encode() for x in names])
    cp = (C.c_void_p * n)(*ptrs)
    cl = (C.c_int64 * n)(*[int(x) for x in lens])
    if lib().sph_fasta_write(os.fsencode(path), n, cn, cp, cl, line_width) != 0:
        raise HostError(_err())


class BamReader:
    """Iterates the eligible read groups of a BAM as FlatBatch objects (numpy copies)."""

    def __init__(self, path, threads=4, keep_records=False):
        self._h = lib().sph_bam_open(os.fsencode(path), threads)
        if not self._h:
            raise HostError(_err())
        self._b = lib().sph_batch_create(None, None)
        self.keep_records = keep_records
        if keep_records:  # -w/--writeBam: whole record bodies travel with the batch
            lib().sph_batch_keep_records(self._b, 1)
        n = lib().sph_bam_n_targets(self._h)
        self.names = [lib().sph_bam_target_name(self._h, i).decode() for i in range(n)]
        self.lens = [int(lib().sph_bam_target_len(self._h, i)) for i in range(n)]

    def set_contig_limits(self, lens):
        arr = (C.c_int64 * len(lens))(*[int(x) for x in lens])
        if lib().sph_bam_set_contig_limits(self._h, arr) != 0:
            raise HostError(_err())

    def next_batch(self, max_groups=4096, max_bytes=256 << 20):
        n = lib().sph_bam_next_batch(self._h, self._b, max_groups, max_bytes)
        if n < 0:
            raise HostError(_err())
        if n == 0:
            return None
        fb = FlatBatch.from_c(lib().sph_batch_view(self._b).contents)
        ri = lib().sph_batch_record_index(self._b)
        fb.record_index = np.ctypeslib.as_array(ri, shape=(fb.n_alns,)).copy()
        if self.keep_records:
            off = C.POINTER(C.c_int64)()
            pool = lib().sph_batch_records(self._b, C.byref(off))
            fb.rec_off = np.ctypeslib.as_array(off, shape=(fb.n_alns + 1,)).copy()
            fb.rec_pool = np.ctypeslib.as_array(pool, shape=(max(int(fb.rec_off[-1]), 1),))[:int(fb.rec_off[-1])].copy()
        return fb

    def header_text(self):
        n = C.c_int64()
        p = lib().sph_bam_header_text(self._h, C.byref(n))
        return C.string_at(p, n.value).decode() if p and n.value else ""


    def __iter__(self):
        while True:
            b = self.next_batch()
            if b is None:
                return
            yield b

    def counts(self):
        a, r = C.c_int64(), C.c_int64()
        lib().sph_bam_counts(self._h, C.byref(a), C.byref(r))
        return a.value, r.value

    def skipped_groups(self):
        return int(lib().sph_bam_skipped_groups(self._h))

    def close(self):
        if self._h:
            lib().sph_batch_destroy(self._b)
            lib().sph_bam_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def load_fasta(path, threads=4):
    """Returns (names, codes uint8 array, offsets int64 array [n+1])."""
    h = lib().sph_fasta_load(os.fsencode(path), threads)
    if not h:
        raise HostError(_err())
    try:
        n = lib().sph_fasta_n(h)
        names = [lib().sph_fasta_name(h, i).decode() for i in range(n)]
        off = np.ctypeslib.as_array(lib().sph_fasta_offsets(h), shape=(n + 1,)).copy()
        total = int(off[-1])
        codes = np.ctypeslib.as_array(lib().sph_fasta_codes(h), shape=(max(total, 1),))[:total].copy()
        return names, codes, off
    finally:
        lib().sph_fasta_free(h)


class Blocks:
    """stHash contig -> block list of the reference, with its two merges and the BED writer."""

    def __init__(self, with_count):
        self._h = lib().sph_blocks_create(1 if with_count else 0)

    def add(self, contig, rfs, rfe, count=1):
        lib().sph_blocks_add_count(self._h, contig.encode(), rfs, rfe, count)

    def merge_v2(self):
        lib().sph_blocks_merge_v2(self._h)

    def merge(self):
        lib().sph_blocks_merge(self._h)

    def total_length(self):
        return int(lib().sph_blocks_total_length(self._h))

    def total_number(self):
        return int(lib().sph_blocks_total_number(self._h))

    def save_bed(self, path):
        if lib().sph_blocks_save_bed(self._h, os.fsencode(path)) != 0:
            raise HostError(_err())

    def rows(self):
        n = lib().sph_blocks_export(self._h, None, 0)
        out = np.zeros((max(n, 1), 4), np.int32)
        lib().sph_blocks_export(self._h, out.ctypes.data_as(C.POINTER(C.c_int32)), n)
        return out[:n]

    def __del__(self):
        try:
            lib().sph_blocks_destroy(self._h)
        except Exception:
            pass


def format_marker_record(qname, flags, scores, contigs, pos, rfe, best_idx):
    n = len(flags)
    cf = (C.c_int32 * n)(*flags)
    cs = (C.c_double * n)(*scores)
    cc = (C.c_char_p * n)(*[c.encode() for c in contigs])
    cp = (C.c_int32 * n)(*pos)
    ce = (C.c_int32 * n)(*rfe)
    q = qname.encode()
    need = lib().sph_format_marker_record(None, 0, q, len(q), n, cf, cs, cc, cp, ce, best_idx)
    buf = C.create_string_buffer(need + 1)
    lib().sph_format_marker_record(buf, need, q, len(q), n, cf, cs, cc, cp, ce, best_idx)
    return buf.raw[:need].decode()


def format_sam_record(rec, names, qual=None):
    """One SAM text line (htslib sam_format1) of a BAM record body; qual (uint8 Phred) overrides QUAL."""
    rec = np.ascontiguousarray(np.frombuffer(bytes(rec), np.uint8))
    n = len(names)
    cn = (C.c_char_p * max(n, 1))(*[x.encode() for x in names])
    q = None if qual is None else np.ascontiguousarray(qual, np.uint8)
    qp = None if q is None else q.ctypes.data
    need = lib().sph_format_sam_record(None, 0, rec.ctypes.data, len(rec), n, cn, qp)
    if need < 0:
        raise HostError(_err())
    buf = C.create_string_buffer(need + 1)
    lib().sph_format_sam_record(buf, need, rec.ctypes.data, len(rec), n, cn, qp)
    return buf.raw[:need].decode()


class SamWriter:
    """<prefix>.quality_modified.out.bam of secphase.c:643-657 (SAM text, see sph_sam.cpp)."""

    def __init__(self, path, reader, threads=1):
        self._w = lib().sph_samw_open_mt(os.fsencode(path), reader._h, threads)
        if not self._w:
            raise HostError(_err())

    def write_current_batch(self, reader, baq_qual=None):
        """Writes the records of the batch `reader.next_batch()` returned last (keep_records=True)."""
        q = None if baq_qual is None else np.ascontiguousarray(baq_qual, np.uint8)
        if lib().sph_samw_write_batch(self._w, reader._b, None if q is None else q.ctypes.data) != 0:
            raise HostError(_err())

    def close(self):
        if self._w:
            rc = lib().sph_samw_close(self._w)
            self._w = None
            if rc != 0:
                raise HostError("closing the SAM output failed")
