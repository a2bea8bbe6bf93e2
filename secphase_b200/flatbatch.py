"""Python view of include/sp_flat_batch.h (ctypes struct + numpy-backed container).

Shared by tests, bench.py, the oracle wrapper and the product's Python host layer: it only
describes data (the BAM record fields of a set of read groups)."""
import ctypes as C

import numpy as np

_I32P = C.POINTER(C.c_int32)
_I64P = C.POINTER(C.c_int64)
_U32P = C.POINTER(C.c_uint32)
_U8P = C.POINTER(C.c_uint8)


class CFlatBatch(C.Structure):
    _fields_ = [
        ("n_groups", C.c_int32),
        ("n_alns", C.c_int32),
        ("grp_aln_off", _I32P),
        ("qname_off", _I64P),
        ("qname_pool", _U8P),
        ("flag", _I32P),
        ("tid", _I32P),
        ("pos", _I32P),
        ("l_qseq", _I32P),
        ("n_cigar", _I32P),
        ("tag_kind", _I32P),
        ("cigar_off", _I64P),
        ("tag_off", _I64P),
        ("seq_off", _I64P),
        ("qual_off", _I64P),
        ("cigar_pool", _U32P),
        ("tag_pool", _U8P),
        ("seq_pool", _U8P),
        ("qual_pool", _U8P),
    ]


_FIELDS = [
    ("grp_aln_off", np.int32), ("qname_off", np.int64), ("qname_pool", np.uint8),
    ("flag", np.int32), ("tid", np.int32), ("pos", np.int32), ("l_qseq", np.int32),
    ("n_cigar", np.int32), ("tag_kind", np.int32),
    ("cigar_off", np.int64), ("tag_off", np.int64), ("seq_off", np.int64), ("qual_off", np.int64),
    ("cigar_pool", np.uint32), ("tag_pool", np.uint8), ("seq_pool", np.uint8), ("qual_pool", np.uint8),
]


class FlatBatch:
    """Owns numpy copies of every array of an sp_flat_batch."""

    def __init__(self, **arrays):
        for name, dt in _FIELDS:
            a = np.ascontiguousarray(arrays[name], dtype=dt)
            if a.size == 0:  # keep a valid pointer for ctypes
                a = np.zeros(1, dtype=dt)[:0].copy()
            setattr(self, name, a)
        self.n_groups = len(self.grp_aln_off) - 1
        self.n_alns = len(self.flag)

    @classmethod
    def from_c(cls, view):
        v = view
        ng, na = v.n_groups, v.n_alns

        def arr(ptr, n, dt):
            if n == 0:
                return np.zeros(0, dtype=dt)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(n,)).copy()

        grp = arr(v.grp_aln_off, ng + 1, np.int32)
        qoff = arr(v.qname_off, ng + 1, np.int64)
        cig_off = arr(v.cigar_off, na + 1, np.int64)
        tag_off = arr(v.tag_off, na + 1, np.int64)
        seq_off = arr(v.seq_off, na + 1, np.int64)
        qual_off = arr(v.qual_off, na + 1, np.int64)
        return cls(
            grp_aln_off=grp, qname_off=qoff,
            qname_pool=arr(v.qname_pool, int(qoff[-1]), np.uint8),
            flag=arr(v.flag, na, np.int32), tid=arr(v.tid, na, np.int32), pos=arr(v.pos, na, np.int32),
            l_qseq=arr(v.l_qseq, na, np.int32), n_cigar=arr(v.n_cigar, na, np.int32),
            tag_kind=arr(v.tag_kind, na, np.int32) if v.tag_kind else np.zeros(na, np.int32),
            cigar_off=cig_off, tag_off=tag_off, seq_off=seq_off, qual_off=qual_off,
            cigar_pool=arr(v.cigar_pool, int(cig_off[-1]), np.uint32),
            tag_pool=arr(v.tag_pool, int(tag_off[-1]), np.uint8),
            seq_pool=arr(v.seq_pool, int(seq_off[-1]), np.uint8),
            qual_pool=arr(v.qual_pool, int(qual_off[-1]), np.uint8),
        )

    def as_c(self):
        s = CFlatBatch()
        s.n_groups = self.n_groups
        s.n_alns = self.n_alns
        for name, dt in _FIELDS:
            a = getattr(self, name)
            ftype = dict(CFlatBatch._fields_)[name]
            setattr(s, name, C.cast(a.ctypes.data, ftype))
        s._keep = self
        return s

    # ---- convenience for tests -------------------------------------------------
    def group_slice(self, g0, g1):
        """Sub-batch of groups [g0,g1) with re-based offsets."""
        a0, a1 = int(self.grp_aln_off[g0]), int(self.grp_aln_off[g1])

        def cut(off, pool):
            lo, hi = int(off[a0]), int(off[a1])
            return off[a0:a1 + 1] - lo, pool[lo:hi]

        cig_off, cig = cut(self.cigar_off, self.cigar_pool)
        tag_off, tag = cut(self.tag_off, self.tag_pool)
        seq_off, seq = cut(self.seq_off, self.seq_pool)
        qual_off, qual = cut(self.qual_off, self.qual_pool)
        q0, q1 = int(self.qname_off[g0]), int(self.qname_off[g1])
        return FlatBatch(
            grp_aln_off=self.grp_aln_off[g0:g1 + 1] - a0, qname_off=self.qname_off[g0:g1 + 1] - q0,
            qname_pool=self.qname_pool[q0:q1], flag=self.flag[a0:a1], tid=self.tid[a0:a1], pos=self.pos[a0:a1],
            l_qseq=self.l_qseq[a0:a1], n_cigar=self.n_cigar[a0:a1], tag_kind=self.tag_kind[a0:a1],
            cigar_off=cig_off, tag_off=tag_off, seq_off=seq_off, qual_off=qual_off,
            cigar_pool=cig, tag_pool=tag, seq_pool=seq, qual_pool=qual)

    def qname(self, g):
        return bytes(self.qname_pool[int(self.qname_off[g]):int(self.qname_off[g + 1])]).decode()

    def cigar_string(self, a):
        ops = "MIDNSHP=X"
        c = self.cigar_pool[int(self.cigar_off[a]):int(self.cigar_off[a]) + int(self.n_cigar[a])]
        return "".join(f"{int(x) >> 4}{ops[int(x) & 15]}" for x in c)

    def tag_string(self, a):
        return bytes(self.tag_pool[int(self.tag_off[a]):int(self.tag_off[a + 1])]).decode()

    @staticmethod
    def assemble(src, groups):
        """New batch from alignments of `src`: groups = [(qname, [(aln index in src, flag or None), ...]), ...].
        An alignment may be used several times (tests craft ineligible / oversized groups this way)."""
        grp, qoff, qpool = [0], [0], bytearray()
        per = {k: [] for k in ("flag", "tid", "pos", "l_qseq", "n_cigar", "tag_kind")}
        offs = {k: [0] for k in ("cigar", "tag", "seq", "qual")}
        pools = {k: [] for k in ("cigar", "tag", "seq", "qual")}
        for qname, alns in groups:
            qpool += qname.encode()
            qoff.append(len(qpool))
            for a, fl in alns:
                for k in per:
                    per[k].append(int(getattr(src, k)[a]))
                if fl is not None:
                    per["flag"][-1] = int(fl)
                for k in offs:
                    off = getattr(src, k + "_off")
                    piece = getattr(src, k + "_pool")[int(off[a]):int(off[a + 1])]
                    pools[k].append(piece)
                    offs[k].append(offs[k][-1] + len(piece))
            grp.append(len(per["flag"]))
        kw = {k: np.array(v, np.int32) for k, v in per.items()}
        kw["grp_aln_off"] = np.array(grp, np.int32)
        kw["qname_off"] = np.array(qoff, np.int64)
        kw["qname_pool"] = np.frombuffer(bytes(qpool), np.uint8)
        for k in offs:
            kw[k + "_off"] = np.array(offs[k], np.int64)
            kw[k + "_pool"] = np.concatenate(pools[k]) if pools[k] else np.zeros(0)
        return FlatBatch(**kw)

    @staticmethod
    def concat(batches):
        if len(batches) == 1:
            return batches[0]
        out = {}
        na = 0
        grp = [np.zeros(1, np.int32)]
        for name in ("qname", "cigar", "tag", "seq", "qual"):
            out[name + "_off"] = [np.zeros(1, np.int64)]
            out[name + "_pool"] = []
        base = dict(qname=0, cigar=0, tag=0, seq=0, qual=0)
        per_aln = {k: [] for k in ("flag", "tid", "pos", "l_qseq", "n_cigar", "tag_kind")}
        for b in batches:
            grp.append(b.grp_aln_off[1:] + na)
            na += b.n_alns
            for name in base:
                off = getattr(b, name + "_off")
                out[name + "_off"].append(off[1:] + base[name])
                pool = getattr(b, name + "_pool")
                out[name + "_pool"].append(pool)
                base[name] += int(off[-1])
            for k in per_aln:
                per_aln[k].append(getattr(b, k))
        kw = {k: np.concatenate(v) for k, v in per_aln.items()}
        kw["grp_aln_off"] = np.concatenate(grp)
        for name in base:
            kw[name + "_off"] = np.concatenate(out[name + "_off"])
            kw[name + "_pool"] = np.concatenate(out[name + "_pool"]) if out[name + "_pool"] else np.zeros(0)
        return FlatBatch(**kw)
