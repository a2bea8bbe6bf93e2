"""-w/--writeBam host side (libsecphase_host, sph_sam.cpp): whole BAM records kept with the batch
and written back as SAM text -- what the reference's sam_open(path, "w") + sam_write1 produce
(secphase.c:182-189, 643-657).  Checked against an independent Python formatter of the flat batch
fields and against hand-written expected lines for every aux type of the SAM spec."""
import struct

import numpy as np
import pytest

from secphase_b200 import hostlib
from tests.conftest import make_case

CIG = "MIDNSHP=X"
NT16 = "=ACMGRSVTWYHKDBN"


def py_sam_line(b, g, a, names, qual=None):
    qn = bytes(b.qname_pool[int(b.qname_off[g]):int(b.qname_off[g + 1])]).decode()
    cig = b.cigar_pool[int(b.cigar_off[a]):int(b.cigar_off[a + 1])]
    cigar = "".join(f"{int(c) >> 4}{CIG[int(c) & 15]}" for c in cig)
    lq = int(b.l_qseq[a])
    sq = b.seq_pool[int(b.seq_off[a]):int(b.seq_off[a]) + (lq + 1) // 2]
    nib = np.empty(2 * len(sq), np.uint8)
    nib[0::2] = sq >> 4
    nib[1::2] = sq & 15
    seq = "".join(NT16[int(x)] for x in nib[:lq])
    q = b.qual_pool[int(b.qual_off[a]):int(b.qual_off[a]) + lq] if qual is None else qual
    qs = bytes((np.asarray(q, np.uint8) + 33).tolist()).decode()
    tag = bytes(b.tag_pool[int(b.tag_off[a]):int(b.tag_off[a + 1])]).decode()
    kind = "cs" if (b.tag_kind is None or int(b.tag_kind[a]) == 0) else "MD"
    mapq = 0 if int(b.flag[a]) & 0x100 else 60  # what hostlib.write_bam stores
    return (f"{qn}\t{int(b.flag[a])}\t{names[int(b.tid[a])]}\t{int(b.pos[a]) + 1}\t{mapq}\t{cigar}\t*\t0\t0\t"
            f"{seq}\t{qs}\t{kind}:Z:{tag}\n")


@pytest.mark.parametrize("preset,over", [("hifi", {}), ("ont", dict(use_md=1)), ("stress", {})])
def test_records_kept_and_formatted(tmp_path, monkeypatch, preset, over):
    s, b, _, _ = make_case(preset, 30, locus_len=200000, **over)
    monkeypatch.setenv("SPH_CHUNK_BYTES", "90000")  # records straddle chunk switches
    monkeypatch.setenv("SPH_HEAD_ROOM", "512")
    p = str(tmp_path / "t.bam")
    hostlib.write_bam(p, s.names, s.lens, b, level=1, threads=2)
    out = str(tmp_path / "o.sam")
    rng = np.random.default_rng(3)
    newq = rng.integers(0, 94, size=b.qual_pool.shape[0]).astype(np.uint8)
    exp_lines = []
    a_base = g_base = 0
    with hostlib.BamReader(p, threads=3, keep_records=True) as r:
        w = hostlib.SamWriter(out, r, threads=1 if preset == "hifi" else 4)
        while True:
            fb = r.next_batch(7)
            if fb is None:
                break
            assert len(fb.rec_off) == fb.n_alns + 1 and fb.rec_off[-1] == len(fb.rec_pool)
            q0 = int(b.qual_off[a_base])
            for g in range(fb.n_groups):
                for a in range(int(fb.grp_aln_off[g]), int(fb.grp_aln_off[g + 1])):
                    rec = fb.rec_pool[int(fb.rec_off[a]):int(fb.rec_off[a + 1])]
                    # the record body is the BAM record: refID, pos, ..., l_seq at the spec's offsets
                    tid, pos = struct.unpack_from("<ii", rec.tobytes(), 0)
                    assert (tid, pos) == (int(fb.tid[a]), int(fb.pos[a]))
                    assert struct.unpack_from("<i", rec.tobytes(), 16)[0] == int(fb.l_qseq[a])
                    assert hostlib.format_sam_record(rec, s.names) == py_sam_line(fb, g, a, s.names)
                    lq = int(fb.l_qseq[a])
                    qa = newq[q0 + int(fb.qual_off[a]):q0 + int(fb.qual_off[a]) + lq]
                    assert hostlib.format_sam_record(rec, s.names, qual=qa) == py_sam_line(fb, g, a, s.names, qual=qa)
                    exp_lines.append(py_sam_line(fb, g, a, s.names, qual=qa))
            w.write_current_batch(r, newq[q0:q0 + int(fb.qual_off[-1])])
            a_base += fb.n_alns
            g_base += fb.n_groups
        hdr = r.header_text()
        w.close()
    assert a_base == b.n_alns and g_base == b.n_groups
    text = open(out).read()
    lines = text.splitlines(keepends=True)
    head = [ln for ln in lines if ln.startswith("@")]
    assert [ln for ln in head if ln.startswith("@SQ")] == [f"@SQ\tSN:{n}\tLN:{ln}\n" for n, ln in zip(s.names, s.lens)]
    if hdr:
        assert text.startswith(hdr.rstrip("\0"))
    assert lines[len(head):] == exp_lines


def test_sam_line_every_aux_type():
    """A hand-built record with every aux type, against the line htslib's sam_format1 prints."""
    qname = b"read/1\0"
    cigar = [(5 << 4) | 4, (10 << 4) | 0, (2 << 4) | 1, (3 << 4) | 2, (4 << 4) | 7, (1 << 4) | 8, (7 << 4) | 5]
    seq = "ACGTNACGTACGTTTGGCCAAG"  # 5S + 10M + 2I + 4= + 1X = 22 bases
    code = {c: i for i, c in enumerate(NT16)}
    nibs = [code[c] for c in seq] + [0]
    packed = bytes((nibs[i] << 4) | nibs[i + 1] for i in range(0, len(seq) + (len(seq) & 1), 2))
    qual = bytes(range(10, 10 + len(seq)))
    aux = (b"XAAQ" + b"Xcc" + struct.pack("<b", -5) + b"XCC" + struct.pack("<B", 200) + b"Xss" + struct.pack("<h", -300) +
           b"XSS" + struct.pack("<H", 60000) + b"Xii" + struct.pack("<i", -70000) + b"XII" + struct.pack("<I", 4000000000) +
           b"Xff" + struct.pack("<f", 0.5) + b"XZZhello world\0" + b"XHH1AE3\0" +
           b"XBBc" + struct.pack("<I", 2) + struct.pack("<bb", -1, 2) +
           b"YBBS" + struct.pack("<I", 3) + struct.pack("<HHH", 1, 2, 65535) +
           b"ZBBf" + struct.pack("<I", 2) + struct.pack("<ff", 1.5, -0.25) +
           b"csZ:10*ag:4\0")
    body = struct.pack("<iiBBHHHiiii", 1, 99, len(qname), 37, 4681, len(cigar), 16 | 256, len(seq), 0, 499, -120)
    body += qname + b"".join(struct.pack("<I", c) for c in cigar) + packed + qual + aux
    names = ["ctgA", "ctgB"]
    exp = ("read/1\t272\tctgB\t100\t37\t5S10M2I3D4=1X7H\tctgA\t500\t-120\t" + seq + "\t" +
           "".join(chr(q + 33) for q in qual) +
           "\tXA:A:Q\tXc:i:-5\tXC:i:200\tXs:i:-300\tXS:i:60000\tXi:i:-70000\tXI:i:4000000000\tXf:f:0.5"
           "\tXZ:Z:hello world\tXH:H:1AE3\tXB:B:c,-1,2\tYB:B:S,1,2,65535\tZB:B:f,1.5,-0.25\tcs:Z::10*ag:4\n")
    assert hostlib.format_sam_record(body, names) == exp
    # mate on the same contig prints '=', missing QUAL prints '*', empty SEQ prints '*' twice
    b2 = bytearray(body)
    struct.pack_into("<i", b2, 20, 1)
    assert hostlib.format_sam_record(bytes(b2), names).split("\t")[6] == "="
    qoff = 32 + len(qname) + 4 * len(cigar) + len(packed)
    b3 = bytearray(body)
    b3[qoff:qoff + len(seq)] = b"\xff" * len(seq)
    assert hostlib.format_sam_record(bytes(b3), names).split("\t")[10] == "*"
    b4 = struct.pack("<iiBBHHHiiii", -1, -1, len(qname), 0, 4680, 0, 4, 0, -1, -1, 0) + qname
    assert hostlib.format_sam_record(b4, names) == "read/1\t4\t*\t0\t0\t*\t*\t0\t0\t*\t*\n"
    # truncated records are rejected, not read past their end
    with pytest.raises(hostlib.HostError):
        hostlib.format_sam_record(body[:-3], names)
