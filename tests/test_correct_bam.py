"""`bin/correct_bam` (host only): the consumer of out.log, kept as a drop-in for the reference's
programs/src/correct_bam.c.  Its record-by-record decisions are checked against a Python
restatement of correct_bam.c:32-105,168-214,354-380 on a synthetic BAM whose records carry
de:f tags, mixed MAPQs, supplementary / unmapped records and short reads."""
import gzip
import os
import struct
import subprocess

import numpy as np
import pytest

from secphase_b200 import hostlib
from secphase_b200.build import CORRECT_BAM
from tests.conftest import make_case


def parse_bam(path):
    """-> (header bytes, names, [record body bytes]) with Python's own gzip (BGZF = multi-member gzip)."""
    raw = gzip.open(path, "rb").read()
    assert raw[:4] == b"BAM\1"
    l_text = struct.unpack_from("<i", raw, 4)[0]
    o = 8 + l_text
    n_ref = struct.unpack_from("<i", raw, o)[0]
    o += 4
    names = []
    for _ in range(n_ref):
        ln = struct.unpack_from("<i", raw, o)[0]
        names.append(raw[o + 4:o + 4 + ln - 1].decode())
        o += 4 + ln + 4
    header = raw[:o]
    recs = []
    while o < len(raw):
        bs = struct.unpack_from("<i", raw, o)[0]
        recs.append(raw[o + 4:o + 4 + bs])
        o += 4 + bs
    return header, names, recs


def fields(rec):
    tid, pos, l_qname, mapq, _bin, n_cigar, flag, l_seq = struct.unpack_from("<iiBBHHHi", rec, 0)
    qname = rec[32:32 + l_qname].split(b"\0")[0].decode()
    cig = struct.unpack_from(f"<{n_cigar}I", rec, 32 + l_qname)
    aux_off = 32 + l_qname + 4 * n_cigar + (l_seq + 1) // 2 + l_seq
    return dict(tid=tid, pos=pos, mapq=mapq, flag=flag, qname=qname, cigar=cig, aux_off=aux_off, l_qname=l_qname)


def build_input(tmp_path):
    """A BAM with: the synthetic read groups (de:f appended), one group made supplementary-bearing,
    an unmapped record, a short read; plus a phasing log, a MAPQ table and an exclude list."""
    s, b, _, _ = make_case("hifi", 40, locus_len=200000, len_mean=7000, len_sd=1500, len_min=3000)
    src = str(tmp_path / "src.bam")
    hostlib.write_bam(src, s.names, s.lens, b)
    header, names, recs = parse_bam(src)
    rng = np.random.default_rng(7)
    out = []
    for k, r in enumerate(recs):
        r = bytearray(r)
        f = fields(r)
        r[9] = int(rng.integers(0, 61))                       # MAPQ
        de = float(rng.choice([0.0, 0.01, 0.05, 0.2]))        # 0.2 > default maxDiv
        r += b"def" + struct.pack("<f", de)
        if k % 17 == 5:
            struct.pack_into("<H", r, 14, f["flag"] | 0x800)  # supplementary
        out.append(bytes(r))
    # an unmapped record and a record of a read that will be excluded
    unm = bytearray(out[0]); struct.pack_into("<H", unm, 14, 4); out.insert(3, bytes(unm))
    path = str(tmp_path / "in.bam")
    with open(path, "wb") as fh:
        body = header + b"".join(struct.pack("<i", len(r)) + r for r in out)
        # BGZF: one gzip member per <= 60000 bytes with the BC extra field, then the EOF marker
        import zlib
        for i in range(0, len(body), 60000):
            chunk = body[i:i + 60000]
            co = zlib.compressobj(6, zlib.DEFLATED, -15)
            comp = co.compress(chunk) + co.flush()
            bsize = len(comp) + 25
            fh.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize) + comp +
                     struct.pack("<II", zlib.crc32(chunk), len(chunk)))
        fh.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    # phasing log in out.log's grammar (secphase.c:32-57,197-200) for every third group: '@' = its first
    # secondary; once with the primary line first, once with the secondary line first; one record whose
    # '@' starts where the '*' does (ignored by the reference)
    log = []
    chosen = {}
    for g in range(0, b.n_groups, 3):
        a0, a1 = int(b.grp_aln_off[g]), int(b.grp_aln_off[g + 1])
        qn = bytes(b.qname_pool[int(b.qname_off[g]):int(b.qname_off[g + 1])]).decode()
        prim = [a for a in range(a0, a1) if not (int(b.flag[a]) & 0x100)][0]
        sec = [a for a in range(a0, a1) if int(b.flag[a]) & 0x100][0]
        lines = []
        for a in range(a0, a1):
            tag = "*" if a == prim else ("@" if a == sec else "!")
            lines.append(f"{tag}\t-12.50\t{s.names[int(b.tid[a])]}\t{int(b.pos[a])}\t{int(b.pos[a]) + 100}\n")
        log.append("#MARKER SCORE\n$\t" + qn + "\n" + "".join(lines) + "\n")
        if (s.names[int(b.tid[prim])], int(b.pos[prim])) != (s.names[int(b.tid[sec])], int(b.pos[sec])):
            chosen[qn] = (s.names[int(b.tid[sec])], int(b.pos[sec]))
    log.append("#MARKER SCORE\n$\tsame_place\n*\t-1.00\tctgX\t100\t900\n@\t-0.50\tctgX\t100\t800\n\n")
    log_path = str(tmp_path / "p.out.log")
    open(log_path, "w").write("".join(log))
    # mapq table: 1-based starts; exclude list
    mq_path = str(tmp_path / "mapq.tsv")
    mq = {}
    with open(mq_path, "w") as fh:
        for k in range(0, len(recs), 5):
            f = fields(recs[k])
            fh.write(f"{f['qname']}\t{names[f['tid']]}\t{f['pos'] + 1}\t{(k * 7) % 90}\n")
            mq.setdefault(f["qname"], []).append((names[f["tid"]], f["pos"], (k * 7) % 90))
    ex_path = str(tmp_path / "exclude.txt")
    excl = {fields(recs[8])["qname"]}
    open(ex_path, "w").write("\n".join(excl) + "\n")
    return path, names, out, chosen, log_path, mq, mq_path, excl, ex_path


def expected(names, recs, chosen, mq, excl, primary_only, no_tag, min_read, min_aln, max_mapq, max_div):
    """correct_bam.c:354-380 restated."""
    res = []
    for r in recs:
        f = fields(r)
        if f["flag"] & 4 or f["qname"] in excl:
            continue
        contig = names[f["tid"]]
        if f["qname"] in chosen:
            prim = chosen[f["qname"]] == (contig, f["pos"])
        else:
            prim = not (f["flag"] & 0x100)
        flag = f["flag"] & ~0x100 if prim else f["flag"] | 0x100
        if not prim and primary_only:
            continue
        rl = sum(c >> 4 for c in f["cigar"] if (c & 15) in (0, 7, 8, 1, 4, 5))
        al = sum(c >> 4 for c in f["cigar"] if (c & 15) in (0, 7, 8))
        if rl < min_read or al < min_aln:
            continue
        mapq = f["mapq"]
        for ctg, st, m in mq.get(f["qname"], []):
            if ctg == contig and st == f["pos"]:
                mapq = m & 255
                break
        if max_mapq < mapq:
            continue
        aux = r[f["aux_off"]:]
        de = struct.unpack_from("<f", aux, aux.index(b"def") + 3)[0] if b"def" in aux else 0.0
        if max_div < de:
            continue
        out = bytearray(r)
        out[9] = mapq
        struct.pack_into("<H", out, 14, flag)
        res.append(bytes(out[:f["aux_off"]] if no_tag else out))
    return res


@pytest.mark.parametrize("opts,kw", [
    ([], {}),
    (["-p", "-t"], dict(primary_only=True, no_tag=True)),
    (["--minReadLen", "6500", "--minAlignmentLen", "6000", "--maxMapq", "50", "--maxDiv", "0.02", "-n", "1"],
     dict(min_read=6500, min_aln=6000, max_mapq=50, max_div=0.02)),
])
def test_correct_bam_matches_the_reference_logic(tmp_path, opts, kw):
    if not os.path.exists(CORRECT_BAM):
        from secphase_b200.build import build_host
        build_host()
    path, names, recs, chosen, log_path, mq, mq_path, excl, ex_path = build_input(tmp_path)
    assert chosen, "the log must re-phase some reads"
    out = str(tmp_path / "out.bam")
    r = subprocess.run([CORRECT_BAM, "-i", path, "-o", out, "-P", log_path, "-M", mq_path, "-e", ex_path] + opts,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    header_in, _, _ = parse_bam(path)
    header_out, names_out, got = parse_bam(out)
    assert header_out == header_in and names_out == names        # sam_hdr_write: header copied through
    par = dict(primary_only=False, no_tag=False, min_read=5000, min_aln=5000, max_mapq=100, max_div=0.12)
    par.update(kw)
    want = expected(names, recs, chosen, mq, excl, **par)
    assert len(got) == len(want)
    assert got == want
    if not opts:
        flags = {(fields(x)["qname"], fields(x)["pos"]): fields(x)["flag"] for x in got}
        promoted = [qn for qn, (ctg, pos) in chosen.items() if (qn, pos) in flags]
        assert promoted
        for qn in promoted:
            assert not flags[(qn, chosen[qn][1])] & 0x100    # the selected secondary is now primary
            others = [fl for (q2, p2), fl in flags.items() if q2 == qn and p2 != chosen[qn][1]]
            assert all(fl & 0x100 for fl in others)          # and the old primary a secondary
    # the output is a valid input for the reader of the hot path
    with hostlib.BamReader(out, threads=2) as rd:
        n = sum(b.n_groups for b in rd)
    assert n >= 0


def test_correct_bam_usage_and_errors(tmp_path):
    if not os.path.exists(CORRECT_BAM):
        from secphase_b200.build import build_host
        build_host()
    r = subprocess.run([CORRECT_BAM, "-h"], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage: correct_bam  -i <INPUT_BAM> -o <OUTPUT_BAM>" in r.stderr   # correct_bam.c:289
    for opt in ("--phasingLog,\t-P", "--mapqTable,\t-M", "--primaryOnly,\t-p", "--maxDiv,\t-d"):
        assert opt in r.stderr
    r = subprocess.run([CORRECT_BAM, "-i", str(tmp_path / "nope.bam"), "-o", str(tmp_path / "o.bam")],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "nope.bam" in r.stderr
    bad = tmp_path / "bad.bam"
    bad.write_bytes(b"not a bam")
    r = subprocess.run([CORRECT_BAM, "-i", str(bad), "-o", str(tmp_path / "o.bam")], capture_output=True, text=True)
    assert r.returncode == 1
