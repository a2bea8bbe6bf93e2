"""Comparison helpers shared by the parity tests (oracle vs hostsim vs CUDA)."""
import numpy as np

ASCII2CODE = np.full(256, 4, np.uint8)
for _c, _v in zip(b"ACGTacgt0123", [0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3]):
    ASCII2CODE[_c] = _v


def encode_reference(synth):
    """codes 0..4 (seq_nt16_int[seq_nt16_table[c]], ptMarker.c:744) + contig offsets."""
    parts = [ASCII2CODE[synth.contig_ascii(i)] for i in range(synth.n_contigs)]
    off = np.zeros(synth.n_contigs + 1, np.int64)
    off[1:] = np.cumsum([len(p) for p in parts])
    return np.concatenate(parts), off


def compare_results(ref, got, check_hmm_rows=True, label="got"):
    """ref: oracle.pyoracle.run() dict; got: dict with the same table names.  Returns a list of
    human-readable mismatch strings (empty = bit-exact)."""
    bad = []

    def eq(name, a, b):
        a = np.asarray(a)
        b = np.asarray(b)
        if a.shape != b.shape:
            bad.append(f"{name}: shape {a.shape} vs {b.shape}")
            return False
        if a.dtype.kind == "f":
            same = a.view(np.int64) == b.view(np.int64)
        else:
            same = a == b
        if not np.all(same):
            idx = np.argwhere(~same)[0]
            bad.append(f"{name}: first mismatch at {tuple(idx)}: oracle={a[tuple(idx)]!r} {label}={b[tuple(idx)]!r} "
                       f"({int((~same).sum())} differing)")
            return False
        return True

    eq("groups", ref["groups"], got["groups"])
    eq("scores(bits)", ref["scores"], got["scores"])
    eq("extents", ref["extents"], got["extents"])
    for st in ("markers_pre", "markers_baq", "markers_final"):
        eq(st + "_off", ref[st + "_off"], got[st + "_off"])
        eq(st, ref[st], got[st])
    eq("block_off", ref["block_off"], got["block_off"])
    eq("blocks", ref["blocks"], got["blocks"])
    return bad
