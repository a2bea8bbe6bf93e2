"""not gpu: the C-ABI library loads and exports every function include/secphase_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "secphase_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(sp_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_declares_the_documented_entry_points():
    names = declared_functions()
    for must in ("sp_create", "sp_destroy", "sp_set_reference_codes", "sp_submit", "sp_wait", "sp_hmm_batch"):
        assert must in names


def test_library_exports_every_declared_symbol():
    import secphase_b200
    from secphase_b200.build import build
    build()
    lib = ctypes.CDLL(secphase_b200.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/secphase_b200.h but not exported"
    assert set(secphase_b200.api.EXPORTED_SYMBOLS) == set(declared_functions())


def test_every_binding_is_typed():
    """ctypes passes an untyped Python int as a 32-bit C int, which truncates a 64-bit sp_ctx*: every
    entry point that takes arguments must have argtypes set by load_library()."""
    import secphase_b200
    L = secphase_b200.load_library()
    for name in secphase_b200.api.EXPORTED_SYMBOLS:
        fn = getattr(L, name)
        if name in ("sp_last_error", "sp_version"):  # no parameters
            assert fn.restype is ctypes.c_char_p
            continue
        assert fn.argtypes is not None, f"{name}: argtypes not set"


def test_host_library_exports_every_declared_symbol():
    """include/secphase_host.h <-> libsecphase_host.so (BAM/FASTA ingest, BED/out.log writers)."""
    from secphase_b200 import hostlib
    from secphase_b200.build import build_host
    build_host()
    text = open(os.path.join(ROOT, "include", "secphase_host.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = sorted(set(re.findall(r"\b(sph_[a-z0-9_]+)\s*\(", text)))
    assert {"sph_bam_open", "sph_bam_next_batch", "sph_fasta_load", "sph_blocks_merge_v2",
            "sph_format_marker_record"} <= set(names)
    lib = ctypes.CDLL(hostlib.LIB_PATH)
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/secphase_host.h but not exported"
    assert os.access(hostlib.CLI_PATH, os.X_OK)


def test_no_cpu_fallback_without_device():
    """On a GPU-less box creating a context must fail loudly (and on a GPU box this is skipped)."""
    import secphase_b200
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(secphase_b200.SecphaseError):
        secphase_b200.Secphase("hifi")


def test_params_presets_follow_the_reference_code():
    import secphase_b200
    h = secphase_b200.params_for("hifi")
    o = secphase_b200.params_for("ont")
    # secphase.c:477-504 (NOT the README values)
    assert (h.indel_threshold, h.set_q, h.min_q, h.min_score, h.prim_margin_score, h.conf_d) == (10, 40, 10, -10, 40.0, 1e-4)
    assert (o.indel_threshold, o.set_q, o.min_q, o.min_score, o.prim_margin_score, o.conf_d) == (20, 20, 10, -10, 20.0, 1e-3)
    assert h.baq_flag == 1 and h.consensus == 1 and h.flank_margin == 500
    d = secphase_b200.params_for(None)
    assert d.baq_flag == 0 and d.consensus == 0
    r = secphase_b200.params_for("hifi", prim_margin_score=20.0, min_score=-50)  # README values stay reachable
    assert r.prim_margin_score == 20.0 and r.min_score == -50
