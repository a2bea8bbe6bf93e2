"""Build libsecphase_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libsecphase_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",            # the HMM also uses explicit __dmul_rn/__dadd_rn; belt and braces
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def sources():
    return [os.path.join(CSRC, "sp_api.cu")]


def deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    d.append(os.path.join(os.path.dirname(HERE), "include", "secphase_b200.h"))
    d.append(os.path.join(os.path.dirname(HERE), "include", "sp_flat_batch.h"))
    return d


def build_variant(name, extra_flags):
    """Tuning aid: libsecphase_b200 with extra nvcc flags (e.g. -DSP_H2_NCR=12) under lib/variants/;
    select it at run time with SECPHASE_B200_LIB=<path>."""
    vdir = os.path.join(LIBDIR, "variants")
    os.makedirs(vdir, exist_ok=True)
    out = os.path.join(vdir, f"lib_{name}.so")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    subprocess.check_call([nvcc] + [f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")] + list(extra_flags) + ["-o", out] + sources())
    return out


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    newest = max(os.path.getmtime(p) for p in deps())
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest:
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + sources()
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(LIBDIR, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed, see " + log)
    return LIB


HOST = os.path.join(HERE, "host")
HOST_LIB = os.path.join(LIBDIR, "libsecphase_host.so")
BINDIR = os.path.join(os.path.dirname(HERE), "bin")
CLI = os.path.join(BINDIR, "secphase")
CORRECT_BAM = os.path.join(BINDIR, "correct_bam")  # consumer of out.log (programs/src/correct_bam.c)
INDEX_TOOL = os.path.join(BINDIR, "secphase_index")  # programs/src/secphase_index.c
HOST_SRCS = ["sph_common.cpp", "sph_bgzf.cpp", "sph_inflate.cpp", "sph_bam.cpp", "sph_fasta.cpp", "sph_output.cpp", "sph_sam.cpp"]
CXX_FLAGS = ["-O2", "-g", "-std=c++17", "-fPIC", "-Wall", "-Wextra", "-pthread"]


def _stale(target, srcs):
    return not os.path.exists(target) or os.path.getmtime(target) < max(os.path.getmtime(p) for p in srcs)


def build_host(force=False):
    """libsecphase_host.so (BGZF/BAM/FASTA ingest, BED/out.log writers; g++ only, no CUDA) and the
    `secphase` executable that links it together with libsecphase_b200.so."""
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(BINDIR, exist_ok=True)
    if not os.path.exists(LIB):  # the secphase executable links the GPU library: build it first
        build()
    cxx = os.environ.get("CXX", "g++")
    inc = [os.path.join(os.path.dirname(HERE), "include", f) for f in ("secphase_host.h", "secphase_b200.h", "sp_flat_batch.h")]
    hdrs = [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".hpp")] + inc
    srcs = [os.path.join(HOST, f) for f in HOST_SRCS]
    if force or _stale(HOST_LIB, srcs + hdrs):
        subprocess.check_call([cxx] + CXX_FLAGS + ["-shared", "-o", HOST_LIB] + srcs + ["-lz"])
    main = os.path.join(HOST, "secphase_main.cpp")
    if force or _stale(CLI, [main, HOST_LIB, LIB] + inc):
        subprocess.check_call([cxx] + CXX_FLAGS + ["-o", CLI, main, "-L" + LIBDIR, "-lsecphase_host", "-lsecphase_b200",
                                                  "-Wl,-rpath,$ORIGIN/../secphase_b200/lib"])
    cb_main = os.path.join(HOST, "correct_bam_main.cpp")
    if force or _stale(CORRECT_BAM, [cb_main, HOST_LIB] + hdrs):
        subprocess.check_call([cxx] + CXX_FLAGS + ["-o", CORRECT_BAM, cb_main, "-L" + LIBDIR, "-lsecphase_host", "-lz",
                                                  "-Wl,-rpath,$ORIGIN/../secphase_b200/lib"])
    ix_main = os.path.join(HOST, "secphase_index_main.cpp")
    if force or _stale(INDEX_TOOL, [ix_main]):
        subprocess.check_call([cxx] + CXX_FLAGS + ["-o", INDEX_TOOL, ix_main, "-lz"])
    return HOST_LIB, CLI


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_host(force="--force" in sys.argv))
