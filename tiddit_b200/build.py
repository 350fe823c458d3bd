"""Build libtdt_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a, and libtdt_bam.so (the host-side
BAM scanner, include/tdt_bam.h) with g++ + zlib.

`python -m tiddit_b200.build` or `tiddit_b200.build.build()`.  nvcc cross-compiles without a GPU.
The built library is git-ignored but travels to the GPU box with the repository snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtdt_b200.so")
BAM_LIB = os.path.join(HERE, "libtdt_bam.so")
TAB_LIB = os.path.join(HERE, "libtdt_tab.so")
SOURCES = ["tdt_api.cu", "tdt_cluster.cu", "tdt_aggregate.cu", "tdt_coverage.cu", "tdt_ploidy.cu", "tdt_gc.cu", "tdt_peer.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O3,-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f not in ("tdt_bam.cpp", "tdt_tab.cpp", "tdt_inflate.h")] + [os.path.join(HERE, "..", "include", "tdt_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_bam(force=False):
    """g++ -O3 -shared: the BGZF/BAM column scanner (host code only, links zlib)."""
    src = os.path.join(CSRC, "tdt_bam.cpp")
    hdr = os.path.join(HERE, "..", "include", "tdt_bam.h")
    inflate = os.path.join(CSRC, "tdt_inflate.h")          # the block decoder the scanner includes
    if not force and os.path.exists(BAM_LIB) and os.path.getmtime(BAM_LIB) > max(map(os.path.getmtime, (src, hdr, inflate))):
        return BAM_LIB
    subprocess.check_call([os.environ.get("CXX", "g++"), "-O3", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden",
                           "-Wall", "-o", BAM_LIB, src, "-lz", "-lpthread"])
    return BAM_LIB


def build_tab(force=False):
    """g++ -O3 -shared: the signal tab-file scanner (host code only; include/tdt_tab.h)."""
    src = os.path.join(CSRC, "tdt_tab.cpp")
    hdr = os.path.join(HERE, "..", "include", "tdt_tab.h")
    if not force and os.path.exists(TAB_LIB) and os.path.getmtime(TAB_LIB) > max(os.path.getmtime(src), os.path.getmtime(hdr)):
        return TAB_LIB
    subprocess.check_call([os.environ.get("CXX", "g++"), "-O3", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden",
                           "-Wall", "-o", TAB_LIB, src, "-lpthread"])
    return TAB_LIB


def build(force=False, verbose=False, out=None, build_dir=None):
    """out / build_dir: build a variant (TDT_NVCC_DEFS) next to the default library without touching it."""
    build_bam(force)
    build_tab(force)
    if out is None and not force and not _stale():
        return LIB
    objs = []
    build_dir = build_dir or os.path.join(HERE, "_build")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(build_dir, src.replace(".cu", ".o"))
        defs = os.environ.get("TDT_NVCC_DEFS", "").split()     # tuning experiments: -DNAME=value ...
        cmd = [_nvcc()] + NVCC_FLAGS + defs + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("== nvcc %s ==\n%s\n" % (src, log))
        failed = failed or p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([_nvcc(), "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-o", out or LIB] + objs)
    return out or LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
