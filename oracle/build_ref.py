#!/usr/bin/env python
"""Compile the REAL reference hot path from /root/reference into oracle/_ref/ (binaries only).

TEST INFRASTRUCTURE ONLY -- nothing under oracle/ is imported by the product package.

What it builds (sources are read where they lie; every intermediate -- the staged copies
Cython wants next to its output and the generated .c files -- goes to a temp dir outside the
repository, only the compiled extension modules are kept):

    oracle/_ref/tiddit/DBSCAN.*.so           <- /root/reference/tiddit/DBSCAN.py   (cythonized as-is)
    oracle/_ref/tiddit/tiddit_cluster.*.so   <- /root/reference/tiddit/tiddit_cluster.pyx
    oracle/_ref/tiddit/tiddit_coverage.*.so  <- /root/reference/tiddit/tiddit_coverage.pyx
    oracle/_ref/tiddit/tiddit_gc.*.so        <- /root/reference/tiddit/tiddit_gc.pyx
    oracle/_ref/tiddit/tiddit_coverage_analysis.*.so <- /root/reference/tiddit/tiddit_coverage_analysis.pyx
    oracle/_ref/tiddit/tiddit_signal.*.so    <- /root/reference/tiddit/tiddit_signal.pyx  (against the pysam stand-in)
    oracle/_ref/tiddit/__init__.py           (empty, generated)
    oracle/_ref/pysam/                       (oracle/ref_shims/pysam: our FastaFile / AlignmentFile / AlignedSegment
                                              stand-ins, libcalignmentfile compiled so that the reference's cimport resolves)

The upstream package runs DBSCAN.py interpreted (setup.py only cythonizes the .pyx files);
compiling it keeps Python semantics (same set iteration, same numpy calls) and is a little
FASTER than the interpreter, so timing it as the CPU baseline is conservative.

oracle/_ref/ is git-ignored but NOT gpurun-ignored: the binaries travel to the GPU box, the
reference sources do not (and /root/reference does not exist there).
"""
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TIDDIT_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
MODULES = [("DBSCAN", "DBSCAN.py"), ("tiddit_cluster", "tiddit_cluster.pyx"),
           ("tiddit_coverage", "tiddit_coverage.pyx"), ("tiddit_gc", "tiddit_gc.pyx"),
           ("tiddit_coverage_analysis", "tiddit_coverage_analysis.pyx"), ("tiddit_signal", "tiddit_signal.pyx")]


def have_ref():
    shim = os.path.join(OUT, "pysam")
    if not (os.path.isdir(shim) and any(f.startswith("libcalignmentfile.") and f.endswith(".so") for f in os.listdir(shim))):
        return False
    return os.path.isdir(os.path.join(OUT, "tiddit")) and all(
        any(f.startswith(name + ".") and f.endswith(".so") for f in os.listdir(os.path.join(OUT, "tiddit")))
        for name, _ in MODULES)


def build(force=False):
    if not os.path.isdir(os.path.join(REF, "tiddit")):
        if have_ref():
            return True
        print("build_ref: %s not present and no prebuilt oracle/_ref -- skipped" % REF)
        return False
    if have_ref() and not force:
        shim_dir = os.path.join(HERE, "ref_shims", "pysam")
        newest_src = max([os.path.getmtime(os.path.join(REF, "tiddit", src)) for _, src in MODULES] +
                         [os.path.getmtime(os.path.join(shim_dir, f)) for f in os.listdir(shim_dir)])
        oldest_out = min(os.path.getmtime(os.path.join(d, f)) for d in (os.path.join(OUT, "tiddit"), os.path.join(OUT, "pysam"))
                         for f in os.listdir(d) if f.endswith(".so"))
        if oldest_out > newest_src:
            return True
    import numpy
    from Cython.Build import cythonize  # noqa: F401  (fail early if missing)

    tmp = tempfile.mkdtemp(prefix="tdt_ref_build_")
    try:
        stage = os.path.join(tmp, "tiddit")
        os.makedirs(stage)
        open(os.path.join(stage, "__init__.py"), "w").close()
        for _, src in MODULES:
            shutil.copy(os.path.join(REF, "tiddit", src), os.path.join(stage, src))
        shutil.copytree(os.path.join(HERE, "ref_shims", "pysam"), os.path.join(tmp, "pysam"))
        setup_py = os.path.join(tmp, "setup.py")
        with open(setup_py, "w") as f:
            f.write(
                "from setuptools import setup\n"
                "from Cython.Build import cythonize\n"
                "import Cython.Compiler.Options as O\n"
                "O.error_on_unknown_names = False  # DBSCAN.py:27 names an undefined helper in dead code\n"
                "import numpy\n"
                "setup(name='tiddit_ref', ext_modules=cythonize(%r, language_level=3, include_path=['.']),\n"
                "      include_dirs=[numpy.get_include()])\n"
                % ([os.path.join("pysam", "libcalignmentfile.pyx")] + [os.path.join("tiddit", s) for _, s in MODULES]))
        env = dict(os.environ, PYTHONPATH=tmp + os.pathsep + os.environ.get("PYTHONPATH", ""))
        subprocess.check_call([sys.executable, setup_py, "-q", "build_ext", "--inplace"], cwd=tmp, env=env)
        os.makedirs(os.path.join(OUT, "tiddit"), exist_ok=True)
        open(os.path.join(OUT, "tiddit", "__init__.py"), "w").close()
        suffix = sysconfig.get_config_var("EXT_SUFFIX")
        for name, _ in MODULES:
            shutil.copy(os.path.join(stage, name + suffix), os.path.join(OUT, "tiddit", name + suffix))
        if os.path.exists(os.path.join(OUT, "pysam.py")):
            os.remove(os.path.join(OUT, "pysam.py"))
        os.makedirs(os.path.join(OUT, "pysam"), exist_ok=True)
        shutil.copy(os.path.join(HERE, "ref_shims", "pysam", "__init__.py"), os.path.join(OUT, "pysam", "__init__.py"))
        shutil.copy(os.path.join(tmp, "pysam", "libcalignmentfile" + suffix),
                    os.path.join(OUT, "pysam", "libcalignmentfile" + suffix))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    print("build_ref: built", ", ".join(n for n, _ in MODULES), "->", OUT)
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    sys.exit(0 if ok else 1)
