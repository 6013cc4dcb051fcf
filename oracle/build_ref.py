"""Build the reference's own Cython evaluators into ``oracle/_ref`` (TEST INFRASTRUCTURE).

The sources are compiled *where they lie* under ``/root/reference`` (they are staged
into a throw-away directory under /tmp only so that Cython can resolve the
``smol.utils.cluster`` cimports; nothing is copied into this repository).  Only the
built ``.so`` files plus empty package ``__init__.py`` markers land in ``oracle/_ref``.

Flags follow the reference build: ``-O3 -ffast-math`` (``setup.py:17-25``) and the
Cython directives of ``tools/build_helpers.py:195-203``.  OpenMP is enabled when the
compiler can link ``-fopenmp`` (the reference's own probe/fallback,
``tools/build_helpers.py:100-183``).

Usage:  python oracle/build_ref.py [--reference /root/reference] [--force]
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

# (relative dir in the reference, file stem)
PYX = [
    ("smol/utils/cluster", "container"),
    ("smol/utils/cluster", "evaluator"),
    ("smol/utils/cluster", "ewald"),
    ("smol/utils/cluster", "correlations"),
    ("smol/utils", "_openmp_helpers"),
]
PXD = ["smol/utils/cluster/container.pxd", "smol/utils/cluster/evaluator.pxd",
       "smol/utils/cluster/struct.pxd"]


def _openmp_ok(cc: str) -> bool:
    code = "#include <omp.h>\nint main(){return omp_get_max_threads()>0?0:1;}\n"
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        with open(src, "w") as f:
            f.write(code)
        r = subprocess.run([cc, "-fopenmp", src, "-o", os.path.join(d, "t")],
                           capture_output=True)
        return r.returncode == 0


def built() -> bool:
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    return all(os.path.exists(os.path.join(OUT, d, s + ext)) for d, s in PYX)


def build(reference: str = "/root/reference", force: bool = False) -> bool:
    """Return True when ``oracle/_ref`` holds the built reference extensions."""
    if built() and not force:
        return True
    if not os.path.isdir(os.path.join(reference, "smol", "utils", "cluster")):
        return built()
    import numpy
    from Cython.Build import cythonize
    from Cython.Compiler import Options  # noqa: F401

    ext_suffix = sysconfig.get_config_var("EXT_SUFFIX")
    cands = [c for c in ("/usr/bin/gcc", os.environ.get("CC"), "gcc") if c and shutil.which(c)]
    cc = next((c for c in cands if _openmp_ok(c)), cands[0])
    omp = _openmp_ok(cc)
    stage = tempfile.mkdtemp(prefix="smolref_")
    try:
        for d, s in PYX:
            os.makedirs(os.path.join(stage, d), exist_ok=True)
            shutil.copy(os.path.join(reference, d, s + ".pyx"), os.path.join(stage, d))
        for p in PXD:
            shutil.copy(os.path.join(reference, p), os.path.join(stage, os.path.dirname(p)))
        for d in ("smol", "smol/utils", "smol/utils/cluster"):
            open(os.path.join(stage, d, "__init__.py"), "w").close()
        cwd = os.getcwd()
        os.chdir(stage)
        try:
            cythonize(
                [os.path.join(d, s + ".pyx") for d, s in PYX],
                include_path=[numpy.get_include(), stage],
                compiler_directives={
                    "language_level": 3, "boundscheck": False, "nonecheck": False,
                    "wraparound": False, "initializedcheck": False, "cdivision": True,
                },
                quiet=True,
            )
        finally:
            os.chdir(cwd)
        inc = [f"-I{sysconfig.get_paths()['include']}", f"-I{numpy.get_include()}"]
        flags = ["-O3", "-ffast-math", "-fPIC", "-shared", "-w",
                 "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION"]
        if omp:
            flags.append("-fopenmp")
        for d, s in PYX:
            out_dir = os.path.join(OUT, d)
            os.makedirs(out_dir, exist_ok=True)
            cmd = [cc, *flags, *inc, os.path.join(stage, d, s + ".c"),
                   "-o", os.path.join(out_dir, s + ext_suffix)]
            subprocess.run(cmd, check=True)
        for d in ("smol", "smol/utils", "smol/utils/cluster"):
            with open(os.path.join(OUT, d, "__init__.py"), "w") as f:
                f.write("# marker written by oracle/build_ref.py (not reference source)\n")
        with open(os.path.join(OUT, "BUILD_INFO.txt"), "w") as f:
            f.write(f"cc={cc}\nflags={' '.join(flags)}\nopenmp={omp}\nreference={reference}\n")
    finally:
        shutil.rmtree(stage, ignore_errors=True)
    return built()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    ok = build(a.reference, a.force)
    print("oracle/_ref built:", ok)
    sys.exit(0 if ok else 1)
