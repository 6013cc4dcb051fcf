"""Build liblmc.so (sm_100a) in-tree with nvcc.  ``python -m smol_b200.build [--force]``."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
LIB = os.path.join(LIBDIR, os.environ.get("LMC_LIB_NAME", "liblmc.so"))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]
GROUPS = (4, 8, 16, 32)


def _sources():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + \
        [os.path.join(HERE, "..", "include", "lmc.h")]
    return deps


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(s) <= t for s in _sources())


def _deps(src):
    """files an object depends on: its source, every header, and (for the API unit only) nothing else"""
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    if "lmc_spec_inst" not in src and "lmc_spec_tf_inst" not in src and "lmc_spec_c64_inst" not in src:
        hdrs = [h for h in hdrs if not h.endswith("lmc_spec.cuh")]
    if "lmc_spec_tf_inst" not in src:
        hdrs = [h for h in hdrs if not h.endswith("lmc_spec_tf.cuh")]
    if "lmc_spec_c64_inst" not in src:
        hdrs = [h for h in hdrs if not h.endswith("lmc_spec_c64.cuh")]
    if "lmc_wl_inst" not in src:
        hdrs = [h for h in hdrs if not h.endswith("lmc_wl.cuh")]
    return [src, os.path.join(HERE, "..", "include", "lmc.h"), *hdrs]


def _compile(args):
    src, obj, extra, force = args
    if (not force and not os.environ.get("LMC_EXTRA_DEFS") and os.path.exists(obj)
            and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in _deps(src))):
        return ["(up to date)", obj], subprocess.CompletedProcess([], 0, "", "")
    cmd = [NVCC, *FLAGS, *extra, *os.environ.get("LMC_EXTRA_DEFS", "").split(), "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return cmd, r


def build(force: bool = False, verbose: bool = False) -> str:
    if up_to_date() and not force:
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj" + os.environ.get("LMC_LIB_NAME", "").replace(".so", ""))
    os.makedirs(objdir, exist_ok=True)
    jobs = [(os.path.join(CSRC, "lmc_api.cu"), os.path.join(objdir, "lmc_api.o"), [], force)]
    for g in GROUPS:
        for wl in (0, 1):
            jobs.append((os.path.join(CSRC, "lmc_run_inst.cu"),
                         os.path.join(objdir, f"lmc_run_g{g}_wl{wl}.o"), [f"-DLMC_G={g}", f"-DLMC_WL={wl}"], force))
    jobs.append((os.path.join(CSRC, "lmc_run_dist_inst.cu"), os.path.join(objdir, "lmc_run_dist.o"), [], force))
    jobs.append((os.path.join(CSRC, "lmc_wl_inst.cu"), os.path.join(objdir, "lmc_wl.o"), [], force))
    jobs.append((os.path.join(CSRC, "lmc_spec_inst.cu"), os.path.join(objdir, "lmc_spec.o"), [], force))
    jobs.append((os.path.join(CSRC, "lmc_spec_inst2.cu"), os.path.join(objdir, "lmc_spec2.o"), [], force))
    jobs.append((os.path.join(CSRC, "lmc_spec_inst3.cu"), os.path.join(objdir, "lmc_spec3.o"), [], force))
    jobs.append((os.path.join(CSRC, "lmc_spec_tf_inst.cu"), os.path.join(objdir, "lmc_spec_tf.o"), [], force))
    jobs.append((os.path.join(CSRC, "lmc_spec_c64_inst.cu"), os.path.join(objdir, "lmc_spec_c64.o"), [], force))
    log = []
    with ThreadPoolExecutor(max_workers=min(os.cpu_count() or 8, len(jobs))) as ex:
        for cmd, r in ex.map(_compile, jobs):
            log.append(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode != 0:
                sys.stderr.write(log[-1])
                raise RuntimeError("nvcc failed")
    with open(os.path.join(LIBDIR, "build.log"), "w") as f:
        f.write("\n".join(log))
    objs = [j[1] for j in jobs]
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
