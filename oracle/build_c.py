"""Build the C restatement of the oracle (TEST INFRASTRUCTURE): oracle/liblmc_oracle.so."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "lmc_oracle.c")
LIB = os.path.join(HERE, "liblmc_oracle.so")


def build(force: bool = False) -> str:
    if os.path.exists(LIB) and not force and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cands = [c for c in ("/usr/bin/gcc", os.environ.get("CC"), "gcc") if c and shutil.which(c)]
    err = None
    for cc in cands:
        for omp in (["-fopenmp"], []):
            # no fast-math, no FMA contraction: bit-for-bit with the Python restatement
            cmd = [cc, "-O2", "-ffp-contract=off", "-fPIC", "-shared", *omp, SRC, "-o", LIB, "-lm"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode == 0:
                return LIB
            err = r.stderr
    raise RuntimeError("could not build oracle C library: " + str(err))


if __name__ == "__main__":
    print(build(force=True))
