"""Build the C part of the oracle (oracle/exact/tspn_exact.c) with gcc.

TEST INFRASTRUCTURE ONLY.  Output: oracle/exact/libtspn_exact.so (git-ignored, travels
to the GPU box with the snapshot).  ``oracle/_ref`` does not exist for this project: the
reference is pure Python (no C/C++ sources to compile) and cannot travel to the GPU box,
so its outputs are frozen under tests/golden/ instead (tests/golden/make_golden.py).
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "exact", "tspn_exact.c")
OUT = os.path.join(HERE, "exact", "libtspn_exact.so")


def build(force: bool = False) -> str:
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
           "-o", OUT, SRC, "-lm"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
