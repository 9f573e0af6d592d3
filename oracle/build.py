"""Build oracle/_portc.so from oracle/portc.c (TEST INFRASTRUCTURE).

Called by ``__graft_entry__.build()`` and lazily by ``oracle.port`` when the shared
object is missing or older than its source.
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "portc.c")
OUT = os.path.join(HERE, "_portc.so")


def build(force: bool = False) -> str:
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-o", OUT, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
