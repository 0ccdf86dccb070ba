"""Build recipe for the CPU oracle (oracle/libmie_oracle.so).  TEST INFRASTRUCTURE.

There is no `oracle/_ref/`: the reference repository has no source files for this
path (0 lines of Python, no C/C++), so nothing of it can be compiled here.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "mie_oracle.c")
LIB = os.path.join(HERE, "libmie_oracle.so")

CFLAGS = ["-O2", "-fPIC", "-shared", "-std=gnu11", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-Wall"]


def build_oracle(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(
        os.path.getmtime(SRC), os.path.getmtime(os.path.abspath(__file__))
    ):
        return LIB
    gcc = shutil.which("gcc")
    if gcc is None:
        raise RuntimeError("gcc not found: cannot build the CPU oracle")
    tmp = LIB + f".tmp{os.getpid()}"
    r = subprocess.run([gcc, *CFLAGS, SRC, "-o", tmp, "-lm"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"oracle build failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build_oracle(force=True))
