"""Build recipe for the oracle's C restatement (test infrastructure).

`python -m oracle.build` or `oracle.build.build()` compiles
oracle/amg_core_restated.c into oracle/_build/liboracle.so with gcc.
The reference itself is pure Python (no C/C++ to compile), so there is no
`oracle/_ref` binary for this repository.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "amg_core_restated.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "liboracle.so")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", OUT, SRC])
    return OUT


if __name__ == "__main__":
    print(build(force=True))
