"""Builds the C++/OpenMP oracle port (oracle/ref_port.cpp) into oracle/_build/.

The reference itself (MCMCDiagnosticTools.jl) is 100 % Julia: there is nothing to compile
from /root/reference, so no oracle/_ref/ is produced (see DESIGN.md "Oracle")."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ref_port.cpp")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libmcd_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    subprocess.run(["g++", "-O3", "-march=x86-64-v3", "-std=c++17", "-fopenmp", "-fPIC", "-shared",
                    "-ffp-contract=off", "-o", LIB, SRC], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
