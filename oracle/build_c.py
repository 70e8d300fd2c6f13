#!/usr/bin/env python
"""Compile oracle/fq_oracle.c into oracle/_build/libfq_oracle.so (gcc; test infrastructure only).

-ffp-contract=off is the point: no FMA contraction, every fp32 operation rounds on its own.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "fq_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libfq_oracle.so")


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [os.environ.get("CC", "gcc"), "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c99", "-shared", "-fPIC",
           "-o", LIB, SRC, "-lm"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("gcc failed building the C oracle")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
