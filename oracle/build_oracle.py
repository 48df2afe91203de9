"""Builds the plain-C oracle (oracle/unet_ref.c -> oracle/_build/libunet_ref.so).
Test infrastructure only.  The reference is pure Python, so there is nothing
to compile into oracle/_ref ("Oracle")."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "libunet_ref.so")


def build(force=False):
    src = os.path.join(HERE, "unet_ref.c")
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(src):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-std=c99", "-o", OUT, src, "-lm"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
