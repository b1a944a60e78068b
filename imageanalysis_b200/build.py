"""Build libiamatch.so for sm_100a with nvcc (in-tree, so the .so travels to
the GPU box with the repo snapshot).  `python -m imageanalysis_b200.build`."""
from __future__ import annotations

import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "lib", "libiamatch.so")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(_HERE, "..", "include", "iamatch.h")]
    return any(os.path.getmtime(s) > t for s in srcs if os.path.isfile(s))


def build(force: bool = False, verbose: bool = True) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... (see csrc/Makefile)."""
    if force:
        subprocess.run(["make", "-C", CSRC, "clean"], check=True, capture_output=not verbose)
    if force or needs_build():
        r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
        if verbose:
            sys.stdout.write(r.stdout[-4000:])
        if r.returncode != 0:
            sys.stderr.write(r.stdout[-8000:] + r.stderr[-8000:])
            raise RuntimeError("nvcc build of libiamatch.so failed")
    if not os.path.exists(LIB):
        raise RuntimeError("libiamatch.so missing after build")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
