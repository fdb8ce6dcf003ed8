"""Build libmds_rdf.so in-tree for sm_100a:  python -m mdsuite_b200._cuda.build [--force]"""
from __future__ import annotations

import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")


def build(force: bool = False, verbose: bool = False, checked: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... (see csrc/Makefile).
    checked=True also builds libmds_rdf_checked.so (-DMDS_CHECKED)."""
    cmd = ["make", "-C", CSRC, "all"] + (["checked"] if checked else [])
    if force:
        cmd.append("-B")
    if not verbose:
        cmd.append("-s")
    subprocess.check_call(cmd)
    out = os.path.join(_HERE, "libmds_rdf.so")
    if not os.path.exists(out):
        raise RuntimeError("build finished but libmds_rdf.so is missing")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
