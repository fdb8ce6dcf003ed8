"""TEST INFRASTRUCTURE ONLY — ctypes loader for oracle/librdf_oracle.so (see rdf_oracle.c).

Parity: see the header of oracle/rdf_oracle.py (control flow pinned, leaf arithmetic restated).  Never imported by the product.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "librdf_oracle.so")
ATOM_MAJOR, FRAME_MAJOR = 0, 1


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "rdf_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.rdf_oracle_counts.restype = ctypes.c_int64
        _lib.rdf_oracle_counts.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p,
            ctypes.c_int, ctypes.c_void_p, ctypes.c_float, ctypes.c_int, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_int]
        _lib.rdf_oracle_max_threads.restype = ctypes.c_int
    return _lib


def rdf_counts_c(positions: np.ndarray, particles_list, box_array, cutoff: float, nbins: int,
                 skip_first: bool = True, layout: int = ATOM_MAJOR, n_threads: int = 0):
    """positions: (N, C, 3) if layout == ATOM_MAJOR (the reference's positions_tensor) else
    (C, N, 3).  Returns ((n_pairs, nbins) int64 counts, triples evaluated)."""
    pos = np.ascontiguousarray(positions, dtype=np.float32)
    if layout == ATOM_MAJOR:
        n_atoms, n_frames = pos.shape[0], pos.shape[1]
    else:
        n_frames, n_atoms = pos.shape[0], pos.shape[1]
    parts = np.ascontiguousarray(particles_list, dtype=np.int64)
    box = np.ascontiguousarray(box_array, dtype=np.float32)
    n_pairs = len(parts) * (len(parts) + 1) // 2
    out = np.zeros((n_pairs, nbins), dtype=np.int64)
    ev = lib().rdf_oracle_counts(pos.ctypes.data, n_atoms, n_frames, layout, parts.ctypes.data,
                                 len(parts), box.ctypes.data, np.float32(cutoff), nbins,
                                 int(skip_first), out.ctypes.data, n_threads)
    if ev < 0:
        raise ValueError("rdf_oracle_counts: bad arguments")
    return out, int(ev)


def max_threads() -> int:
    return int(lib().rdf_oracle_max_threads())
