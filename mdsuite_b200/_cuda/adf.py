"""
Python binding (ctypes) of the C ABI in include/mds_adf.h — the angular-distribution path of
``libmds_rdf.so``.  The ADF calculator routes the triple search and angle binning of its batch loop
(reference mdsuite/calculators/angular_distribution_function.py:374-405) here.  No CPU fallback:
without the built library or a CUDA device the calls raise.
"""
from __future__ import annotations

import ctypes
import itertools
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

from mdsuite_b200 import _cuda
from mdsuite_b200._cuda import MdsCudaError

ABI_VERSION = 1
MAX_NEIGHBOURS = 1024
RANGE_HI = 3.15

EXPORTED_SYMBOLS = (
    "mds_adf_create", "mds_adf_destroy", "mds_adf_compute", "mds_adf_compute_device",
    "mds_adf_stats_get", "mds_adf_build_threshold_table", "mds_adf_neighbour_window",
    "mds_adf_last_error", "mds_adf_abi_version",
)


class _Config(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32),
        ("n_species", ctypes.c_int32),
        ("counts", ctypes.POINTER(ctypes.c_int64)),
        ("nbins", ctypes.c_int32),
        ("cutoff", ctypes.c_float),
        ("box", ctypes.c_float * 3),
        ("norm_power", ctypes.c_int32),
        ("device", ctypes.c_int32),
        ("max_frames_in_flight", ctypes.c_int32),
    ]


class _Stats(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32),
        ("n_combinations", ctypes.c_int32),
        ("frames", ctypes.c_int64),
        ("triples", ctypes.c_double),
        ("pair_tests", ctypes.c_double),
        ("kernel_launches", ctypes.c_int64),
        ("kernel_ms", ctypes.c_float),
        ("compute_ms", ctypes.c_float),
        ("max_neighbours", ctypes.c_int32),
        ("sm_count", ctypes.c_int32),
        ("cells", ctypes.c_int32 * 3),
        ("reserved", ctypes.c_int32),
        ("frames_far", ctypes.c_int64),
    ]


@dataclass
class ADFStats:
    n_combinations: int
    frames: int
    triples: float
    pair_tests: float
    kernel_launches: int
    kernel_ms: float
    compute_ms: float
    max_neighbours: int
    sm_count: int
    cells: tuple = (0, 0, 0)
    frames_far: int = 0


_declared = False


def load() -> ctypes.CDLL:
    global _declared
    lib = _cuda.load()
    if _declared:
        return lib
    vp, i64 = ctypes.c_void_p, ctypes.c_int64
    lib.mds_adf_create.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(_Config)]
    lib.mds_adf_destroy.argtypes = [vp]
    lib.mds_adf_destroy.restype = None
    lib.mds_adf_compute.argtypes = [vp, vp, i64, vp]
    lib.mds_adf_compute_device.argtypes = [vp, vp, i64, vp, vp]
    lib.mds_adf_stats_get.argtypes = [vp, ctypes.POINTER(_Stats)]
    lib.mds_adf_build_threshold_table.argtypes = [ctypes.c_int, vp]
    lib.mds_adf_neighbour_window.argtypes = [ctypes.c_float, ctypes.POINTER(ctypes.c_float),
                                             ctypes.POINTER(ctypes.c_float)]
    lib.mds_adf_last_error.restype = ctypes.c_char_p
    lib.mds_adf_abi_version.restype = ctypes.c_int
    if lib.mds_adf_abi_version() != ABI_VERSION:
        raise MdsCudaError(f"ADF ABI version {lib.mds_adf_abi_version()} != {ABI_VERSION}; rebuild")
    _declared = True
    return lib


def _check(rc: int, what: str):
    if rc != 0:
        msg = load().mds_adf_last_error()
        raise MdsCudaError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")


def species_combinations(n_species: int) -> List[Tuple[int, int, int]]:
    """Row order of the output: itertools.combinations_with_replacement(range(S), 3)."""
    return list(itertools.combinations_with_replacement(range(n_species), 3))


def build_threshold_table(nbins: int) -> np.ndarray:
    """cos_edges[0..nbins] (float32) — see include/mds_adf.h."""
    out = np.empty(int(nbins) + 1 if nbins >= 0 else 1, dtype=np.float32)
    _check(load().mds_adf_build_threshold_table(int(nbins), out.ctypes.data),
           "mds_adf_build_threshold_table")
    return out


def neighbour_window(cutoff: float) -> Tuple[np.float32, np.float32]:
    """(s_lo, s_hi): j is a neighbour of i iff s_lo <= |r_ij|^2 < s_hi."""
    lo, hi = ctypes.c_float(), ctypes.c_float()
    _check(load().mds_adf_neighbour_window(float(cutoff), ctypes.byref(lo), ctypes.byref(hi)),
           "mds_adf_neighbour_window")
    return np.float32(lo.value), np.float32(hi.value)


class ADFHandle:
    """One GPU-resident ADF engine for a fixed (species counts, bins, cutoff, box, norm_power)."""

    def __init__(self, counts: Sequence[int], nbins: int, cutoff: float, box: Sequence[float],
                 norm_power: int = 4, device: int = 0, max_frames_in_flight: int = 0):
        self._lib = load()
        self._h = ctypes.c_void_p()
        self.counts = [int(c) for c in counts]
        self.n_atoms = sum(self.counts)
        self.nbins = int(nbins)
        self.n_combinations = len(species_combinations(len(self.counts)))
        self.device = int(device)
        arr = (ctypes.c_int64 * len(self.counts))(*self.counts)
        cfg = _Config(ctypes.sizeof(_Config), len(self.counts), arr, self.nbins,
                      float(np.float32(cutoff)),
                      (ctypes.c_float * 3)(*[float(np.float32(b)) for b in box]),
                      int(norm_power), self.device, int(max_frames_in_flight))
        _check(self._lib.mds_adf_create(ctypes.byref(self._h), ctypes.byref(cfg)), "mds_adf_create")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.mds_adf_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def compute(self, positions: np.ndarray) -> np.ndarray:
        """(F, N, 3) float32 host positions -> (F, n_combinations, nbins) float64 weight sums."""
        pos = np.ascontiguousarray(positions, dtype=np.float32)
        if pos.ndim != 3 or pos.shape[1:] != (self.n_atoms, 3):
            raise ValueError(f"positions must have shape (F, {self.n_atoms}, 3), got {pos.shape}")
        out = np.empty((pos.shape[0], self.n_combinations, self.nbins), dtype=np.float64)
        _check(self._lib.mds_adf_compute(self._h, pos.ctypes.data, pos.shape[0], out.ctypes.data),
               "mds_adf_compute")
        return out

    def compute_device(self, positions, out=None, stream: int = None):
        """torch CUDA tensors: positions (F, N, 3) float32 -> out (F, n_combinations, nbins)
        float64 on the same device (allocated if not given)."""
        import torch
        if not (positions.is_cuda and positions.dtype == torch.float32 and positions.is_contiguous()
                and positions.dim() == 3 and tuple(positions.shape[1:]) == (self.n_atoms, 3)):
            raise ValueError("positions must be a contiguous CUDA float32 tensor (F, n_atoms, 3)")
        shape = (positions.shape[0], self.n_combinations, self.nbins)
        if out is None:
            out = torch.empty(shape, dtype=torch.float64, device=positions.device)
        elif not (out.is_cuda and out.dtype == torch.float64 and out.is_contiguous()
                  and tuple(out.shape) == shape):
            raise ValueError(f"out must be a contiguous CUDA float64 tensor {shape}")
        if stream is None:
            stream = torch.cuda.current_stream(positions.device).cuda_stream
        _check(self._lib.mds_adf_compute_device(self._h, positions.data_ptr(), positions.shape[0],
                                                out.data_ptr(), ctypes.c_void_p(stream)),
               "mds_adf_compute_device")
        return out

    def stats(self) -> ADFStats:
        raw = _Stats()
        raw.struct_size = ctypes.sizeof(_Stats)
        _check(self._lib.mds_adf_stats_get(self._h, ctypes.byref(raw)), "mds_adf_stats_get")
        return ADFStats(int(raw.n_combinations), int(raw.frames), float(raw.triples),
                        float(raw.pair_tests), int(raw.kernel_launches), float(raw.kernel_ms),
                        float(raw.compute_ms), int(raw.max_neighbours), int(raw.sm_count),
                        tuple(int(c) for c in raw.cells), int(raw.frames_far))


# ---- handle reuse (as for the RDF, mdsuite_b200._cuda.acquire_handle) ---------------------------
# Creating a handle allocates device buffers that then grow to the batch size; a calculator call on
# a few thousand configurations spends as long on that as on its kernels.  One parked handle per
# device; the next call with the same configuration takes it back.
_PARKED = {}


def _key(counts, nbins, cutoff, box, norm_power, device):
    return (tuple(int(c) for c in counts), int(nbins), float(np.float32(cutoff)),
            tuple(float(np.float32(b)) for b in box), int(norm_power), int(device))


def acquire_handle(counts, nbins, cutoff, box, norm_power=4, device=0) -> "ADFHandle":
    key = _key(counts, nbins, cutoff, box, norm_power, device)
    parked = _PARKED.pop(int(device), None)
    if parked is not None:
        if parked[0] == key:
            return parked[1]
        parked[1].close()
    h = ADFHandle(counts, nbins, cutoff, box, norm_power=norm_power, device=device)
    h._park_key = key
    return h


def park_handle(h: "ADFHandle"):
    key = getattr(h, "_park_key", None)
    if key is None or not h._h.value:
        h.close()
        return
    old = _PARKED.pop(h.device, None)
    if old is not None:
        old[1].close()
    _PARKED[h.device] = (key, h)


def drop_parked_handles():
    for _key_, h in list(_PARKED.values()):
        h.close()
    _PARKED.clear()
