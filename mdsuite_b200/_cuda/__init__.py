"""
mdsuite_b200._cuda — Python binding (ctypes) of the C ABI in include/mds_rdf.h.

This is the extension the north star calls ``mdsuite/_cuda``: the RDF calculator routes the body
of its batch loop (reference mdsuite/calculators/radial_distribution_function.py:846-885) here.
There is NO CPU fallback: if ``libmds_rdf.so`` has not been built, or no CUDA device is usable,
the calls raise.

Build in-tree with ``python -m mdsuite_b200._cuda.build`` (or ``make -C mdsuite_b200/_cuda/csrc``).
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MDS_RDF_LIB selects another build of the same ABI, e.g. the self-checking libmds_rdf_checked.so
# (make -C csrc checked; see rdf_common.cuh MDS_CHECKED)
LIB_PATH = os.environ.get("MDS_RDF_LIB") or os.path.join(_HERE, "libmds_rdf.so")

ABI_VERSION = 7
MDS_OK = 0
PARITY_SKIP_FIRST = 0x1
PATH_ALLPAIRS = 0x2
PATH_CELLLIST = 0x4
LAYOUT_FRAME_MAJOR = 0  # (n_frames, n_atoms, 3)
LAYOUT_ATOM_MAJOR = 1  # (n_atoms, n_frames, 3) — the reference's positions_tensor

EXPORTED_SYMBOLS = (
    "mds_rdf_create", "mds_rdf_destroy", "mds_rdf_submit", "mds_rdf_submit_device",
    "mds_rdf_counts", "mds_rdf_device_counts", "mds_rdf_allreduce", "mds_rdf_join", "mds_rdf_wait", "mds_rdf_wait_host_copies", "mds_rdf_host_ticket", "mds_rdf_wait_host_ticket", "mds_rdf_plan_host_batches", "mds_rdf_synchronize",
    "mds_rdf_reset", "mds_rdf_set_item_shard",
    "mds_rdf_stats_get", "mds_rdf_threshold_table", "mds_rdf_build_threshold_table",
    "mds_bench_shared_atomics",
    "mds_last_error", "mds_abi_version",
)


class MdsCudaError(RuntimeError):
    """Raised for every non-zero status of the C ABI (message from mds_last_error())."""


class _Config(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32),
        ("n_species", ctypes.c_int32),
        ("counts", ctypes.POINTER(ctypes.c_int64)),
        ("nbins", ctypes.c_int32),
        ("cutoff", ctypes.c_float),
        ("box", ctypes.c_float * 3),
        ("flags", ctypes.c_uint32),
        ("device", ctypes.c_int32),
        ("max_frames_in_flight", ctypes.c_int32),
    ]


class _Stats(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32),
        ("path", ctypes.c_int32),
        ("frames", ctypes.c_int64),
        ("pairs_evaluated", ctypes.c_double),
        ("pairs_allpairs_equiv", ctypes.c_double),
        ("pairs_binned", ctypes.c_double),
        ("frames_general_minimage", ctypes.c_int64),
        ("kernel_launches", ctypes.c_int64),
        ("kernel_ms", ctypes.c_float),
        ("submit_ms", ctypes.c_float),
        ("n_pairs", ctypes.c_int32),
        ("nbins", ctypes.c_int32),
        ("sm_count", ctypes.c_int32),
        ("variant", ctypes.c_int32),
        ("pair_kernel_ms", ctypes.c_float),
        ("pair_kernel_launches", ctypes.c_int32),
        ("dense", ctypes.c_int32),
        ("frames_far", ctypes.c_int32),
        ("cells", ctypes.c_int32 * 3),
        ("item_shards", ctypes.c_int32),
        ("cell_stencil", ctypes.c_int32 * 3),
        ("cell_build", ctypes.c_int32),
        ("pairs_tested", ctypes.c_double),
    ]


@dataclass
class RDFStats:
    path: int
    frames: int
    pairs_evaluated: float
    pairs_allpairs_equiv: float
    pairs_binned: float
    frames_general_minimage: int
    kernel_launches: int
    kernel_ms: float
    submit_ms: float
    n_pairs: int
    nbins: int
    sm_count: int
    variant: int
    pair_kernel_ms: float
    pair_kernel_launches: int
    dense: int
    frames_far: int = 0
    cells: tuple = (0, 0, 0)
    item_shards: int = 1
    cell_stencil: tuple = (0, 0, 0)
    pairs_tested: float = 0.0
    cell_build: int = 0  # cell path: 1 = the last batch's lists came from the one-kernel build

    @property
    def path_name(self) -> str:
        return "celllist" if self.path == PATH_CELLLIST else "allpairs"


_lib = None


def load() -> ctypes.CDLL:
    """dlopen libmds_rdf.so and declare the prototypes.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MdsCudaError(
            f"{LIB_PATH} is missing: build it with `python -m mdsuite_b200._cuda.build` "
            "(nvcc, sm_100a).  mdsuite_b200 has no CPU fallback for the RDF hot path.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    lib.mds_rdf_create.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(_Config)]
    lib.mds_rdf_destroy.argtypes = [vp]
    lib.mds_rdf_destroy.restype = None
    lib.mds_rdf_submit.argtypes = [vp, vp, i64, i32, i32]
    lib.mds_rdf_submit_device.argtypes = [vp, vp, i64, i32, vp]
    lib.mds_rdf_counts.argtypes = [vp, vp, i32]
    lib.mds_rdf_device_counts.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp)]
    lib.mds_rdf_allreduce.argtypes = [vp, vp]
    lib.mds_rdf_join.argtypes = [vp, vp]
    lib.mds_rdf_wait.argtypes = [vp, vp]
    lib.mds_rdf_wait_host_copies.argtypes = [vp]
    lib.mds_rdf_host_ticket.argtypes = [vp]
    lib.mds_rdf_host_ticket.restype = ctypes.c_int64
    lib.mds_rdf_plan_host_batches.argtypes = [ctypes.c_int64, ctypes.c_int32, ctypes.c_int, ctypes.c_int,
                                              ctypes.POINTER(ctypes.c_int64), ctypes.c_int64]
    lib.mds_rdf_plan_host_batches.restype = ctypes.c_int64
    lib.mds_rdf_wait_host_ticket.argtypes = [vp, ctypes.c_int64]
    lib.mds_rdf_synchronize.argtypes = [vp]
    lib.mds_rdf_reset.argtypes = [vp]
    lib.mds_rdf_set_item_shard.argtypes = [vp, ctypes.c_int32, ctypes.c_int32]
    lib.mds_rdf_stats_get.argtypes = [vp, ctypes.POINTER(_Stats)]
    lib.mds_rdf_threshold_table.argtypes = [vp, vp, ctypes.POINTER(ctypes.c_float)]
    lib.mds_rdf_build_threshold_table.argtypes = [ctypes.c_float, i32, vp,
                                                  ctypes.POINTER(ctypes.c_float)]
    lib.mds_bench_shared_atomics.argtypes = [i32, i32, i32, ctypes.POINTER(ctypes.c_double),
                                             ctypes.POINTER(ctypes.c_double),
                                             ctypes.POINTER(ctypes.c_double)]
    lib.mds_last_error.restype = ctypes.c_char_p
    lib.mds_abi_version.restype = ctypes.c_int
    if lib.mds_abi_version() != ABI_VERSION:
        raise MdsCudaError(f"ABI version {lib.mds_abi_version()} != {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def _check(rc: int, what: str):
    if rc != MDS_OK:
        msg = load().mds_last_error()
        raise MdsCudaError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")


def n_species_pairs(n_species: int) -> int:
    return n_species * (n_species + 1) // 2


class RDFHandle:
    """One handle per (thread, GPU): accumulates pair-distance histograms of submitted frames.

    counts : atoms per species in ``args.species`` order (the reference's ``particles_list``,
             calculators/radial_distribution_function.py:691-717)
    box    : ``experiment.box_array`` (cast to float32 like :450)
    parity : reproduce the reference's first-atom quirk (Q1, :637-638) — the default, since the
             acceptance criterion is bit-exact histograms against the reference
    """

    def __init__(self, counts: Sequence[int], nbins: int, cutoff: float, box: Sequence[float],
                 parity: bool = True, device: int = 0, max_frames_in_flight: int = 0,
                 path: Optional[str] = None):
        lib = load()
        self._lib = lib
        self._h = ctypes.c_void_p()
        self.counts_per_species = [int(c) for c in counts]
        self.n_atoms = int(sum(self.counts_per_species))
        self.n_species = len(self.counts_per_species)
        self.n_pairs = n_species_pairs(self.n_species)
        self.nbins = int(nbins)
        self.device = int(device)
        flags = PARITY_SKIP_FIRST if parity else 0
        if path == "allpairs":
            flags |= PATH_ALLPAIRS
        elif path == "celllist":
            flags |= PATH_CELLLIST
        elif path not in (None, "auto"):
            raise ValueError(f"unknown path {path!r}")
        arr = (ctypes.c_int64 * self.n_species)(*self.counts_per_species)
        cfg = _Config()
        cfg.struct_size = ctypes.sizeof(_Config)
        cfg.n_species = self.n_species
        cfg.counts = ctypes.cast(arr, ctypes.POINTER(ctypes.c_int64))
        cfg.nbins = self.nbins
        cfg.cutoff = float(np.float32(cutoff))
        b32 = np.asarray(box, dtype=np.float32)
        if b32.shape != (3,):
            raise ValueError("box must have 3 entries")
        cfg.box = (ctypes.c_float * 3)(*[float(v) for v in b32])
        cfg.flags = flags
        cfg.device = self.device
        cfg.max_frames_in_flight = int(max_frames_in_flight)
        _check(lib.mds_rdf_create(ctypes.byref(self._h), ctypes.byref(cfg)), "mds_rdf_create")
        self._keepalive = []

    # -- lifetime -------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.mds_rdf_destroy(self._h)
            self._h = ctypes.c_void_p()
        self._keepalive = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- data path ------------------------------------------------------------------------
    def _shape_to_frames(self, shape, layout) -> int:
        if len(shape) != 3 or shape[2] != 3:
            raise ValueError(f"positions must have shape (.., .., 3), got {tuple(shape)}")
        atoms_axis = 1 if layout == LAYOUT_FRAME_MAJOR else 0
        if shape[atoms_axis] != self.n_atoms:
            raise ValueError(f"expected {self.n_atoms} atoms on axis {atoms_axis}, got "
                             f"{shape[atoms_axis]}")
        return int(shape[1 - atoms_axis])

    def submit(self, positions, layout: int = LAYOUT_FRAME_MAJOR, pinned: bool = False):
        """Accumulate frames from HOST memory: a C-contiguous float32 NumPy array or a CPU
        torch tensor, shape (F, N, 3) for LAYOUT_FRAME_MAJOR or (N, F, 3) for LAYOUT_ATOM_MAJOR.
        Asynchronous; the array is kept alive until the next synchronising call."""
        if hasattr(positions, "data_ptr"):  # torch tensor
            t = positions
            if t.device.type != "cpu":
                return self.submit_device(t, layout)
            import torch
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("host tensor must be contiguous float32")
            n_frames = self._shape_to_frames(tuple(t.shape), layout)
            ptr, pinned = t.data_ptr(), pinned or t.is_pinned()
            self._keepalive.append(t)
        else:
            a = positions
            if not isinstance(a, np.ndarray) or a.dtype != np.float32 or not a.flags.c_contiguous:
                raise ValueError("positions must be a C-contiguous float32 ndarray")
            n_frames = self._shape_to_frames(a.shape, layout)
            ptr = a.ctypes.data
            self._keepalive.append(a)
        _check(self._lib.mds_rdf_submit(self._h, ptr, n_frames, layout, int(bool(pinned))),
               "mds_rdf_submit")
        return n_frames

    def submit_device(self, tensor, layout: int = LAYOUT_FRAME_MAJOR, stream: Optional[int] = None):
        """Accumulate frames already resident on the handle's GPU (a contiguous float32 CUDA
        torch tensor).  Ordered after ``stream`` (default: torch's current stream)."""
        import torch
        if tensor.device.type != "cuda" or tensor.dtype != torch.float32 or not tensor.is_contiguous():
            raise ValueError("device tensor must be a contiguous float32 CUDA tensor")
        if tensor.device.index not in (None, self.device):
            raise ValueError(f"tensor lives on cuda:{tensor.device.index}, handle on cuda:{self.device}")
        n_frames = self._shape_to_frames(tuple(tensor.shape), layout)
        if stream is None:
            stream = torch.cuda.current_stream(tensor.device).cuda_stream
        self._keepalive.append(tensor)
        _check(self._lib.mds_rdf_submit_device(self._h, tensor.data_ptr(), n_frames, layout,
                                               ctypes.c_void_p(stream)), "mds_rdf_submit_device")
        return n_frames

    def join(self, stream: Optional[int] = None):
        """Make ``stream`` (default: torch's current stream) wait for the submitted work, so
        events recorded on that stream bracket it.  Does not block the host."""
        if stream is None:
            import torch
            stream = torch.cuda.current_stream(self.device).cuda_stream
        _check(self._lib.mds_rdf_join(self._h, ctypes.c_void_p(stream)), "mds_rdf_join")

    def wait(self, stream: Optional[int] = None):
        """Make the library's stream wait for work already enqueued on ``stream`` (default:
        torch's current stream), e.g. an allreduce over :meth:`device_counts`."""
        if stream is None:
            import torch
            stream = torch.cuda.current_stream(self.device).cuda_stream
        _check(self._lib.mds_rdf_wait(self._h, ctypes.c_void_p(stream)), "mds_rdf_wait")

    def wait_host_copies(self):
        """Block until the host buffers passed to :meth:`submit` so far may be overwritten."""
        _check(self._lib.mds_rdf_wait_host_copies(self._h), "mds_rdf_wait_host_copies")
        self._keepalive = []

    def host_ticket(self) -> int:
        """Ticket of the latest page-locked host submit (their count so far)."""
        return int(self._lib.mds_rdf_host_ticket(self._h))

    def wait_host_ticket(self, ticket: int):
        """Block until the copies of host submit number ``ticket`` have left its buffer (later
        submits may still be in flight)."""
        _check(self._lib.mds_rdf_wait_host_ticket(self._h, int(ticket)), "mds_rdf_wait_host_ticket")

    def synchronize(self):
        _check(self._lib.mds_rdf_synchronize(self._h), "mds_rdf_synchronize")
        self._keepalive = []

    def counts(self) -> np.ndarray:
        """(n_pairs, nbins) uint64 histogram accumulated so far (waits for submitted work)."""
        out = np.zeros((self.n_pairs, self.nbins), dtype=np.uint64)
        _check(self._lib.mds_rdf_counts(self._h, out.ctypes.data, 1), "mds_rdf_counts")
        self._keepalive = []
        return out

    def device_counts(self):
        """(device pointer, cudaStream_t) of the uint64 accumulator, for NCCL reductions."""
        ptr, stream = ctypes.c_void_p(), ctypes.c_void_p()
        _check(self._lib.mds_rdf_device_counts(self._h, ctypes.byref(ptr), ctypes.byref(stream)),
               "mds_rdf_device_counts")
        return ptr.value, (stream.value or 0)

    def allreduce(self, comm):
        """Sum the accumulated counts over the ranks of a raw NCCL communicator (an ``ncclComm_t``
        as an integer / ``c_void_p``, or a :class:`mdsuite_b200._cuda.nccl.Communicator`): one
        in-place ncclAllReduce on the library's stream, ordered with the kernels."""
        ptr = getattr(comm, "handle", comm)
        _check(self._lib.mds_rdf_allreduce(self._h, ctypes.c_void_p(ptr)), "mds_rdf_allreduce")

    def reset(self):
        _check(self._lib.mds_rdf_reset(self._h), "mds_rdf_reset")

    def set_item_shard(self, index: int, count: int):
        """Strong scaling within frames: this handle takes every ``count``-th (work item, frame)
        unit, starting at ``index``, of the frames submitted from now on; the sum over the
        ``count`` handles' histograms is the full result.  (0, 1) switches it off."""
        _check(self._lib.mds_rdf_set_item_shard(self._h, int(index), int(count)),
               "mds_rdf_set_item_shard")

    def stats(self) -> RDFStats:
        st = _Stats()
        st.struct_size = ctypes.sizeof(_Stats)
        _check(self._lib.mds_rdf_stats_get(self._h, ctypes.byref(st)), "mds_rdf_stats_get")
        self._keepalive = []
        return RDFStats(st.path, st.frames, st.pairs_evaluated, st.pairs_allpairs_equiv,
                        st.pairs_binned, st.frames_general_minimage, st.kernel_launches,
                        st.kernel_ms, st.submit_ms, st.n_pairs, st.nbins, st.sm_count, st.variant,
                        st.pair_kernel_ms, st.pair_kernel_launches, st.dense, st.frames_far,
                        tuple(st.cells), st.item_shards, tuple(st.cell_stencil), st.pairs_tested,
                        st.cell_build)

    def threshold_table(self):
        """(edges[0..nbins] float32, s_cut): the squared-distance thresholds the kernels use."""
        edges = np.zeros(self.nbins + 1, dtype=np.float32)
        s_cut = ctypes.c_float()
        _check(self._lib.mds_rdf_threshold_table(self._h, edges.ctypes.data, ctypes.byref(s_cut)),
               "mds_rdf_threshold_table")
        return edges, np.float32(s_cut.value)


# ---- handle reuse ------------------------------------------------------------------------------
# Creating and destroying a handle costs ~50 ms (cudaMalloc / cudaFree of the frame buffers, the
# threshold table, the work schedule) — more than the kernels of a few hundred frames.  The
# calculator therefore parks its handle here when a run ends and the next run with the same
# configuration takes it back (after a reset).  One parked handle per device.
_PARKED = {}


def _handle_key(counts, nbins, cutoff, box, parity, device, path):
    return (tuple(int(c) for c in counts), int(nbins), float(np.float32(cutoff)),
            tuple(float(v) for v in np.asarray(box, dtype=np.float32)), bool(parity), int(device),
            path or "auto")


def acquire_handle(counts, nbins, cutoff, box, parity=True, device=0, path=None) -> "RDFHandle":
    """A reset handle for this configuration: the parked one if it matches, else a new one."""
    key = _handle_key(counts, nbins, cutoff, box, parity, device, path)
    parked = _PARKED.pop(int(device), None)
    if parked is not None:
        if parked[0] == key:
            parked[1].reset()
            return parked[1]
        parked[1].close()
    h = RDFHandle(counts, nbins, cutoff, box, parity=parity, device=device, path=path)
    h._park_key = key
    return h


def park_handle(h: "RDFHandle"):
    """Give a handle back for reuse (its streams are synchronised first)."""
    key = getattr(h, "_park_key", None)
    if key is None or not h._h.value:
        h.close()
        return
    h.synchronize()
    old = _PARKED.pop(h.device, None)
    if old is not None:
        old[1].close()
    _PARKED[h.device] = (key, h)


def drop_parked_handles():
    for _key, h in list(_PARKED.values()):
        h.close()
    _PARKED.clear()


def plan_host_batches(n_frames: int, batch_frames: int, compute_idle: bool, cell_path: bool):
    """Batch sizes ``mds_rdf_submit`` cuts a call of ``n_frames`` host frames into (pure host
    arithmetic of the library; no GPU needed)."""
    lib = load()
    n = lib.mds_rdf_plan_host_batches(n_frames, batch_frames, int(compute_idle), int(cell_path), None, 0)
    if n < 0:
        raise ValueError("mds_rdf_plan_host_batches: invalid argument")
    out = (ctypes.c_int64 * max(1, n))()
    lib.mds_rdf_plan_host_batches(n_frames, batch_frames, int(compute_idle), int(cell_path), out, n)
    return [int(out[k]) for k in range(n)]


def build_threshold_table(cutoff: float, nbins: int):
    """Host-only: (edges[0..nbins] float32, s_cut) for float32(cutoff) — needs no GPU."""
    edges = np.zeros(nbins + 1, dtype=np.float32)
    s_cut = ctypes.c_float()
    _check(load().mds_rdf_build_threshold_table(float(np.float32(cutoff)), int(nbins),
                                                edges.ctypes.data, ctypes.byref(s_cut)),
           "mds_rdf_build_threshold_table")
    return edges, np.float32(s_cut.value)


def bench_shared_atomics(device: int = 0, nbins: int = 4096, iters: int = 4096):
    """(conflict-free lanes/clk/SM, random-bin lanes/clk/SM, SM clock MHz during the test)."""
    a, b, c = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    _check(load().mds_bench_shared_atomics(device, nbins, iters, ctypes.byref(a), ctypes.byref(b),
                                           ctypes.byref(c)), "mds_bench_shared_atomics")
    return a.value, b.value, c.value
