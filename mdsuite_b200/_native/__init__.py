"""
mdsuite_b200._native — ctypes binding of the host-side native library (include/mds_io.h).

Native replacements for the pure-Python steps on the way into the trajectory store
(SURVEY.md §8f): the LAMMPS text-dump reader.  Built in-tree with ``make -C
mdsuite_b200/_native/csrc`` (g++, no CUDA); ``load()`` builds it on first use if it is missing.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmds_io.so")
ABI_VERSION = 3

EXPORTED_SYMBOLS = (
    "mds_dump_open", "mds_xyz_open", "mds_dump_header_line", "mds_dump_close", "mds_dump_info_get", "mds_dump_column_name",
    "mds_dump_first_frame_labels", "mds_dump_read_columns", "mds_dump_read_columns_mapped",
    "mds_parse_float32", "mds_h5_read_chunks", "mds_inflate",
    "mds_io_last_error", "mds_io_abi_version",
)


class MdsIOError(RuntimeError):
    """Raised for every non-zero status of the native I/O library."""


class _DumpInfo(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32),
        ("n_columns", ctypes.c_int32),
        ("n_atoms", ctypes.c_int64),
        ("n_configurations", ctypes.c_int64),
        ("file_bytes", ctypes.c_int64),
        ("timestep0", ctypes.c_int64),
        ("timestep1", ctypes.c_int64),
        ("box_lo", ctypes.c_double * 3),
        ("box_hi", ctypes.c_double * 3),
        ("id_column", ctypes.c_int32),
        ("species_column", ctypes.c_int32),
    ]


class H5Chunk(ctypes.Structure):
    """mds_h5_chunk (include/mds_io.h)."""
    _fields_ = [("address", ctypes.c_int64), ("nbytes", ctypes.c_int64),
                ("offsets", ctypes.c_int64 * 4), ("filter_mask", ctypes.c_uint32),
                ("reserved", ctypes.c_uint32)]


class H5Layout(ctypes.Structure):
    """mds_h5_layout (include/mds_io.h)."""
    _fields_ = [("struct_size", ctypes.c_uint32), ("rank", ctypes.c_int32),
                ("elem_size", ctypes.c_int32), ("n_filters", ctypes.c_int32),
                ("filter_ids", ctypes.c_int32 * 8), ("filter_cd0", ctypes.c_int32 * 8),
                ("verify_checksums", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("chunk_dims", ctypes.c_int64 * 4), ("box_lo", ctypes.c_int64 * 4),
                ("box_hi", ctypes.c_int64 * 4), ("out_strides", ctypes.c_int64 * 4)]


def build(force: bool = False) -> str:
    cmd = ["make", "-s", "-C", os.path.join(_HERE, "csrc")]
    if force:
        cmd.append("-B")
    subprocess.check_call(cmd)
    if not os.path.exists(LIB_PATH):
        raise MdsIOError("build finished but libmds_io.so is missing")
    return LIB_PATH


_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    srcs = [os.path.join(_HERE, "csrc", f) for f in ("lammps_reader.cpp", "h5_chunks.cpp")]
    if not os.path.exists(LIB_PATH) or any(os.path.exists(src) and
                                           os.path.getmtime(LIB_PATH) < os.path.getmtime(src)
                                           for src in srcs):
        build()
    lib = ctypes.CDLL(LIB_PATH)
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
    lib.mds_dump_open.argtypes = [ctypes.POINTER(vp), ctypes.c_char_p, ctypes.c_int]
    lib.mds_xyz_open.argtypes = [ctypes.POINTER(vp), ctypes.c_char_p, ctypes.c_int]
    lib.mds_dump_header_line.argtypes = [vp, i64, ctypes.c_int, ctypes.c_char_p, i64]
    lib.mds_dump_header_line.restype = i64
    lib.mds_dump_close.argtypes = [vp]
    lib.mds_dump_close.restype = None
    lib.mds_dump_info_get.argtypes = [vp, ctypes.POINTER(_DumpInfo)]
    lib.mds_dump_column_name.argtypes = [vp, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
    lib.mds_dump_first_frame_labels.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int]
    lib.mds_dump_read_columns.argtypes = [vp, i64, i64, vp, i32, ctypes.c_int, vp]
    lib.mds_dump_read_columns_mapped.argtypes = [vp, i64, i64, vp, i32, ctypes.c_int, vp, vp]
    lib.mds_parse_float32.argtypes = [vp, i64, i32, vp]
    lib.mds_h5_read_chunks.argtypes = [ctypes.c_int, ctypes.POINTER(H5Layout), vp, i64, vp, ctypes.c_int]
    lib.mds_inflate.argtypes = [vp, i64, vp, i64, ctypes.POINTER(i64), ctypes.c_int]
    lib.mds_io_last_error.restype = ctypes.c_char_p
    lib.mds_io_abi_version.restype = ctypes.c_int
    if lib.mds_io_abi_version() != ABI_VERSION:
        raise MdsIOError(f"libmds_io ABI {lib.mds_io_abi_version()} != {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def _check(rc: int, what: str):
    if rc != 0:
        msg = load().mds_io_last_error()
        raise MdsIOError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")


@dataclass
class DumpInfo:
    n_atoms: int
    n_configurations: int
    columns: List[str]
    box_lo: List[float]
    box_hi: List[float]
    timestep0: int
    timestep1: int
    id_column: int
    species_column: int
    file_bytes: int


class LammpsDump:
    """An mmap'ed LAMMPS text dump with an index of its configurations."""

    LABEL_STRIDE = 32
    _OPEN = "mds_dump_open"

    def __init__(self, path: str, n_threads: int = 0):
        self._lib = load()
        self._h = ctypes.c_void_p()
        _check(getattr(self._lib, self._OPEN)(ctypes.byref(self._h), os.fsencode(path),
                                              int(n_threads)), self._OPEN)
        raw = _DumpInfo()
        raw.struct_size = ctypes.sizeof(_DumpInfo)
        _check(self._lib.mds_dump_info_get(self._h, ctypes.byref(raw)), "mds_dump_info_get")
        cols = []
        buf = ctypes.create_string_buffer(64)
        for i in range(raw.n_columns):
            _check(self._lib.mds_dump_column_name(self._h, i, buf, 64), "mds_dump_column_name")
            cols.append(buf.value.decode())
        self.info = DumpInfo(int(raw.n_atoms), int(raw.n_configurations), cols, list(raw.box_lo),
                             list(raw.box_hi), int(raw.timestep0), int(raw.timestep1),
                             int(raw.id_column), int(raw.species_column), int(raw.file_bytes))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.mds_dump_close(self._h)
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

    def header_line(self, frame: int, line: int) -> str:
        """Header line ``line`` (0-based) of configuration ``frame``, without the line end."""
        size = 4096
        while True:
            buf = ctypes.create_string_buffer(size)
            n = self._lib.mds_dump_header_line(self._h, int(frame), int(line), buf, size)
            if n < 0:
                _check(int(n), "mds_dump_header_line")
            if n < size:
                return buf.value.decode()
            size = int(n) + 1

    def first_frame_labels(self, column: int, sort_by_id: bool = True) -> List[str]:
        n = self.info.n_atoms
        buf = ctypes.create_string_buffer(max(1, n * self.LABEL_STRIDE))
        _check(self._lib.mds_dump_first_frame_labels(self._h, int(column), int(sort_by_id), buf,
                                                     self.LABEL_STRIDE),
               "mds_dump_first_frame_labels")
        raw = buf.raw
        return [raw[k * self.LABEL_STRIDE:(k + 1) * self.LABEL_STRIDE].split(b"\0", 1)[0].decode()
                for k in range(n)]

    def read_columns(self, frame0: int, n_frames: int, columns: Sequence[int],
                     sort_by_id: bool = True, out: np.ndarray = None,
                     dest: np.ndarray = None) -> np.ndarray:
        """(n_frames, n_atoms, len(columns)) float32 = float32(float64(token)), atoms in id order;
        with ``dest`` the atom of rank k lands in row dest[k]."""
        cols = np.ascontiguousarray(columns, dtype=np.int32)
        dest_ptr = None
        if dest is not None:
            dest = np.ascontiguousarray(dest, dtype=np.int32)
            if dest.shape != (self.info.n_atoms,):
                raise ValueError("dest must have one entry per atom")
            dest_ptr = dest.ctypes.data
        shape = (int(n_frames), self.info.n_atoms, len(cols))
        if out is None:
            out = np.empty(shape, dtype=np.float32)
        elif out.shape != shape or out.dtype != np.float32 or not out.flags.c_contiguous:
            raise ValueError(f"out must be a C-contiguous float32 array of shape {shape}")
        _check(self._lib.mds_dump_read_columns_mapped(self._h, int(frame0), int(n_frames),
                                                      cols.ctypes.data, len(cols), int(sort_by_id),
                                                      dest_ptr, out.ctypes.data),
               "mds_dump_read_columns")
        return out


class XyzDump(LammpsDump):
    """An mmap'ed extended XYZ file; same calls as LammpsDump, atoms always in file order
    (``sort_by_id=False``), the key=value line of a configuration is ``header_line(frame, 1)``."""

    _OPEN = "mds_xyz_open"


def parse_float32(tokens: Sequence[str]) -> np.ndarray:
    """float32((double) token) for every string — the library's number conversion on its own."""
    enc = [t.encode() for t in tokens]
    stride = max([len(t) for t in enc] + [1]) + 1
    buf = b"".join(t.ljust(stride, b"\0") for t in enc)
    out = np.empty(len(enc), dtype=np.float32)
    _check(load().mds_parse_float32(buf, len(enc), stride, out.ctypes.data), "mds_parse_float32")
    return out


def inflate(data: bytes, size: int, verify: bool = True) -> bytes:
    """The library's own inflate on one zlib stream of known decompressed size (tests, audits)."""
    out = ctypes.create_string_buffer(max(1, size))
    n = ctypes.c_int64()
    _check(load().mds_inflate(data, len(data), out, size, ctypes.byref(n), int(verify)), "mds_inflate")
    return out.raw[:n.value]


def h5_read_chunks(fd: int, chunks, chunk_dims, elem_size: int, filters, box_lo, box_hi,
                   out: np.ndarray, out_strides, n_threads: int = 0, verify: bool = True):
    """Decode HDF5 chunks and place their part of the box [box_lo, box_hi) into ``out``.

    chunks: iterable of (address, nbytes, filter_mask, offsets); filters: [(id, first client value
    or 0), ...] in pipeline order; out_strides: element strides of ``out`` per dataset dimension."""
    lib = load()
    rank = len(chunk_dims)
    lay = H5Layout()
    lay.struct_size = ctypes.sizeof(H5Layout)
    lay.rank, lay.elem_size, lay.n_filters = rank, int(elem_size), len(filters)
    for i, (fid, cd0) in enumerate(filters):
        lay.filter_ids[i], lay.filter_cd0[i] = int(fid), int(cd0)
    lay.verify_checksums = int(verify)
    for d in range(rank):
        lay.chunk_dims[d], lay.box_lo[d], lay.box_hi[d] = int(chunk_dims[d]), int(box_lo[d]), int(box_hi[d])
        lay.out_strides[d] = int(out_strides[d])
    chunks = list(chunks)
    arr = (H5Chunk * max(1, len(chunks)))()
    for k, (address, nbytes, mask, offsets) in enumerate(chunks):
        arr[k].address, arr[k].nbytes, arr[k].filter_mask = int(address), int(nbytes), int(mask)
        for d in range(rank):
            arr[k].offsets[d] = int(offsets[d])
    _check(lib.mds_h5_read_chunks(int(fd), ctypes.byref(lay), arr, len(chunks), out.ctypes.data,
                                  int(n_threads)), "mds_h5_read_chunks")
    return out
