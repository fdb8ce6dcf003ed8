"""
A from-scratch reader for the subset of HDF5 that an MDSuite ``database.hdf5`` uses
(SURVEY.md §8f rank 1).  libhdf5 / h5py are not dependencies of this package.

What the reference writes (mdsuite/database/simulation_database.py:449-499 through h5py with
default library bounds): superblock version 0, old-style groups (symbol-table message -> v1
B-tree of symbol-table nodes + local heap), version-1 object headers, one dataset per
``<species>/<property>`` of shape (n_atoms, n_configurations, 3) float32 little-endian,
``chunks=True`` (chunked layout version 3, v1 B-tree chunk index), ``compression="gzip"``
(filter pipeline with deflate; shuffle is honoured if present), resizable along axis 1.

Supported beyond that, because they are cheap and files in the wild have them: superblock
versions 0-1, object-header continuation blocks, layout versions 1-3 (compact, contiguous,
chunked), dataspace versions 1-2, filter-pipeline versions 1-2, fixed-point and IEEE float
datatypes of either byte order, symbol tables spanning several nodes and levels.
NOT supported (raises ``HDF5FormatError``): superblock versions 2-3, version-2 object headers
("OHDR"), new-style groups (link messages / fractal heaps), v2 B-trees, virtual or external
storage, filters other than deflate / shuffle / fletcher32, variable-length or compound types.

PARITY NOTE: there is no HDF5 implementation in the build image to cross-check against.  The
reader is tested on files produced by this package's own minimal writer
(tests/hdf5_writer.py, written from the format specification) and on the one foreign HDF5 file the
image holds (a MATLAB v7.3 file shipped with SciPy's tests).
"""
from __future__ import annotations

import struct
import zlib
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class HDF5FormatError(ValueError):
    """The file uses a feature outside the supported subset, or is not HDF5."""


@dataclass
class _Message:
    type: int
    flags: int
    data: bytes


@dataclass
class ChunkRef:
    offsets: Tuple[int, ...]  # element offsets of the chunk's first element (without the type dim)
    address: int
    nbytes: int
    filter_mask: int


@dataclass
class Dataset:
    """One dataset: shape/dtype/layout and the means to read it."""
    file: "H5File"
    name: str
    shape: Tuple[int, ...]
    maxshape: Tuple[Optional[int], ...]
    dtype: np.dtype
    layout: str  # "chunked" | "contiguous" | "compact"
    chunks: Optional[Tuple[int, ...]] = None
    filters: List[Tuple[int, Tuple[int, ...]]] = field(default_factory=list)  # (id, client data)
    _address: int = UNDEF
    _size: int = 0
    _compact: bytes = b""
    _btree: int = UNDEF
    _index: Optional[Dict[Tuple[int, ...], ChunkRef]] = None

    # ---- chunk index -----------------------------------------------------------------------
    def chunk_index(self) -> Dict[Tuple[int, ...], ChunkRef]:
        if self._index is None:
            self._index = {}
            if self.layout == "chunked" and self._btree != UNDEF:
                for ref in self.file._walk_chunk_btree(self._btree, len(self.shape)):
                    self._index[ref.offsets] = ref
        return self._index

    def _decode_chunk(self, ref: ChunkRef) -> np.ndarray:
        raw = self.file._read(ref.address, ref.nbytes)
        # filters are undone in reverse order of the pipeline; bit i of the mask = filter i skipped
        for i in range(len(self.filters) - 1, -1, -1):
            fid, cdata = self.filters[i]
            if ref.filter_mask & (1 << i):
                continue
            if fid == 1:  # deflate
                raw = zlib.decompress(raw)
            elif fid == 2:  # shuffle: bytes of each element de-interleaved
                size = cdata[0] if cdata else self.dtype.itemsize
                n = len(raw) // size
                raw = np.frombuffer(raw, dtype=np.uint8, count=n * size).reshape(size, n).T.tobytes()
            elif fid == 3:  # fletcher32: trailing 4-byte checksum
                raw = raw[:-4]
            else:
                raise HDF5FormatError(f"dataset {self.name}: filter id {fid} is not supported")
        arr = np.frombuffer(raw, dtype=self.dtype, count=int(np.prod(self.chunks)))
        return arr.reshape(self.chunks)

    # ---- reading ---------------------------------------------------------------------------
    def read(self) -> np.ndarray:
        """The whole dataset as a NumPy array."""
        return self.read_hyperslab(tuple(slice(0, n) for n in self.shape))

    def read_box_into(self, lo: Sequence[int], hi: Sequence[int], out: np.ndarray,
                      out_strides: Sequence[int], threads: int = 0, verify: bool = True) -> bool:
        """The box [lo, hi) of a chunked dataset decoded by the native library on all host cores
        (include/mds_io.h, mds_h5_read_chunks: pread + hand-written inflate + strided placement)
        straight into ``out``: the element at dataset index i lands at flat position
        sum_d (i_d - lo_d) * out_strides[d].  ``out`` must be pre-filled with the fill value.
        Returns False when the dataset is outside what the native path takes (not chunked, rank
        above 4, a filter it does not know) — the caller then uses read_hyperslab."""
        if self.layout != "chunked" or len(self.shape) > 4 or len(self.filters) > 8:
            return False
        if any(fid not in (1, 2, 3) for fid, _ in self.filters):
            return False
        from mdsuite_b200 import _native
        index = self.chunk_index()
        ranges = [range(l // c * c, h, c) for l, h, c in zip(lo, hi, self.chunks)]
        wanted = [ref for offs in np.ndindex(*[len(r) for r in ranges])
                  if (ref := index.get(tuple(r[i] for r, i in zip(ranges, offs)))) is not None]
        base = self.file._base
        _native.h5_read_chunks(
            self.file._fh.fileno(),
            [(base + ref.address, ref.nbytes, ref.filter_mask, ref.offsets) for ref in wanted],
            self.chunks, self.dtype.itemsize, [(fid, cd[0] if cd else 0) for fid, cd in self.filters],
            lo, hi, out, out_strides, n_threads=threads, verify=verify)
        return True

    def read_hyperslab(self, slices: Sequence[slice], threads: int = 8) -> np.ndarray:
        """Axis-aligned box [start, stop) per dimension (unit steps)."""
        lo = [s.start or 0 for s in slices]
        hi = [self.shape[d] if s.stop is None else min(s.stop, self.shape[d])
              for d, s in enumerate(slices)]
        out_shape = tuple(max(0, h - l) for l, h in zip(lo, hi))
        out = np.zeros(out_shape, dtype=self.dtype)  # missing chunks read as the fill value 0
        if 0 in out_shape:
            return out
        if self.layout == "compact":
            full = np.frombuffer(self._compact, dtype=self.dtype,
                                 count=int(np.prod(self.shape))).reshape(self.shape)
            return np.array(full[tuple(slice(l, h) for l, h in zip(lo, hi))])
        if self.layout == "contiguous":
            if self._address == UNDEF:
                return out
            raw = self.file._read(self._address, int(np.prod(self.shape)) * self.dtype.itemsize)
            full = np.frombuffer(raw, dtype=self.dtype).reshape(self.shape)
            return np.array(full[tuple(slice(l, h) for l, h in zip(lo, hi))])
        index = self.chunk_index()
        ranges = [range(l // c * c, h, c) for l, h, c in zip(lo, hi, self.chunks)]
        wanted = [ref for offs in np.ndindex(*[len(r) for r in ranges])
                  if (ref := index.get(tuple(r[i] for r, i in zip(ranges, offs)))) is not None]

        def place(ref: ChunkRef):
            chunk = self._decode_chunk(ref)
            src, dst = [], []
            for d, off in enumerate(ref.offsets):
                a, b = max(lo[d], off), min(hi[d], off + self.chunks[d])
                src.append(slice(a - off, b - off))
                dst.append(slice(a - lo[d], b - lo[d]))
            out[tuple(dst)] = chunk[tuple(src)]

        if threads > 1 and len(wanted) > 1:
            with ThreadPoolExecutor(max_workers=threads) as pool:  # zlib releases the GIL
                list(pool.map(place, wanted))
        else:
            for ref in wanted:
                place(ref)
        return out


class H5File:
    """Read-only view of an HDF5 file of the supported subset."""

    def __init__(self, path: str):
        self.path = path
        self._fh = open(path, "rb")
        self._mm = np.memmap(path, dtype=np.uint8, mode="r")
        self._base = 0
        self._parse_superblock()
        self._groups: Dict[str, Dict[str, int]] = {}

    def close(self):
        self._mm = None
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- low level -------------------------------------------------------------------------
    def _read(self, address: int, n: int) -> bytes:
        a = self._base + address
        if a < 0 or a + n > self._mm.shape[0]:
            raise HDF5FormatError(f"read of {n} bytes at {address} is outside the file")
        return self._mm[a:a + n].tobytes()

    def _uint(self, buf: bytes, pos: int, size: int) -> int:
        return int.from_bytes(buf[pos:pos + size], "little")

    def _parse_superblock(self):
        size = self._mm.shape[0]
        start = 0
        while True:  # the signature may sit at 0, 512, 1024, ... (user block)
            if start + 8 > size:
                raise HDF5FormatError(f"{self.path}: no HDF5 signature found")
            if self._mm[start:start + 8].tobytes() == SIGNATURE:
                break
            start = 512 if start == 0 else start * 2
        sb = self._mm[start:start + 128].tobytes()
        version = sb[8]
        if version not in (0, 1):
            raise HDF5FormatError(f"superblock version {version} is not supported (only 0 and 1: "
                                  f"files written with libver='earliest', the h5py default)")
        self.size_of_offsets, self.size_of_lengths = sb[13], sb[14]
        O = self.size_of_offsets
        pos = 24 if version == 0 else 28
        base = self._uint(sb, pos, O)
        self._base = start if base == 0 else base  # addresses are relative to the base address
        pos += 4 * O  # base, free-space info, end of file, driver info
        # root group symbol table entry
        self._root_header = self._uint(sb, pos + O, O)
        cache_type = self._uint(sb, pos + 2 * O, 4)
        self._root_cache = None
        if cache_type == 1:
            scratch = pos + 2 * O + 8
            self._root_cache = (self._uint(sb, scratch, O), self._uint(sb, scratch + O, O))

    # ---- object headers ----------------------------------------------------------------------
    def _messages(self, address: int) -> List[_Message]:
        head = self._read(address, 16)
        if head[:4] == b"OHDR":
            raise HDF5FormatError("version-2 object headers are not supported (file written with "
                                  "libver='latest')")
        if head[0] != 1:
            raise HDF5FormatError(f"object header version {head[0]} at {address}")
        n_msgs = self._uint(head, 2, 2)
        first_size = self._uint(head, 8, 4)
        blocks = [(address + 16, first_size)]
        out: List[_Message] = []
        while blocks and len(out) < n_msgs:
            addr, length = blocks.pop(0)
            buf = self._read(addr, length)
            pos = 0
            while pos + 8 <= length and len(out) < n_msgs:
                mtype, msize, mflags = struct.unpack_from("<HHB", buf, pos)
                data = buf[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x0010:  # continuation
                    O, L = self.size_of_offsets, self.size_of_lengths
                    blocks.append((self._uint(data, 0, O), self._uint(data, O, L)))
                out.append(_Message(mtype, mflags, data))
        return out

    # ---- groups --------------------------------------------------------------------------------
    def _heap_string(self, heap_address: int, offset: int) -> str:
        O, L = self.size_of_offsets, self.size_of_lengths
        head = self._read(heap_address, 8 + 2 * L + O)
        if head[:4] != b"HEAP":
            raise HDF5FormatError(f"no local heap at {heap_address}")
        seg_size = self._uint(head, 8, L)
        seg_addr = self._uint(head, 8 + 2 * L, O)
        seg = self._read(seg_addr, seg_size)
        end = seg.index(b"\0", offset)
        return seg[offset:end].decode("utf-8")

    def _walk_group_btree(self, address: int, heap: int) -> Iterator[Tuple[str, int]]:
        O, L = self.size_of_offsets, self.size_of_lengths
        head = self._read(address, 8 + 2 * O)
        if head[:4] == b"SNOD":
            n = self._uint(head, 6, 2)
            entry = 2 * O + 24
            buf = self._read(address + 8, n * entry)
            for k in range(n):
                name_off = self._uint(buf, k * entry, O)
                header = self._uint(buf, k * entry + O, O)
                yield self._heap_string(heap, name_off), header
            return
        if head[:4] != b"TREE" or head[4] != 0:
            raise HDF5FormatError(f"expected a group B-tree node at {address}")
        used = self._uint(head, 6, 2)
        body = self._read(address + 8 + 2 * O, (used + 1) * L + used * O)
        pos = L  # key 0
        for _ in range(used):
            child = self._uint(body, pos, O)
            pos += O + L  # child, next key
            yield from self._walk_group_btree(child, heap)

    def _group_entries(self, header_address: int, cache=None) -> Dict[str, int]:
        btree = heap = None
        if cache is not None:
            btree, heap = cache
        else:
            for m in self._messages(header_address):
                if m.type == 0x0011:
                    O = self.size_of_offsets
                    btree, heap = self._uint(m.data, 0, O), self._uint(m.data, O, O)
                elif m.type in (0x0002, 0x0006):
                    raise HDF5FormatError("new-style groups (link messages) are not supported")
        if btree is None:
            return {}
        return dict(self._walk_group_btree(btree, heap))

    def _resolve(self, path: str) -> int:
        """Object header address of /a/b/c."""
        address, cache = self._root_header, self._root_cache
        walked = ""
        for part in [p for p in path.split("/") if p]:
            key = walked or "/"
            if key not in self._groups:
                self._groups[key] = self._group_entries(address, cache)
            entries = self._groups[key]
            if part not in entries:
                raise KeyError(f"{path!r}: no object {part!r} in group {key!r}")
            address, cache = entries[part], None
            walked += "/" + part
        return address

    def listdir(self, path: str = "/") -> List[str]:
        address = self._resolve(path)
        cache = self._root_cache if path.strip("/") == "" else None
        key = "/" + "/".join(p for p in path.split("/") if p) if path.strip("/") else "/"
        if key not in self._groups:
            self._groups[key] = self._group_entries(address, cache)
        return list(self._groups[key])

    def __contains__(self, path: str) -> bool:
        try:
            self._resolve(path)
            return True
        except KeyError:
            return False

    def is_dataset(self, path: str) -> bool:
        return any(m.type == 0x0008 for m in self._messages(self._resolve(path)))

    # ---- datasets ------------------------------------------------------------------------------
    def _parse_dataspace(self, d: bytes):
        L = self.size_of_lengths
        version, rank, flags = d[0], d[1], d[2]
        if version == 1:
            pos = 8
        elif version == 2:
            pos = 4
        else:
            raise HDF5FormatError(f"dataspace version {version}")
        dims = tuple(self._uint(d, pos + k * L, L) for k in range(rank))
        pos += rank * L
        maxdims = dims
        if flags & 1:
            raw = [self._uint(d, pos + k * L, L) for k in range(rank)]
            maxdims = tuple(None if v == (1 << (8 * L)) - 1 else v for v in raw)
        return dims, maxdims

    def _parse_datatype(self, d: bytes) -> np.dtype:
        cls, version = d[0] & 0x0F, d[0] >> 4
        bits0 = d[1]
        size = self._uint(d, 4, 4)
        order = ">" if bits0 & 1 else "<"
        if cls == 0:  # fixed point
            signed = bool(bits0 & 0x08)
            return np.dtype(f"{order}{'i' if signed else 'u'}{size}")
        if cls == 1:  # floating point: accept the IEEE layouts only
            if size not in (2, 4, 8):
                raise HDF5FormatError(f"float of {size} bytes")
            return np.dtype(f"{order}f{size}")
        raise HDF5FormatError(f"datatype class {cls} (version {version}) is not supported")

    def _parse_filters(self, d: bytes) -> List[Tuple[int, Tuple[int, ...]]]:
        version, n = d[0], d[1]
        pos = 8 if version == 1 else 2
        out = []
        for _ in range(n):
            fid = self._uint(d, pos, 2)
            pos += 2
            name_len = 0
            if version == 1 or fid >= 256:
                name_len = self._uint(d, pos, 2)
                pos += 2
            pos += 2  # flags
            n_cd = self._uint(d, pos, 2)
            pos += 2
            if version == 1:
                name_len = (name_len + 7) & ~7
            pos += name_len
            cdata = tuple(self._uint(d, pos + 4 * k, 4) for k in range(n_cd))
            pos += 4 * n_cd
            if version == 1 and n_cd % 2:
                pos += 4
            out.append((fid, cdata))
        return out

    def dataset(self, path: str) -> Dataset:
        msgs = self._messages(self._resolve(path))
        by_type = {m.type: m for m in msgs}
        if 0x0008 not in by_type:
            raise KeyError(f"{path!r} is not a dataset")
        dims, maxdims = self._parse_dataspace(by_type[0x0001].data)
        dtype = self._parse_datatype(by_type[0x0003].data)
        filters = self._parse_filters(by_type[0x000B].data) if 0x000B in by_type else []
        ds = Dataset(self, path, dims, maxdims, dtype, "contiguous", filters=filters)
        lay = by_type[0x0008].data
        O, L = self.size_of_offsets, self.size_of_lengths
        version = lay[0]
        if version == 3:
            cls = lay[1]
            if cls == 0:
                size = self._uint(lay, 2, 2)
                ds.layout, ds._compact = "compact", lay[4:4 + size]
            elif cls == 1:
                ds.layout, ds._address, ds._size = "contiguous", self._uint(lay, 2, O), self._uint(lay, 2 + O, L)
            elif cls == 2:
                ndim = lay[2]
                ds.layout, ds._btree = "chunked", self._uint(lay, 3, O)
                cd = [self._uint(lay, 3 + O + 4 * k, 4) for k in range(ndim)]
                ds.chunks = tuple(cd[:-1])  # the last entry is the element size
            else:
                raise HDF5FormatError(f"layout class {cls}")
        elif version in (1, 2):
            ndim, cls = lay[1], lay[2]
            pos = 8
            if cls != 0:
                addr = self._uint(lay, pos, O)
                pos += O
            cd = [self._uint(lay, pos + 4 * k, 4) for k in range(ndim)]
            pos += 4 * ndim
            if cls == 0:
                size = self._uint(lay, pos, 4)
                ds.layout, ds._compact = "compact", lay[pos + 4:pos + 4 + size]
            elif cls == 1:
                ds.layout, ds._address = "contiguous", addr
            else:
                ds.layout, ds._btree, ds.chunks = "chunked", addr, tuple(cd[:-1])
        else:
            raise HDF5FormatError(f"data layout version {version} is not supported")
        if ds.layout == "chunked" and len(ds.chunks) != len(dims):
            raise HDF5FormatError(f"{path}: chunk rank {len(ds.chunks)} != dataset rank {len(dims)}")
        return ds

    def _walk_chunk_btree(self, address: int, rank: int) -> Iterator[ChunkRef]:
        O = self.size_of_offsets
        head = self._read(address, 8 + 2 * O)
        if head[:4] != b"TREE" or head[4] != 1:
            raise HDF5FormatError(f"expected a chunk B-tree node at {address}")
        level, used = head[5], self._uint(head, 6, 2)
        key_size = 8 + 8 * (rank + 1)
        body = self._read(address + 8 + 2 * O, used * (key_size + O) + key_size)
        pos = 0
        for _ in range(used):
            nbytes, mask = struct.unpack_from("<II", body, pos)
            offsets = struct.unpack_from(f"<{rank + 1}Q", body, pos + 8)
            child = self._uint(body, pos + key_size, O)
            pos += key_size + O
            if level == 0:
                yield ChunkRef(tuple(offsets[:rank]), child, nbytes, mask)
            else:
                yield from self._walk_chunk_btree(child, rank)

    def walk(self, path: str = "/") -> Iterator[str]:
        """Every dataset path below ``path``."""
        for name in self.listdir(path):
            child = (path.rstrip("/") + "/" + name)
            if self.is_dataset(child):
                yield child.lstrip("/")
            else:
                yield from self.walk(child)
