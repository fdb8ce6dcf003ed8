"""TEST INFRASTRUCTURE ONLY — a minimal HDF5 writer producing the structures h5py emits for an
MDSuite ``database.hdf5`` with default library bounds (superblock v0, symbol-table groups, v1
object headers, chunked v3 layout with a v1 B-tree chunk index, deflate / shuffle filters,
resizable dataspaces), written from the HDF5 File Format Specification v1.1/2.0.  It exists so that
mdsuite_b200/database/hdf5_reader.py can be exercised without libhdf5, which the image lacks."""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


class H5Writer:
    def __init__(self, user_block: int = 0, leaf_k: int = 4, chunk_k: int = 32):
        self.buf = bytearray(b"\0" * user_block)
        self.base = user_block
        self.leaf_k, self.chunk_k = leaf_k, chunk_k
        self.buf += b"\0" * 96  # superblock v0 with 8-byte offsets is 96 bytes: 24 + 4*8 + 40
        self.tree = {}  # nested dict: name -> dict (group) | ("dataset", kwargs)

    # ---- allocation ------------------------------------------------------------------------
    def _alloc(self, data: bytes) -> int:
        self.buf += b"\0" * (-len(self.buf) % 8)
        addr = len(self.buf) - self.base
        self.buf += data
        return addr

    # ---- public API -------------------------------------------------------------------------
    def create_dataset(self, path: str, data: np.ndarray, chunks=None, maxshape=None,
                       gzip: int = 4, shuffle: bool = False, layout: str = "chunked",
                       continuation: bool = False, filter_version: int = 1):
        node = self.tree
        parts = [p for p in path.split("/") if p]
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = ("dataset", dict(data=np.ascontiguousarray(data), chunks=chunks,
                                           maxshape=maxshape, gzip=gzip, shuffle=shuffle,
                                           layout=layout, continuation=continuation,
                                           filter_version=filter_version))

    def create_group(self, path: str):
        node = self.tree
        for p in [q for q in path.split("/") if q]:
            node = node.setdefault(p, {})

    # ---- object headers ---------------------------------------------------------------------
    @staticmethod
    def _message(mtype: int, data: bytes, flags: int = 0) -> bytes:
        data = _pad8(data)
        return struct.pack("<HHB3x", mtype, len(data), flags) + data

    def _object_header(self, messages, continuation=False) -> int:
        if continuation and len(messages) > 2:
            # first block: first message + a continuation message; the rest lives elsewhere
            tail = b"".join(messages[1:])
            tail_addr = self._alloc(tail)
            cont = self._message(0x0010, struct.pack("<QQ", tail_addr, len(tail)))
            body = messages[0] + cont
            n = len(messages) + 1
        else:
            body, n = b"".join(messages), len(messages)
        head = struct.pack("<BxHII4x", 1, n, 1, len(body))
        return self._alloc(head + body)

    def _dataspace(self, shape, maxshape) -> bytes:
        rank = len(shape)
        flags = 1 if maxshape is not None else 0
        out = struct.pack("<BBB5x", 1, rank, flags)
        out += b"".join(struct.pack("<Q", int(d)) for d in shape)
        if maxshape is not None:
            out += b"".join(struct.pack("<Q", UNDEF if m is None else int(m)) for m in maxshape)
        return out

    @staticmethod
    def _datatype(dtype: np.dtype) -> bytes:
        dtype = np.dtype(dtype)
        big = 1 if dtype.byteorder == ">" else 0
        if dtype.kind == "f":
            size = dtype.itemsize
            exp_bits, mant_bits, bias = {4: (8, 23, 127), 8: (11, 52, 1023), 2: (5, 10, 15)}[size]
            bits = bytes([0x20 | big, size * 8 - 1, 0])
            props = struct.pack("<HHBBBBI", 0, size * 8, mant_bits, exp_bits, 0, mant_bits, bias)
            return bytes([0x11]) + bits + struct.pack("<I", size) + props
        if dtype.kind in "iu":
            bits = bytes([(0x08 if dtype.kind == "i" else 0) | big, 0, 0])
            return bytes([0x10]) + bits + struct.pack("<I", dtype.itemsize) + struct.pack(
                "<HH", 0, dtype.itemsize * 8)
        raise ValueError(dtype)

    @staticmethod
    def _filters(gzip, shuffle, itemsize, version) -> bytes:
        flt = []
        if shuffle:
            flt.append((2, b"shuffle\0", (itemsize,)))
        if gzip is not None:
            flt.append((1, b"deflate\0", (gzip,)))
        if not flt:
            return b""
        if version == 1:
            out = struct.pack("<BB6x", 1, len(flt))
            for fid, name, cd in flt:
                name = _pad8(name)
                out += struct.pack("<HHHH", fid, len(name), 1, len(cd)) + name
                out += b"".join(struct.pack("<I", v) for v in cd)
                if len(cd) % 2:
                    out += b"\0" * 4
            return out
        out = struct.pack("<BB", 2, len(flt))
        for fid, _name, cd in flt:
            out += struct.pack("<HHH", fid, 1, len(cd)) + b"".join(struct.pack("<I", v) for v in cd)
        return out

    # ---- chunk storage -----------------------------------------------------------------------
    def _write_chunks(self, data, chunks, gzip, shuffle):
        rank = data.ndim
        grid = [range(0, data.shape[d], chunks[d]) for d in range(rank)]
        refs = []
        for idx in np.ndindex(*[len(g) for g in grid]):
            off = [g[i] for g, i in zip(grid, idx)]
            block = np.zeros(chunks, dtype=data.dtype)  # edge chunks are stored full size
            src = tuple(slice(o, min(o + c, n)) for o, c, n in zip(off, chunks, data.shape))
            block[tuple(slice(0, s.stop - s.start) for s in src)] = data[src]
            raw = block.tobytes()
            if shuffle:
                size = data.dtype.itemsize
                raw = np.frombuffer(raw, np.uint8).reshape(-1, size).T.tobytes()
            if gzip is not None:
                raw = zlib.compress(raw, gzip)
            refs.append((tuple(off), self._alloc(raw), len(raw)))
        return refs

    def _chunk_btree(self, refs, rank, shape) -> int:
        key = lambda nbytes, off: struct.pack("<II", nbytes, 0) + b"".join(  # noqa: E731
            struct.pack("<Q", o) for o in list(off) + [0])
        end_key = key(0, shape)  # the last key names the first chunk beyond the data
        cap = 2 * self.chunk_k

        def node(level, entries, last):
            body = b"".join(k + struct.pack("<Q", child) for k, child in entries) + last
            return self._alloc(b"TREE" + struct.pack("<BBHQQ", 1, level, len(entries), UNDEF, UNDEF) + body)

        level, items = 0, [(key(n, off), addr) for off, addr, n in refs]
        while True:
            groups = [items[i:i + cap] for i in range(0, max(len(items), 1), cap)]
            nodes = []
            for gi, g in enumerate(groups):
                last = groups[gi + 1][0][0] if gi + 1 < len(groups) else end_key
                nodes.append((g[0][0] if g else end_key, node(level, g, last)))
            if len(nodes) == 1:
                return nodes[0][1]
            items, level = nodes, level + 1

    def _write_dataset(self, spec) -> int:
        data = spec["data"]
        msgs = [self._message(0x0001, self._dataspace(data.shape, spec["maxshape"])),
                self._message(0x0003, self._datatype(data.dtype), flags=1)]
        if spec["layout"] == "chunked":
            chunks = tuple(spec["chunks"] or data.shape)
            refs = self._write_chunks(data, chunks, spec["gzip"], spec["shuffle"])
            btree = self._chunk_btree(refs, data.ndim, data.shape)
            flt = self._filters(spec["gzip"], spec["shuffle"], data.dtype.itemsize, spec["filter_version"])
            if flt:
                msgs.append(self._message(0x000B, flt, flags=1))
            lay = struct.pack("<BBBQ", 3, 2, data.ndim + 1, btree) + b"".join(
                struct.pack("<I", c) for c in list(chunks) + [data.dtype.itemsize])
        elif spec["layout"] == "contiguous":
            addr = self._alloc(data.tobytes())
            lay = struct.pack("<BBQQ", 3, 1, addr, data.nbytes)
        else:
            raw = data.tobytes()
            lay = struct.pack("<BBH", 3, 0, len(raw)) + raw
        msgs.append(self._message(0x0008, lay))
        msgs.append(self._message(0x0000, b"\0" * 16))  # a null message, as libhdf5 leaves them
        return self._object_header(msgs, spec["continuation"])

    # ---- groups ---------------------------------------------------------------------------------
    def _write_group(self, entries: dict):
        """-> (object header address, btree address, heap address)"""
        children = {}
        for name, node in entries.items():
            if isinstance(node, tuple):
                children[name] = self._write_dataset(node[1])
            else:
                children[name] = self._write_group(node)[0]
        names = sorted(children, key=lambda s: s.encode())
        heap_data = bytearray(b"\0" * 8)  # offset 0: the empty name
        offsets = {}
        for name in names:
            offsets[name] = len(heap_data)
            heap_data += _pad8(name.encode() + b"\0")
        heap_data += b"\0" * 16
        seg = self._alloc(bytes(heap_data))
        heap = self._alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, seg))
        cap = 2 * self.leaf_k
        snods = []
        for i in range(0, max(len(names), 1), cap):
            part = names[i:i + cap]
            body = b"".join(struct.pack("<QQII16x", offsets[n], children[n], 0, 0) for n in part)
            body += b"\0" * (40 * (cap - len(part)))
            addr = self._alloc(b"SNOD" + struct.pack("<BxH", 1, len(part)) + body)
            snods.append((offsets[part[-1]] if part else 0, addr))
        body = struct.pack("<Q", 0)
        for last_name_off, addr in snods:
            body += struct.pack("<QQ", addr, last_name_off)
        btree = self._alloc(b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF) + body)
        header = self._object_header([self._message(0x0011, struct.pack("<QQ", btree, heap))])
        return header, btree, heap

    def tobytes(self) -> bytes:
        header, btree, heap = self._write_group(self.tree)
        eof = len(self.buf) - self.base
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBxBBBxHHI", 0, 0, 0, 0, 8, 8, self.leaf_k, 16, 0)
        sb += struct.pack("<QQQQ", self.base, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, header, 1, 0) + struct.pack("<QQ", btree, heap)
        assert len(sb) == 96
        self.buf[self.base:self.base + 96] = sb
        return bytes(self.buf)

    def save(self, path: str):
        with open(path, "wb") as fh:
            fh.write(self.tobytes())
