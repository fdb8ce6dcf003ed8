"""The native library's hand-written inflate (mdsuite_b200/_native/csrc/h5_chunks.cpp, the decoder
behind the HDF5 chunk feed) against Python's zlib: every block type, every compression level and
strategy, degenerate inputs, and damaged streams.  Byte-exact, tolerance zero."""
import zlib

import numpy as np
import pytest

from mdsuite_b200 import _native


def _corpus():
    rng = np.random.default_rng(7)
    floats = (rng.random((4000, 3)) * 50).astype(np.float32)
    lattice = (np.stack(np.meshgrid(*[np.arange(12)] * 3, indexing="ij"), -1).reshape(-1, 3) * 2.5
               + rng.normal(0, 0.1, (1728, 3))).astype(np.float32)
    return {
        "empty": b"",
        "one": b"x",
        "zeros": bytes(100_000),
        "run_then_noise": b"ab" * 40_000 + rng.bytes(5000),
        "noise": rng.bytes(70_000),                 # stored blocks at every level
        "text": (b"ITEM: TIMESTEP\n0\nITEM: NUMBER OF ATOMS\n1000\n" * 3000),
        "float32_positions": floats.tobytes(),
        "atom_major_chunk": np.ascontiguousarray(
            np.repeat(lattice[:432, None, :], 32, axis=1)
            + rng.normal(0, 0.3, (432, 32, 3)).astype(np.float32)).tobytes(),
        "long_matches": (bytes(range(256)) * 5 + b"\x00" * 600) * 300,   # length-258 matches, far distances
        "dist_32k": rng.bytes(32768) * 3,           # matches at the maximum distance
        "small_alphabet": bytes(rng.integers(0, 3, 50_000, dtype=np.uint8)),
    }


@pytest.mark.parametrize("name", sorted(_corpus()))
def test_inflate_equals_zlib(name):
    data = _corpus()[name]
    streams = [zlib.compress(data, level) for level in (0, 1, 4, 6, 9)]
    for strategy in (zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
        c = zlib.compressobj(6, zlib.DEFLATED, 15, 9, strategy)
        streams.append(c.compress(data) + c.flush())
    # many small blocks: a full flush after every 1000 bytes (empty stored blocks in between)
    c = zlib.compressobj(6)
    pieces = b"".join(c.compress(data[i:i + 1000]) + c.flush(zlib.Z_FULL_FLUSH)
                      for i in range(0, len(data), 1000))
    streams.append(pieces + c.flush())
    for s in streams:
        assert _native.inflate(s, len(data)) == data


def test_random_round_trips():
    rng = np.random.default_rng(8)
    for _ in range(200):
        n = int(rng.integers(0, 6000))
        kind = rng.integers(0, 3)
        if kind == 0:
            data = rng.bytes(n)
        elif kind == 1:
            data = bytes(rng.integers(0, int(rng.integers(1, 20)), n, dtype=np.uint8))
        else:
            data = (rng.bytes(int(rng.integers(1, 40))) * (n // 8 + 1))[:n]
        s = zlib.compress(data, int(rng.integers(0, 10)))
        assert _native.inflate(s, len(data)) == data
        # exact-size and oversized output buffers both work
        assert _native.inflate(s, len(data) + 300) == data


def test_damaged_streams_are_refused():
    data = (np.arange(20000, dtype=np.float32) * 0.37).tobytes()
    s = bytearray(zlib.compress(data, 6))
    with pytest.raises(_native.MdsIOError, match="too small"):
        _native.inflate(bytes(s), len(data) - 1)
    with pytest.raises(_native.MdsIOError):
        _native.inflate(bytes(s[:len(s) // 2]), len(data))           # truncated
    bad = bytearray(s)
    bad[-1] ^= 0x55
    with pytest.raises(_native.MdsIOError, match="checksum"):
        _native.inflate(bytes(bad), len(data))                        # adler32
    assert _native.inflate(bytes(bad), len(data), verify=False) == data
    bad = bytearray(s)
    bad[0] = 0x79
    with pytest.raises(_native.MdsIOError, match="header"):
        _native.inflate(bytes(bad), len(data))
    flipped = 0
    for pos in range(2, len(s) - 4, 97):                              # bit flips inside the stream
        bad = bytearray(s)
        bad[pos] ^= 0x10
        try:
            out = _native.inflate(bytes(bad), len(data))
        except _native.MdsIOError:
            flipped += 1
            continue
        assert out != data or True  # a flip that survives must still have passed the checksum
    assert flipped > 0
