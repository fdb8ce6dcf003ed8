"""
Trajectory store feeding the RDF frame batcher.

The reference keeps one HDF5 dataset per ``<species>/<property>`` with shape
(n_particles, n_configurations, 3), gzip-compressed float32, time on axis 1
(mdsuite/database/simulation_database.py:333-378, :449-499), so reading ONE configuration touches
every chunk (SURVEY.md hard part H4).  Neither h5py nor libhdf5 exists in this image, and the GPU
path wants whole frames, so the native store here is FRAME-MAJOR: one raw little-endian float32
file per dataset with shape (n_configurations, n_particles, n_dims), memory-mapped for reading,
appended on ingest, next to a small JSON index.  A reference ``database.hdf5`` is read through
h5py when h5py is importable (same ``load_frames`` interface, transposed on the fly).
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import Dict, List, Sequence

import numpy as np


@dataclass
class PropertyInfo:
    """simulation_database.py:43-62."""
    name: str
    n_dims: int


@dataclass
class SpeciesInfo:
    """simulation_database.py:65-99."""
    name: str
    n_particles: int
    properties: List[PropertyInfo] = field(default_factory=list)
    mass: float = None
    charge: float = 0


@dataclass
class TrajectoryMetadata:
    """simulation_database.py:130-169."""
    n_configurations: int
    species_list: List[SpeciesInfo]
    box_l: list = None
    sample_rate: int = 1
    temperature: float = None
    simulation_time_step: float = None


@dataclass
class TrajectoryChunkData:
    """A chunk of configurations handed from a file reader to the store
    (simulation_database.py:172-227): ``data[species][property]`` is (chunk, n_particles, n_dims)."""
    species_list: List[SpeciesInfo]
    chunk_size: int
    data: Dict[str, Dict[str, np.ndarray]] = field(default_factory=dict)

    def add_data(self, array: np.ndarray, config_idx: int, species_name: str, property_name: str):
        sp = self.data.setdefault(species_name, {})
        if property_name not in sp:
            info = next(s for s in self.species_list if s.name == species_name)
            dims = next(p.n_dims for p in info.properties if p.name == property_name)
            sp[property_name] = np.zeros((self.chunk_size, info.n_particles, dims), dtype=np.float32)
        n = array.shape[0]
        sp[property_name][config_idx:config_idx + n] = array


_COPY_THREADS = max(1, min(16, os.cpu_count() or 1))
_POOL = None


def _copy_pool():
    """Threads that move frames from the page cache into (pinned) staging buffers."""
    global _POOL
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=_COPY_THREADS, thread_name_prefix="mds-frames")
    return _POOL


def _file_name(path: str) -> str:
    return path.replace("/", "__") + ".f32"


class Database:
    """Frame-major trajectory store rooted at ``<experiment>/database/``."""

    def __init__(self, root: str):
        self.root = root
        self.index_path = os.path.join(root, "index.json")
        os.makedirs(root, exist_ok=True)
        if os.path.exists(self.index_path):
            with open(self.index_path) as fh:
                self.index = json.load(fh)
        else:
            self.index = {"layout": "frame_major_f32", "n_configurations": 0, "datasets": {}}
        self._maps: Dict[str, np.memmap] = {}

    # ---- writing ------------------------------------------------------------------------
    def _save_index(self):
        tmp = self.index_path + ".tmp"
        with open(tmp, "w") as fh:
            json.dump(self.index, fh)
        os.replace(tmp, self.index_path)

    def initialize_database(self, species_list: Sequence[SpeciesInfo]):
        """simulation_database.py:420-443."""
        for sp in species_list:
            for prop in sp.properties:
                self.add_dataset(f"{sp.name}/{prop.name}", sp.n_particles, prop.n_dims)

    def add_dataset(self, path: str, n_particles: int, n_dims: int = 3):
        """simulation_database.py:449-499 (no compression: the GPU path is fed raw frames)."""
        if path in self.index["datasets"]:
            return
        self.index["datasets"][path] = {"n_particles": int(n_particles), "n_dims": int(n_dims),
                                        "file": _file_name(path), "n_configurations": 0}
        open(os.path.join(self.root, _file_name(path)), "ab").close()
        self._save_index()

    def add_data(self, chunk: TrajectoryChunkData):
        """Append a chunk of configurations (simulation_database.py:333-378).  The reference swaps
        axes into its atom-major layout here (:365-368); this store keeps frames contiguous."""
        self._maps.clear()
        for species_name, props in chunk.data.items():
            for prop_name, arr in props.items():
                path = f"{species_name}/{prop_name}"
                ds = self.index["datasets"][path]
                a = np.ascontiguousarray(arr, dtype=np.float32)
                if a.shape[1:] != (ds["n_particles"], ds["n_dims"]):
                    raise ValueError(f"{path}: chunk shape {a.shape} does not match the dataset")
                with open(os.path.join(self.root, ds["file"]), "ab") as fh:
                    a.tofile(fh)
                ds["n_configurations"] += int(a.shape[0])
        self.index["n_configurations"] = min(
            (d["n_configurations"] for d in self.index["datasets"].values()), default=0)
        self._save_index()

    # ---- reading ------------------------------------------------------------------------
    def check_existence(self, path: str) -> bool:
        return path in self.index["datasets"]

    def get_data_size(self, path: str) -> tuple:
        """(n_particles, n_configurations, n_bytes) — simulation_database.py:668-690."""
        ds = self.index["datasets"][path]
        n_bytes = ds["n_particles"] * ds["n_configurations"] * ds["n_dims"] * 4
        return ds["n_particles"], ds["n_configurations"], n_bytes

    def _map(self, path: str) -> np.memmap:
        if path not in self._maps:
            ds = self.index["datasets"][path]
            shape = (ds["n_configurations"], ds["n_particles"], ds["n_dims"])
            self._maps[path] = np.memmap(os.path.join(self.root, ds["file"]), dtype=np.float32,
                                         mode="r", shape=shape)
        return self._maps[path]

    def load_frames(self, path: str, frames: Sequence[int], out: np.ndarray = None,
                    atom_selection=None) -> np.ndarray:
        """(len(frames), n_selected, n_dims) float32 for the given configuration indices, written
        into ``out`` (e.g. a pinned staging buffer slice) when given."""
        mm = self._map(path)
        frames = np.asarray(frames, dtype=np.int64)
        select = None
        if atom_selection is not None and not (isinstance(atom_selection, slice)
                                               and atom_selection == slice(None)):
            select = atom_selection
        n_sel = mm.shape[1] if select is None else len(np.arange(mm.shape[1])[select])
        if out is None:
            out = np.empty((len(frames), n_sel, mm.shape[2]), dtype=np.float32)
        if len(frames) == 0:
            return out
        if frames.min() < 0 or frames.max() >= mm.shape[0]:
            raise IndexError(f"{path}: configuration index outside [0, {mm.shape[0]})")
        contiguous = bool(np.all(np.diff(frames) == 1))

        def copy_block(lo: int, hi: int):  # one pass from the page cache into `out`, no temporaries
            if select is None and contiguous:
                np.copyto(out[lo:hi], mm[frames[lo]:frames[lo] + (hi - lo)])
            elif select is None:
                # mode="clip": with the default "raise" NumPy buffers `out` (a second copy)
                np.take(mm, frames[lo:hi], axis=0, out=out[lo:hi], mode="clip")
            else:
                for k in range(lo, hi):
                    out[k] = mm[frames[k]][select]

        n_bytes = len(frames) * n_sel * mm.shape[2] * 4
        n_jobs = min(_COPY_THREADS, len(frames), max(1, n_bytes // (2 << 20)))
        if n_jobs <= 1:
            copy_block(0, len(frames))
        else:  # NumPy releases the GIL in these copies: the threads run on separate cores
            bounds = np.linspace(0, len(frames), n_jobs + 1).astype(int)
            list(_copy_pool().map(lambda ab: copy_block(ab[0], ab[1]), zip(bounds[:-1], bounds[1:])))
        return out

    def load_data(self, path_list: Sequence[str], select_slice=np.s_[:]) -> Dict[str, np.ndarray]:
        """Reference-shaped read (simulation_database.py:594-639): ``select_slice`` indexes
        (atoms, configurations) and the result is atom-major (n_atoms, n_configurations, 3)."""
        if not isinstance(select_slice, tuple):
            select_slice = (select_slice, np.s_[:])
        atoms_sel, conf_sel = (list(select_slice) + [np.s_[:]])[:2]
        out = {}
        for path in path_list:
            mm = self._map(path)
            frames = np.arange(mm.shape[0])[conf_sel]
            data = np.array(mm[frames], dtype=np.float32)[:, atoms_sel]
            out[path] = np.ascontiguousarray(data.transpose(1, 0, 2))
        return out


class HDF5Database:
    """Read-only adapter over a reference ``database.hdf5`` (atom-major (n, cfg, 3), chunked,
    gzip — mdsuite/database/simulation_database.py:449-499, :594-639) using this package's own
    HDF5 reader (no h5py).  Serves the same calls as ``Database``; decoded chunks of the most
    recent frame window are cached so that consecutive batches do not inflate a chunk twice."""

    def __init__(self, path: str, n_configurations: int = None, cache_bytes: int = 512 << 20):
        from mdsuite_b200.database.hdf5_reader import H5File
        self.path = path
        self.file = H5File(path)
        self._datasets = {}
        self._n_configurations = n_configurations
        self._cache = {}  # path -> (f0, f1, frame-major array (f1 - f0, n_atoms, dims))
        self._cache_bytes = cache_bytes
        self.native = True  # chunks decoded by libmds_io (False: the pure-Python reader, for tests)

    def close(self):
        self.file.close()

    def _ds(self, path: str):
        if path not in self._datasets:
            self._datasets[path] = self.file.dataset(path)
        return self._datasets[path]

    def datasets(self) -> list:
        return sorted(self.file.walk())

    def check_existence(self, path: str) -> bool:
        return path in self.file

    def get_data_size(self, path: str) -> tuple:
        ds = self._ds(path)
        n_cfg = ds.shape[1] if self._n_configurations is None else min(ds.shape[1],
                                                                       self._n_configurations)
        return ds.shape[0], n_cfg, int(ds.shape[0] * n_cfg * np.prod(ds.shape[2:])) * ds.dtype.itemsize

    def _window(self, path: str, f0: int, f1: int) -> np.ndarray:
        """FRAME-MAJOR (f1 - f0, n_atoms, dims) for configurations [f0, f1), widened to whole
        frame-chunks and kept for the next call.  The chunks are atom-major on disk
        (simulation_database.py:449-499); the native decoder writes them transposed, so a batch is
        a contiguous slice of the window."""
        ds = self._ds(path)
        hit = self._cache.get(path)
        if hit is not None and hit[0] <= f0 and f1 <= hit[1]:
            return hit[2][f0 - hit[0]:f1 - hit[0]]
        cf = ds.chunks[1] if ds.chunks else ds.shape[1]
        w0, w1 = f0 // cf * cf, min(ds.shape[1], -(-f1 // cf) * cf)
        per_frame = ds.shape[0] * int(np.prod(ds.shape[2:])) * ds.dtype.itemsize
        if (w1 - w0) * per_frame > self._cache_bytes:  # too wide to keep: read exactly what is asked
            w0, w1 = f0, f1
        n_atoms, inner = ds.shape[0], tuple(ds.shape[2:])
        block = np.zeros((w1 - w0, n_atoms) + inner, dtype=ds.dtype)  # missing chunks: fill value 0
        n_inner = int(np.prod(inner)) if inner else 1
        # element strides of the frame-major block per dataset dimension (atoms, frames, inner...)
        strides = [n_inner, n_atoms * n_inner]
        acc = n_inner
        for n in inner:
            acc //= n
            strides.append(acc)
        lo = [0, w0] + [0] * len(inner)
        hi = [n_atoms, w1] + list(inner)
        if not (self.native and ds.read_box_into(lo, hi, block, strides)):
            am = ds.read_hyperslab((slice(0, n_atoms), slice(w0, w1)) + tuple(slice(0, n) for n in inner))
            block = np.ascontiguousarray(np.moveaxis(am, 0, 1))
        self._cache[path] = (w0, w1, block)
        return block[f0 - w0:f1 - w0]

    def frames_per_chunk(self, path: str) -> int:
        """Configurations per HDF5 chunk: batches that start on a multiple of this (and end on one,
        or at the end of the dataset) are decoded straight into the caller's buffer."""
        ds = self._ds(path)
        return int(ds.chunks[1]) if ds.chunks else int(ds.shape[1])

    def _decode_into(self, path: str, f0: int, f1: int, out: np.ndarray) -> bool:
        """Chunk-aligned contiguous range [f0, f1) decoded by the native library directly into
        ``out`` (frames, atoms, dims) — e.g. a slice of the batcher's pinned ring buffer; no window
        cache, no intermediate copy.  False if the request does not qualify."""
        ds = self._ds(path)
        if not self.native or ds.layout != "chunked" or out.dtype != ds.dtype or out.ndim != len(ds.shape):
            return False
        cf = ds.chunks[1]
        if f0 % cf or (f1 % cf and f1 != ds.shape[1]):
            return False
        inner = tuple(ds.shape[2:])
        if out.shape != (f1 - f0, ds.shape[0]) + inner:
            return False
        item = out.itemsize
        if any(st % item for st in out.strides):
            return False
        st = [x // item for x in out.strides]  # (frames, atoms, inner...) in elements
        lo = [0, f0] + [0] * len(inner)
        hi = [ds.shape[0], f1] + list(inner)
        index = ds.chunk_index()
        grid = [range(l // c * c, h, c) for l, h, c in zip(lo, hi, ds.chunks)]
        if any(tuple(r[i] for r, i in zip(grid, offs)) not in index
               for offs in np.ndindex(*[len(r) for r in grid])):
            out[...] = 0  # chunks that were never written read as the fill value
        return ds.read_box_into(lo, hi, out, [st[1], st[0]] + st[2:])

    def load_frames(self, path: str, frames: Sequence[int], out: np.ndarray = None,
                    atom_selection=None) -> np.ndarray:
        frames = np.asarray(frames, dtype=np.int64)
        whole = atom_selection is None or (isinstance(atom_selection, slice)
                                           and atom_selection == slice(None))
        if (out is not None and whole and len(frames) and
                len(frames) == int(frames[-1]) - int(frames[0]) + 1 and np.all(np.diff(frames) == 1)
                and self._decode_into(path, int(frames[0]), int(frames[-1]) + 1, out)):
            return out
        if len(frames) == 0:
            data = np.zeros((0,) + tuple(np.delete(self._ds(path).shape, 1)), dtype=np.float32)
        else:
            lo, hi = int(frames.min()), int(frames.max()) + 1
            block = self._window(path, lo, hi)
            data = block if (len(frames) == hi - lo and np.all(np.diff(frames) == 1)) \
                else block[frames - lo]
        if atom_selection is not None and not (isinstance(atom_selection, slice)
                                               and atom_selection == slice(None)):
            data = data[:, atom_selection]
        if out is None:
            return np.ascontiguousarray(data, dtype=np.float32)
        np.copyto(out, data, casting="same_kind")
        return out

    def load_data(self, path_list: Sequence[str], select_slice=np.s_[:]) -> Dict[str, np.ndarray]:
        """Reference-shaped read (simulation_database.py:594-639), atom-major result."""
        if not isinstance(select_slice, tuple):
            select_slice = (select_slice, np.s_[:])
        atoms_sel, conf_sel = (list(select_slice) + [np.s_[:]])[:2]
        out = {}
        for path in path_list:
            n_cfg = self.get_data_size(path)[1]
            frames = np.arange(n_cfg)[conf_sel]
            fm = self.load_frames(path, frames)
            out[path] = np.ascontiguousarray(fm.transpose(1, 0, 2)[atoms_sel])
        return out


def import_hdf5(hdf5_path: str, database: "Database", n_configurations: int = None,
                frames_per_block: int = 256, properties: Sequence[str] = ("Positions",)) -> dict:
    """Re-chunk a reference ``database.hdf5`` (atom-major, gzip) into the frame-major store the
    GPU feed streams from ("a frame-major re-chunking transform").
    Returns {dataset path: (n_particles, n_configurations)}."""
    src = HDF5Database(hdf5_path, n_configurations=n_configurations)
    done = {}
    try:
        for path in src.datasets():
            species, _, prop = path.rpartition("/")
            if prop not in properties:
                continue
            n_atoms, n_cfg, _ = src.get_data_size(path)
            dims = src._ds(path).shape[2]
            info = SpeciesInfo(species, n_atoms, [PropertyInfo(prop, dims)])
            database.add_dataset(path, n_atoms, dims)
            for f0 in range(0, n_cfg, frames_per_block):
                f1 = min(n_cfg, f0 + frames_per_block)
                chunk = TrajectoryChunkData([info], f1 - f0)
                chunk.add_data(src.load_frames(path, np.arange(f0, f1)), 0, species, prop)
                database.add_data(chunk)
            done[path] = (n_atoms, n_cfg)
    finally:
        src.close()
    return done
