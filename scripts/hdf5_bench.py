"""Throughput of the from-scratch HDF5 reader on a reference-shaped database (atom-major, chunked,
gzip): frame-major batches as the RDF batcher requests them, and the import transform.
Usage: python scripts/hdf5_bench.py [n_frames]     (prints one JSON line)"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from hdf5_writer import H5Writer  # noqa: E402  (no libhdf5 in the image: the test writer makes the file)
from mdsuite_b200 import synthetic  # noqa: E402
from mdsuite_b200.database.simulation_database import Database, HDF5Database, import_hdf5  # noqa: E402

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 512
# every frame its own Gaussian displacements: float32 positions barely compress (ratio ~0.9), as
# in a real trajectory — repeated frames would flatter the inflate rate
sysm = synthetic.nacl_melt(12, n_frames, seed=2)
pos = sysm.positions
am = np.ascontiguousarray(pos.transpose(1, 0, 2))
n = sysm.counts[0]
with tempfile.TemporaryDirectory() as tmp:
    w = H5Writer()
    # h5py's guess_chunk on (6912, F, 3) float32 lands near (432, F/8, 3); use that shape
    cf = 32
    w.create_dataset("Na/Positions", am[:n], chunks=(432, cf, 3), maxshape=(n, None, 3))
    w.create_dataset("Cl/Positions", am[n:], chunks=(432, cf, 3), maxshape=(n, None, 3))
    path = os.path.join(tmp, "database.hdf5")
    w.save(path)
    raw_mb = pos.nbytes / 1e6
    out = {"atoms": sysm.n_atoms, "frames": n_frames, "raw_mb": raw_mb,
           "file_mb": os.path.getsize(path) / 1e6, "cores": os.cpu_count(),
           "decoder": "libmds_io: pread + own inflate + strided placement, all cores"}
    db = HDF5Database(path)
    db.load_frames("Na/Positions", np.arange(0, 1))  # chunk index, library load
    db._cache.clear()
    t0 = time.perf_counter()
    batch = 4 * cf  # whole HDF5 chunks per batch, as the calculator's frame batcher picks them
    # as the frame batcher calls it: both species into one (frames, atoms, 3) staging buffer
    stage = np.empty((batch, sysm.n_atoms, 3), dtype=np.float32)
    for f0 in range(0, n_frames, batch):
        idx = np.arange(f0, min(n_frames, f0 + batch))
        db.load_frames("Na/Positions", idx, out=stage[:len(idx), :n])
        db.load_frames("Cl/Positions", idx, out=stage[:len(idx), n:])
    dt = time.perf_counter() - t0
    ok = np.array_equal(stage[:len(idx)], pos[idx])
    out["batched_read"] = {"seconds": dt, "raw_mb_per_s": raw_mb / dt, "batch_frames": batch,
                           "last_batch_matches": bool(ok)}
    db.close()
    store = Database(os.path.join(tmp, "store"))
    t0 = time.perf_counter()
    import_hdf5(path, store, frames_per_block=64)
    dt = time.perf_counter() - t0
    out["import_to_frame_store"] = {"seconds": dt, "raw_mb_per_s": raw_mb / dt}
print(json.dumps(out))
