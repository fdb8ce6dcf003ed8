"""Ingestion throughput (host path): the native LAMMPS dump reader against
the pure-Python restatement of the reference's reader, on a cfg2-shaped synthetic dump.
Usage: python scripts/ingest_bench.py [n_frames] [n_cells]      (prints one JSON line)"""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdsuite_b200 import _native, synthetic  # noqa: E402
from mdsuite_b200.file_io.lammps_trajectory_files import LAMMPSTrajectoryFile  # noqa: E402
from oracle.lammps_oracle import LAMMPSTrajectoryFileOracle  # noqa: E402

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n_cells = int(sys.argv[2]) if len(sys.argv) > 2 else 12
sysm = synthetic.nacl_melt(n_cells, 4, seed=2)
sysm.positions = np.tile(sysm.positions, ((n_frames + 3) // 4, 1, 1))[:n_frames]
with tempfile.TemporaryDirectory() as tmp:
    path = os.path.join(tmp, "bench.lammpstraj")
    synthetic.write_lammps_dump(sysm, path)
    size = os.path.getsize(path)
    with open(path, "rb") as fh:  # warm the page cache: this measures parsing, not the disk
        while fh.read(1 << 24):
            pass
    out = {"file_mb": size / 1e6, "atoms": sysm.n_atoms, "frames": n_frames,
           "cores": os.cpu_count()}
    for threads in (1, 0):
        t0 = time.perf_counter()
        proc = LAMMPSTrajectoryFile(path, n_threads=threads)
        md = proc.metadata
        t1 = time.perf_counter()
        n = 0
        for chunk in proc.get_configurations_generator():
            n += chunk.chunk_size
        t2 = time.perf_counter()
        key = "native_1_thread" if threads == 1 else "native_all_threads"
        out[key] = {"open_index_s": t1 - t0, "parse_s": t2 - t1, "mb_per_s": size / 1e6 / (t2 - t0),
                    "atom_frames_per_s": sysm.n_atoms * n / (t2 - t0)}
    n_py = max(1, min(n_frames, 6))
    t0 = time.perf_counter()
    oracle = LAMMPSTrajectoryFileOracle(path, chunk_configurations=n_py)
    _ = oracle.metadata
    t1 = time.perf_counter()
    next(iter(oracle.get_configurations_generator()))
    t2 = time.perf_counter()
    per_frame = (t2 - t1) / n_py
    out["python_restatement"] = {"analyse_s": t1 - t0, "parse_s_per_frame": per_frame,
                                 "mb_per_s": size / n_frames / 1e6 / per_frame,
                                 "atom_frames_per_s": sysm.n_atoms / per_frame,
                                 "sample": f"first {n_py} configurations"}
    out["speedup_all_threads_vs_python"] = (out["native_all_threads"]["atom_frames_per_s"]
                                            / out["python_restatement"]["atom_frames_per_s"])
print(json.dumps(out))
