"""Write tests/golden/reference_database.hdf5 with h5py / libhdf5 exactly as the reference does, so
that mdsuite_b200's from-scratch HDF5 reader can be pinned against a file it did not write itself.

NOT RUNNABLE IN THE BUILD IMAGE (neither h5py nor libhdf5 exists there); run it wherever h5py is
installed:

    python tests/golden/make_hdf5_golden.py          # writes reference_database.hdf5 + .json

What it reproduces (reference mdsuite/database/simulation_database.py):
  * add_dataset :448-499 — ``create_dataset(path, shape, maxshape=(n, None, 3),
    compression="gzip", chunks=True)`` and the ``starting_index`` attribute;
  * add_data :333-378 — batches written with ``dataset[:, start:stop, :] = ...`` after a
    ``resize`` along axis 1 (resize_datasets :380-410), which leaves partially filled edge chunks;
  * the group layout ``<species>/<property>``.

The arrays are seeded (mdsuite_b200.synthetic), so the test (tests/test_hdf5_reader.py::
test_reference_written_file) regenerates them and needs only the .hdf5 file.  Until that file has
been produced and committed the reader stays UNPINNED against libhdf5."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from mdsuite_b200 import synthetic  # noqa: E402

CASES = {"n_cells": 3, "frames": 150, "seed": 4711, "batch": 64}


def golden_arrays():
    sysm = synthetic.nacl_melt(CASES["n_cells"], CASES["frames"], seed=CASES["seed"])
    am = np.ascontiguousarray(sysm.positions.transpose(1, 0, 2))  # (N, frames, 3), atom-major
    n_na = sysm.counts[0]
    return {"Na/Positions": am[:n_na], "Cl/Positions": am[n_na:],
            "Na/Velocities": (am[:n_na] * np.float32(0.5)).astype(np.float32)}


def main():
    import h5py as hf  # noqa: WPS433 - only where the script is meant to run
    path = os.path.join(HERE, "reference_database.hdf5")
    if os.path.exists(path):
        os.remove(path)
    arrays = golden_arrays()
    batch = CASES["batch"]
    # initial size = one batch, as add_dataset is first called with the first chunk's shape
    with hf.File(path, "a") as database:
        for name, arr in arrays.items():
            shape = (arr.shape[0], batch, 3)
            database.create_dataset(name, shape, maxshape=(arr.shape[0], None, 3),
                                    compression="gzip", chunks=True)
            database[name].attrs["starting_index"] = 0
    n_frames = CASES["frames"]
    for start in range(0, n_frames, batch):
        stop = min(n_frames, start + batch)
        with hf.File(path, "r+") as database:
            for name, arr in arrays.items():
                ds = database[name]
                if ds.shape[1] < stop:
                    ds.resize(stop, axis=1)  # resize_datasets :380-410
                ds[:, start:stop, :] = arr[:, start:stop, :]
    meta = {"h5py": hf.__version__, "hdf5": hf.version.hdf5_version, "cases": CASES,
            "datasets": {}}
    with hf.File(path, "r") as database:
        for name in arrays:
            ds = database[name]
            meta["datasets"][name] = {"shape": list(ds.shape), "chunks": list(ds.chunks),
                                      "compression": ds.compression,
                                      "compression_opts": ds.compression_opts}
    with open(os.path.join(HERE, "reference_database.json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    print("wrote", path, meta)


if __name__ == "__main__":
    main()
