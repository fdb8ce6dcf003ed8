"""Generate tests/golden/reader_golden.json by EXECUTING THE REFERENCE'S OWN trajectory readers

    mdsuite/file_io/lammps_trajectory_files.py   (LAMMPSTrajectoryFile)
    mdsuite/file_io/extxyz_files.py              (EXTXYZFile)
    mdsuite/file_io/tabular_text_files.py        (TabularTextFileProcessor)
    mdsuite/database/simulation_database.py      (TrajectoryChunkData, metadata classes)
    mdsuite/utils/meta_functions.py              (sort_array_by_column, optimize_batch_size)

loaded unmodified from /root/reference, with stand-ins only for the packages they import at module
level and never touch on this path (tensorflow, h5py, GPUtil).  Every fixture file under
tests/golden/ named in FILES goes through ``processor.metadata`` and
``processor.get_configurations_generator()``; the float64 chunk values are stored as the float32
the reference's HDF5 datasets hold (database/simulation_database.py:491-497 creates them with
dtype float32), hex encoded.

Runs only where /root/reference exists.  Re-run:  python tests/golden/make_reader_golden.py
"""
import copy
import importlib.util
import json
import os
import sys
import types
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/mdsuite"

FILES = [
    ("lammps_golden.lammpstraj", "lammps", {}),
    ("lammps_unsorted_types.lammpstraj", "lammps", {}),
    ("extxyz_golden.extxyz", "extxyz", {}),
    ("extxyz_forces_first.extxyz", "extxyz", {}),
    # a property the reference does not know ("Z:I:1") in front of known ones: its column counter
    # does not advance (extxyz_files.py:279-288), so "pos" is read from columns 1-3 — recorded as is
    ("extxyz_unknown_property.extxyz", "extxyz", {}),
    # reader keyword arguments: rows taken in file order; user-defined properties
    ("lammps_golden.lammpstraj", "lammps", {"trajectory_is_sorted_by_ids": True}),
    ("lammps_unsorted_types.lammpstraj", "lammps",
     {"custom_data_map": {"Scaled_Charge_Pair": ["q", "xs"]}}),
    ("extxyz_custom_property.extxyz", "extxyz", {"custom_data_map": {"Reduced_Momentum": "redmom"}}),
]


def case_key(fname, kwargs):
    """Key of a case in reader_golden.json: the file name, plus its keyword arguments if any."""
    return fname if not kwargs else fname + "?" + json.dumps(kwargs, sort_keys=True)


def load_reference():
    stubs = {}
    for name in ("tensorflow", "h5py", "GPUtil"):
        stubs[name] = mock.MagicMock()
    for pkg in ("mdsuite", "mdsuite.file_io", "mdsuite.utils", "mdsuite.database"):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REF, *pkg.split(".")[1:])]
        stubs[pkg] = m
    sys.modules.update(stubs)

    def load(modname, relpath):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, relpath))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        parent, _, leaf = modname.rpartition(".")
        spec.loader.exec_module(mod)
        setattr(sys.modules[parent], leaf, mod)
        return mod

    sys.modules["mdsuite"].utils = sys.modules["mdsuite.utils"]
    sys.modules["mdsuite"].database = sys.modules["mdsuite.database"]
    sys.modules["mdsuite"].file_io = sys.modules["mdsuite.file_io"]
    load("mdsuite.utils.exceptions", "utils/exceptions.py")
    load("mdsuite.utils.units", "utils/units.py")
    load("mdsuite.utils.meta_functions", "utils/meta_functions.py")
    load("mdsuite.database.simulation_database", "database/simulation_database.py")
    load("mdsuite.database.mdsuite_properties", "database/mdsuite_properties.py")
    load("mdsuite.file_io.file_read", "file_io/file_read.py")
    load("mdsuite.file_io.tabular_text_files", "file_io/tabular_text_files.py")
    lammps = load("mdsuite.file_io.lammps_trajectory_files", "file_io/lammps_trajectory_files.py")
    extxyz = load("mdsuite.file_io.extxyz_files", "file_io/extxyz_files.py")
    return lammps, extxyz


def hex32(a):
    return np.ascontiguousarray(a, dtype="<f4").tobytes().hex()


def run(processor):
    md = processor.metadata
    out = {
        "n_configurations": int(md.n_configurations),
        "box_l": [float(v) for v in md.box_l],
        "sample_rate": None if md.sample_rate is None else int(md.sample_rate),
        "species": [{"name": sp.name, "n_particles": int(sp.n_particles),
                     "properties": [[p.name, int(p.n_dims)] for p in sp.properties]}
                    for sp in md.species_list],
        "species_line_idx": {k: [int(i) for i in v] for k, v in
                             processor.tabular_text_reader_data
                             .species_name_to_line_idx_dict.items()},
    }
    chunks = [c.get_data() for c in processor.get_configurations_generator()]
    data = {}
    for sp in md.species_list:
        data[sp.name] = {}
        for p in sp.properties:
            full = np.concatenate([c[sp.name][p.name] for c in chunks], axis=0)
            assert full.shape == (md.n_configurations, sp.n_particles, p.n_dims)
            with np.errstate(over="ignore"):
                data[sp.name][p.name] = {"shape": list(full.shape),
                                         "float32_hex": hex32(full.astype(np.float32))}
    out["data"] = data
    return out


def main():
    lammps, extxyz = load_reference()
    golden = {"_doc": "output of the reference's own LAMMPSTrajectoryFile / EXTXYZFile "
                      "(tests/golden/make_reader_golden.py); data = float32 of the chunk values, "
                      "little-endian hex, C order (configuration, particle, dimension)"}
    for fname, kind, kwargs in FILES:
        path = os.path.join(HERE, fname)
        cls = lammps.LAMMPSTrajectoryFile if kind == "lammps" else extxyz.EXTXYZFile
        key = case_key(fname, kwargs)
        # the reference writes column indices into the user's custom_data_map lists: give it a copy
        golden[key] = run(cls(path, **copy.deepcopy(kwargs)))
        golden[key]["file"] = fname
        golden[key]["kwargs"] = kwargs
        print(key, golden[key]["n_configurations"], golden[key]["box_l"],
              golden[key]["sample_rate"], [s["name"] for s in golden[key]["species"]])
    with open(os.path.join(HERE, "reader_golden.json"), "w") as fh:
        json.dump(golden, fh, indent=1)
        fh.write("\n")


if __name__ == "__main__":
    main()
