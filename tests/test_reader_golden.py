"""Trajectory readers against the REFERENCE'S OWN OUTPUT: tests/golden/reader_golden.json holds what
the reference's LAMMPSTrajectoryFile / EXTXYZFile (executed unmodified by
tests/golden/make_reader_golden.py) produce for the fixture files next to it — metadata, species
line indices and every property as the float32 its HDF5 store would hold.  The native-backed
product readers and the pure-Python restatements under oracle/ must both equal it bit for bit.

Then the extended XYZ product reader against its restatement on seeded synthetic files, and the
error behaviour of the header helpers."""
import json
import os

import numpy as np
import pytest

from mdsuite_b200 import _native, synthetic
from mdsuite_b200.file_io import extxyz_files
from mdsuite_b200.file_io.extxyz_files import EXTXYZFile
from mdsuite_b200.file_io.lammps_trajectory_files import LAMMPSTrajectoryFile
from oracle.extxyz_oracle import EXTXYZFileOracle
from oracle.lammps_oracle import LAMMPSTrajectoryFileOracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

with open(os.path.join(GOLDEN, "reader_golden.json")) as _fh:
    REFERENCE = {k: v for k, v in json.load(_fh).items() if not k.startswith("_")}

READERS = {
    ".lammpstraj": {"product": LAMMPSTrajectoryFile, "oracle": LAMMPSTrajectoryFileOracle},
    ".extxyz": {"product": EXTXYZFile, "oracle": EXTXYZFileOracle},
}


def _chunk_data(chunk):
    return chunk.data if hasattr(chunk, "data") else chunk.get_data()


def _read_all(proc):
    md = proc.metadata
    chunks = [_chunk_data(c) for c in proc.get_configurations_generator()]
    data = {}
    for sp in md.species_list:
        data[sp.name] = {p.name: np.concatenate([c[sp.name][p.name] for c in chunks], axis=0)
                         for p in sp.properties}
    return md, data


def _make(case_key, which):
    """The reader under test for a case of reader_golden.json (file name + keyword arguments)."""
    want = REFERENCE[case_key]
    fname, kwargs = want.get("file", case_key), dict(want.get("kwargs", {}))
    cls = READERS[os.path.splitext(fname)[1]][which]
    return cls(os.path.join(GOLDEN, fname), **kwargs), want, fname


@pytest.mark.parametrize("which", ["product", "oracle"])
@pytest.mark.parametrize("case_key", sorted(REFERENCE))
def test_reader_equals_the_reference_reader(case_key, which):
    proc, want, fname = _make(case_key, which)
    md, data = _read_all(proc)
    assert md.n_configurations == want["n_configurations"]
    assert [float(v) for v in md.box_l] == want["box_l"]
    assert md.sample_rate == want["sample_rate"]
    got_species = [{"name": sp.name, "n_particles": sp.n_particles,
                    "properties": [[p.name, p.n_dims] for p in sp.properties]}
                   for sp in md.species_list]
    key = lambda s: s["name"]
    assert [s["name"] for s in got_species] == [s["name"] for s in want["species"]]
    for g, w in zip(sorted(got_species, key=key), sorted(want["species"], key=key)):
        assert g["n_particles"] == w["n_particles"]
        assert sorted(map(tuple, g["properties"])) == sorted(map(tuple, w["properties"]))
    for sp, props in want["data"].items():
        for prop, enc in props.items():
            ref = np.frombuffer(bytes.fromhex(enc["float32_hex"]), dtype="<f4").reshape(enc["shape"])
            with np.errstate(over="ignore"):
                got = np.asarray(data[sp][prop]).astype(np.float32)
            assert got.shape == ref.shape, (sp, prop)
            assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (fname, sp, prop)


@pytest.mark.parametrize("case_key", sorted(k for k in REFERENCE if ".extxyz" in k))
def test_species_line_indices_equal_the_reference(case_key):
    proc, want, _ = _make(case_key, "product")
    assert proc._analyse()["species"] == want["species_line_idx"]


@pytest.mark.parametrize("case", [
    dict(n_cells=2, frames=5, interleave=True, chunk_bytes=256 << 20),
    dict(n_cells=3, frames=7, interleave=True, chunk_bytes=3 * 216 * 3 * 4),   # 3 frames per chunk
    dict(n_cells=3, frames=4, interleave=False, chunk_bytes=1),               # 1 frame per chunk
])
def test_extxyz_native_reader_equals_python_restatement(tmp_path, case):
    system = synthetic.nacl_melt(case["n_cells"], case["frames"], seed=4242)
    path = str(tmp_path / "melt.extxyz")
    synthetic.write_extxyz(system, path, time_stride=3.0, interleave=case["interleave"])
    md_p, data_p = _read_all(EXTXYZFile(path, chunk_bytes=case["chunk_bytes"]))
    md_o, data_o = _read_all(EXTXYZFileOracle(path, chunk_configurations=2))
    assert md_p.n_configurations == md_o.n_configurations == case["frames"]
    assert md_p.box_l == md_o.box_l and md_p.sample_rate == md_o.sample_rate == 3
    assert [(s.name, s.n_particles) for s in md_p.species_list] == \
           [(s.name, s.n_particles) for s in md_o.species_list]
    for sp in data_o:
        got = np.asarray(data_p[sp]["Positions"], dtype=np.float32)
        want = np.asarray(data_o[sp]["Positions"]).astype(np.float32)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # the positions written with 9 significant digits read back as the float32 that were written
    total = sum(np.asarray(data_p[sp]["Positions"]).shape[1] for sp in data_p)
    assert total == system.n_atoms
    if not case["interleave"]:
        got = np.concatenate([data_p[sp]["Positions"] for sp in system.species], axis=1)
        assert np.array_equal(got.view(np.uint32), system.positions.view(np.uint32))


def test_extxyz_through_add_experiment(tmp_path):
    """The suffix picks the reader (reference experiment.py:62-86) and the store holds the file's
    float32 positions, species-contiguous."""
    import mdsuite_b200 as mds
    system = synthetic.nacl_melt(2, 3, seed=99)
    path = str(tmp_path / "melt.extxyz")
    synthetic.write_extxyz(system, path, interleave=False)
    project = mds.Project("xyz_project", storage_path=str(tmp_path))
    exp = project.add_experiment("melt", timestep=0.002, temperature=1400.0, units="metal",
                                 simulation_data=path)
    assert exp.number_of_configurations == 3
    assert {k: v.n_particles for k, v in exp.species.items()} == system.species
    assert [float(v) for v in exp.box_array] == [float(v) for v in system.box]


def test_extxyz_header_helpers_where_the_reference_raises():
    """Lattice behind other keys, and a single configuration: the reference raises IndexError /
    StopIteration (extxyz_files.py:224, :127); read as intended here."""
    header = 'Properties=species:S:1:pos:R:3 energy=-1.5 Lattice="4.0 0 0 0 5.0 0 0 0 6.0" time=7'
    assert extxyz_files.get_box_l(header) == [4.0, 5.0, 6.0]
    assert extxyz_files.get_time(header) == 7.0
    assert extxyz_files.get_time("time=2.5,") == 2.5
    assert extxyz_files.get_time('Lattice="1 0 0 0 1 0 0 0 1"') is None
    with pytest.raises(RuntimeError):
        extxyz_files.get_box_l("Properties=species:S:1:pos:R:3")
    with pytest.raises(RuntimeError):
        extxyz_files.get_property_to_column_idx_dict('Lattice="1 0 0 0 1 0 0 0 1"',
                                                     extxyz_files.VAR_NAMES)
    with pytest.raises(RuntimeError):
        extxyz_files.get_property_to_column_idx_dict("Properties=pos:R:3", extxyz_files.VAR_NAMES)


def test_extxyz_single_configuration_and_custom_property(tmp_path):
    path = tmp_path / "one.extxyz"
    path.write_text('2\nLattice="9 0 0 0 9 0 0 0 9" Properties=species:S:1:pos:R:3:redmom:R:2\n'
                    "He 1 2 3 0.25 0.5\nNe 4 5 6 0.75 1.0\n")
    proc = EXTXYZFile(str(path), custom_data_map={"Reduced_Momentum": "redmom"})
    md, data = _read_all(proc)
    assert md.n_configurations == 1 and md.sample_rate is None and md.box_l == [9.0, 9.0, 9.0]
    assert np.asarray(data["Ne"]["Reduced_Momentum"]).tolist() == [[[0.75, 1.0]]]
    assert np.asarray(data["He"]["Positions"]).tolist() == [[[1.0, 2.0, 3.0]]]


def test_extxyz_format_errors_are_reported(tmp_path):
    bad = tmp_path / "bad.extxyz"
    bad.write_text("two\nLattice=...\n")
    with pytest.raises(_native.MdsIOError, match="atom count"):
        _native.XyzDump(str(bad))
    ragged = tmp_path / "ragged.extxyz"
    ragged.write_text('2\nLattice="9 0 0 0 9 0 0 0 9" Properties=species:S:1:pos:R:3\n'
                      "He 1 2 3\nNe 4 5\n")
    with pytest.raises(_native.MdsIOError, match="columns"):
        list(EXTXYZFile(str(ragged)).get_configurations_generator())
    short = tmp_path / "short.extxyz"
    short.write_text('2\nLattice="9 0 0 0 9 0 0 0 9" Properties=species:S:1:pos:R:3\n'
                     "He 1 2 3\nNe 4 5 6\n2\n")
    with pytest.raises(_native.MdsIOError, match="whole number"):
        _native.XyzDump(str(short))
    wide = tmp_path / "wide.extxyz"
    wide.write_text('1\nLattice="9 0 0 0 9 0 0 0 9" Properties=species:S:1:pos:R:3:force:R:3\n'
                    "He 1 2 3\n")
    with pytest.raises(ValueError, match="columns"):
        EXTXYZFile(str(wide)).metadata
