"""Native trajectory ingestion (include/mds_io.h, mdsuite_b200/_native): the C++ LAMMPS dump reader
against (1) a hand-written golden dump, (2) the pure-Python restatement of the reference's reader
(oracle/lammps_oracle.py) on seeded synthetic dumps, (3) NumPy's string -> float64 -> float32
conversion on adversarial number strings.  Bit-exact float32 everywhere (tolerance zero)."""
import json
import os
import re

import numpy as np
import pytest

from mdsuite_b200 import _native, synthetic
from mdsuite_b200.file_io.lammps_trajectory_files import LAMMPSTrajectoryFile
from oracle.lammps_oracle import LAMMPSTrajectoryFileOracle

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_library_exports_every_symbol_of_the_header():
    lib = _native.load()
    with open(os.path.join(ROOT, "include", "mds_io.h")) as fh:
        declared = sorted(set(re.findall(r"MDS_IO_API\s+[\w\s\*]+?\b(mds_\w+)\s*\(", fh.read())))
    assert len(declared) >= 9
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mds_io.h but not exported"
    assert sorted(_native.EXPORTED_SYMBOLS) == declared
    assert lib.mds_io_abi_version() == _native.ABI_VERSION


def test_golden_dump_by_hand():
    with open(os.path.join(HERE, "golden", "lammps_golden.json")) as fh:
        g = json.load(fh)
    want = np.array([[[int(h, 16) for h in atom] for atom in frame] for frame in g["positions_hex"]],
                    dtype=np.uint32)
    proc = LAMMPSTrajectoryFile(os.path.join(HERE, "golden", "lammps_golden.lammpstraj"))
    md = proc.metadata
    assert md.n_configurations == g["n_configurations"]
    assert md.box_l == g["box_l"] and md.sample_rate == g["sample_rate"]
    assert {s.name: s.n_particles for s in md.species_list} == g["species"]
    chunks = list(proc.get_configurations_generator())
    assert len(chunks) == 1
    got = np.concatenate([chunks[0].data["Na"]["Positions"], chunks[0].data["Cl"]["Positions"]], 1)
    assert got.dtype == np.float32
    assert np.array_equal(got.view(np.uint32), want)
    assert "Velocities" in chunks[0].data["Na"]


def _metadata_tuple(md):
    return (md.n_configurations, [float(v) for v in md.box_l], md.sample_rate,
            [(s.name, s.n_particles, [(p.name, p.n_dims) for p in s.properties])
             for s in md.species_list])


def _write_dump(path, rng, n_atoms, n_frames, shuffle=True, element=True, extra_cols=True,
                crlf=False, fmt="%.9g", trailing_newline=True):
    names = ["Na", "Cl", "K"]
    kinds = rng.integers(0, 3, n_atoms)
    lines = []
    for f in range(n_frames):
        pos = rng.random((n_atoms, 3)) * [20.0, 30.0, 40.0] - [0.0, 5.0, 0.0]
        vel = rng.normal(size=(n_atoms, 3))
        lines += ["ITEM: TIMESTEP", str(100 * f), "ITEM: NUMBER OF ATOMS", str(n_atoms),
                  "ITEM: BOX BOUNDS pp pp pp", "0 20", "-5.0 25.0", "0.0e0 4.0e1"]
        cols = ["id", "element" if element else "type", "x", "y", "z"]
        if extra_cols:
            cols = ["id", "type", "q"] + cols[1:] + ["vx", "vy", "vz"] if element else cols + ["vx", "vy", "vz"]
        lines.append("ITEM: ATOMS " + " ".join(cols))
        order = rng.permutation(n_atoms) if shuffle else np.arange(n_atoms)
        for i in order:
            label = names[kinds[i]] if element else str(kinds[i] + 1)
            row = {"id": str(i + 1), "type": str(kinds[i] + 1), "element": label, "q": "0.5",
                   "x": fmt % pos[i, 0], "y": fmt % pos[i, 1], "z": fmt % pos[i, 2],
                   "vx": "%g" % vel[i, 0], "vy": "%g" % vel[i, 1], "vz": "%g" % vel[i, 2]}
            lines.append(" ".join(row[c] for c in cols) + (" " if f % 2 else ""))
    text = ("\r\n" if crlf else "\n").join(lines) + (("\r\n" if crlf else "\n") if trailing_newline else "")
    with open(path, "w", newline="") as fh:
        fh.write(text)


@pytest.mark.parametrize("case", [
    dict(n_atoms=37, n_frames=5),
    dict(n_atoms=64, n_frames=3, shuffle=False, element=False, extra_cols=False),
    dict(n_atoms=50, n_frames=4, crlf=True, fmt="%.17g"),
    dict(n_atoms=21, n_frames=6, fmt="%.6e", trailing_newline=False),
    dict(n_atoms=1, n_frames=1),
])
def test_native_reader_equals_python_restatement(tmp_path, case):
    path = str(tmp_path / "t.lammpstraj")
    _write_dump(path, np.random.default_rng(11), **case)
    native = LAMMPSTrajectoryFile(path, chunk_bytes=4096, n_threads=3)
    oracle = LAMMPSTrajectoryFileOracle(path, chunk_configurations=2)
    assert _metadata_tuple(native.metadata) == _metadata_tuple(oracle.metadata)

    def gather(proc):
        out = {}
        for chunk in proc.get_configurations_generator():
            for sp, props in chunk.data.items():
                for prop, arr in props.items():
                    out.setdefault((sp, prop), []).append(np.array(arr))
        return {k: np.concatenate(v) for k, v in out.items()}

    a, b = gather(native), gather(oracle)
    assert a.keys() == b.keys()
    for key in a:
        assert a[key].dtype == np.float32 and a[key].shape == b[key].shape
        assert np.array_equal(a[key].view(np.uint32), b[key].view(np.uint32)), key


def test_ids_that_are_not_a_permutation_of_1_to_n(tmp_path):
    path = str(tmp_path / "ids.lammpstraj")
    with open(path, "w") as fh:
        fh.write("ITEM: TIMESTEP\n0\nITEM: NUMBER OF ATOMS\n3\nITEM: BOX BOUNDS pp pp pp\n0 1\n0 1\n0 1\n"
                 "ITEM: ATOMS id type x y z\n700 1 0.7 0 0\n5 2 0.05 0 0\n12 1 0.12 0 0\n")
    with _native.LammpsDump(path) as d:
        x = d.read_columns(0, 1, [2])[0, :, 0]
        assert x.tolist() == [np.float32(0.05), np.float32(0.12), np.float32(0.7)]
        assert d.first_frame_labels(1) == ["2", "1", "1"]
        assert d.read_columns(0, 1, [2], sort_by_id=False)[0, :, 0].tolist() == [
            np.float32(0.7), np.float32(0.05), np.float32(0.12)]


def test_number_conversion_equals_numpy_double_then_float():
    rng = np.random.default_rng(5)
    tokens = ["0", "-0", "+0.0", "1", "-1", ".5", "5.", "1e0", "1E+00", "1e-0", "0.1", "0.2", "0.3",
              "16777217", "16777216.999999999", "0.30000001192092896", "3.4028235e38", "3.4028236e38",
              "1e39", "-1e39", "1.17549435e-38", "1.4e-45", "7e-46", "4.9406564584124654e-324",
              "2.2250738585072014e-308", "1e-400", "1e400", "123456789012345678901234567890",
              "0.000000000000000000000000000001", "9007199254740993", "9007199254740992",
              "1.7976931348623157e308", "nan", "inf", "-inf", "NaN", "Infinity", "1e22", "1e23",
              "8.5e-23", "123456789.123456789e-5", "00012.5000"]
    for _ in range(4000):
        digits = int(rng.integers(1, 22))
        mant = "".join(rng.choice(list("0123456789"), digits))
        point = int(rng.integers(0, digits + 1))
        s = ("-" if rng.random() < 0.3 else "") + mant[:point] + "." + mant[point:]
        if rng.random() < 0.5:
            s += "e%+d" % int(rng.integers(-45, 40))
        tokens.append(s)
    for _ in range(2000):
        tokens.append("%.9g" % (rng.normal() * 10.0 ** rng.integers(-8, 8)))
        tokens.append("%.17g" % rng.random())
        tokens.append("%g" % (rng.random() * 100))
    got = _native.parse_float32(tokens)
    with np.errstate(over="ignore"):
        want = np.array([np.float64(t) for t in tokens]).astype(np.float32)
    same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
    bad = [(tokens[i], got[i], want[i]) for i in np.flatnonzero(~same)[:5]]
    assert not bad, bad


def test_format_errors_are_reported(tmp_path):
    p = tmp_path / "bad.lammpstraj"
    p.write_text("ITEM: TIMESTEP\n0\nITEM: NUMBER OF ATOMS\n2\nITEM: BOX BOUNDS pp pp pp\n0 1\n0 1\n0 1\n"
                 "ITEM: ATOMS id type x y z\n1 1 0 0 0\n")  # one atom line missing
    with pytest.raises(_native.MdsIOError, match="whole number of configurations"):
        _native.LammpsDump(str(p))
    p.write_text("ITEM: TIMESTEP\n0\nITEM: NUMBER OF ATOMS\n2\nITEM: BOX BOUNDS pp pp pp\n0 1\n0 1\n0 1\n"
                 "ITEM: ATOMS id type x y z\n1 1 0 0 0\n2 1 0 0\n")
    with _native.LammpsDump(str(p)) as d:
        with pytest.raises(_native.MdsIOError, match="has 4 columns, expected 5"):
            d.read_columns(0, 1, [2, 3, 4])
    with pytest.raises(_native.MdsIOError, match="cannot open"):
        _native.LammpsDump(str(tmp_path / "missing.lammpstraj"))
    p.write_text("hello\n" * 12)
    with pytest.raises(_native.MdsIOError, match="ITEM: TIMESTEP"):
        _native.LammpsDump(str(p))


def test_cfg1_dump_roundtrip_through_the_native_reader(tmp_path):
    """configs[0] of BASELINE.json: the 1,000-atom NaCl melt written as a LAMMPS dump and read back
    must reproduce the float32 positions bit for bit (%.9g round-trips float32)."""
    sysm = synthetic.nacl_melt(5, 7, seed=1)
    path = str(tmp_path / "nacl.lammpstraj")
    synthetic.write_lammps_dump(sysm, path)
    proc = LAMMPSTrajectoryFile(path)
    chunks = list(proc.get_configurations_generator())
    got = np.concatenate([np.concatenate([c.data["Na"]["Positions"], c.data["Cl"]["Positions"]], 1)
                          for c in chunks])
    assert np.array_equal(got.view(np.uint32), sysm.positions.view(np.uint32))


def test_every_reference_property_and_custom_columns(tmp_path):
    """The reference's column-name table (positions in four flavours, velocities, forces, charge,
    per-atom energies, stress ...) plus a custom_data_map entry."""
    path = tmp_path / "props.lammpstraj"
    cols = ["id", "type", "x", "y", "z", "xu", "yu", "zu", "q", "c_KE", "c_Stress[1]", "c_Stress[2]",
            "c_Stress[3]", "c_Stress[4]", "c_Stress[5]", "c_Stress[6]", "rp_x", "rp_y", "rp_z"]
    rows = []
    for i in (2, 1, 3):
        rows.append(" ".join([str(i), "1"] + ["%d.%d" % (i, k) for k in range(len(cols) - 2)]))
    frame = ("ITEM: TIMESTEP\n0\nITEM: NUMBER OF ATOMS\n3\nITEM: BOX BOUNDS pp pp pp\n0 5\n0 5\n0 5\n"
             "ITEM: ATOMS " + " ".join(cols) + "\n" + "\n".join(rows) + "\n")
    path.write_text(frame)
    proc = LAMMPSTrajectoryFile(str(path), custom_data_map={"Reduced_Momentum": ["rp_x", "rp_y", "rp_z"]})
    props = {p.name: p.n_dims for p in proc.metadata.species_list[0].properties}
    assert props == {"Positions": 3, "Unwrapped_Positions": 3, "Charge": 1, "Kinetic_Energy": 1,
                     "Stress": 6, "Reduced_Momentum": 3}
    chunk = next(iter(proc.get_configurations_generator()))
    data = chunk.data["1"]
    assert data["Charge"][0, :, 0].tolist() == [np.float32(1.6), np.float32(2.6), np.float32(3.6)]
    assert data["Stress"].shape == (1, 3, 6)
    assert data["Reduced_Momentum"][0, 2].tolist() == [np.float32(3.14), np.float32(3.15), np.float32(3.16)]
