"""The from-scratch HDF5 reader (mdsuite_b200/database/hdf5_reader.py) on files made by the test
writer (tests/hdf5_writer.py, the structures h5py emits for an MDSuite database.hdf5) and on the one
foreign HDF5 file the image holds (a MATLAB v7.3 file from SciPy's test data)."""
import glob
import os
import sys

import numpy as np
import pytest

from mdsuite_b200.database.hdf5_reader import H5File, HDF5FormatError

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from hdf5_writer import H5Writer  # noqa: E402


def _positions(rng, n, cfg):
    return (rng.random((n, cfg, 3)) * 20).astype(np.float32)


def test_mdsuite_shaped_database_roundtrip(tmp_path):
    rng = np.random.default_rng(1)
    na, cl = _positions(rng, 37, 120), _positions(rng, 41, 120)
    w = H5Writer()
    # h5py's guess_chunk splits every axis, including the last one
    w.create_dataset("Na/Positions", na, chunks=(10, 16, 3), maxshape=(37, None, 3), gzip=4)
    w.create_dataset("Cl/Positions", cl, chunks=(16, 32, 1), maxshape=(41, None, 3), gzip=9,
                     shuffle=True, continuation=True)
    w.create_dataset("Na/Velocities", na * 2, chunks=(37, 120, 3), maxshape=(37, None, 3), gzip=1,
                     filter_version=2)
    path = str(tmp_path / "database.hdf5")
    w.save(path)
    with H5File(path) as f:
        assert sorted(f.listdir("/")) == ["Cl", "Na"]
        assert sorted(f.listdir("/Na")) == ["Positions", "Velocities"]
        assert "Na/Positions" in f and "Na/Forces" not in f and "K" not in f
        assert sorted(f.walk()) == ["Cl/Positions", "Na/Positions", "Na/Velocities"]
        ds = f.dataset("Na/Positions")
        assert ds.shape == (37, 120, 3) and ds.maxshape == (37, None, 3)
        assert ds.dtype == np.dtype("<f4") and ds.chunks == (10, 16, 3) and ds.layout == "chunked"
        assert [fid for fid, _ in ds.filters] == [1]
        assert np.array_equal(ds.read(), na)
        box = ds.read_hyperslab((slice(3, 29), slice(15, 50), slice(0, 3)))
        assert np.array_equal(box, na[3:29, 15:50])
        cl_ds = f.dataset("Cl/Positions")
        assert [fid for fid, _ in cl_ds.filters] == [2, 1]
        assert np.array_equal(cl_ds.read(), cl)
        assert np.array_equal(cl_ds.read_hyperslab((slice(0, 41), slice(100, 120), slice(1, 2))),
                              cl[:, 100:120, 1:2])
        assert np.array_equal(f.dataset("Na/Velocities").read(), na * 2)


def test_many_chunks_many_groups_and_other_layouts(tmp_path):
    """> 64 chunks (two-level chunk B-tree), > 8 group members (several symbol-table nodes),
    contiguous and compact layouts, integer and float64 types, a 512-byte user block."""
    rng = np.random.default_rng(2)
    big = _positions(rng, 50, 90)
    w = H5Writer(user_block=512)
    w.create_dataset("big/Positions", big, chunks=(5, 9, 3), maxshape=(50, None, 3))  # 100 chunks
    names = [f"species_{k:02d}" for k in range(21)]
    for k, name in enumerate(names):
        w.create_dataset(f"{name}/Positions", big[:2, :3] + k, chunks=(2, 3, 3))
    ints = np.arange(24, dtype=np.int32).reshape(4, 6)
    w.create_dataset("misc/contiguous", ints, layout="contiguous")
    w.create_dataset("misc/compact", np.linspace(0, 1, 7), layout="compact")
    w.create_dataset("misc/unfiltered", big[:4, :5], chunks=(3, 2, 3), gzip=None)
    w.create_group("empty")
    path = str(tmp_path / "many.h5")
    w.save(path)
    with H5File(path) as f:
        assert sorted(f.listdir("/")) == sorted(names + ["big", "misc", "empty"])
        assert f.listdir("/empty") == []
        ds = f.dataset("big/Positions")
        assert len(ds.chunk_index()) == 100
        assert np.array_equal(ds.read(), big)
        for k, name in enumerate(names):
            assert np.array_equal(f.dataset(f"{name}/Positions").read(), big[:2, :3] + k)
        assert np.array_equal(f.dataset("misc/contiguous").read(), ints)
        assert f.dataset("misc/contiguous").dtype == np.dtype("<i4")
        assert np.array_equal(f.dataset("misc/compact").read(), np.linspace(0, 1, 7))
        assert np.array_equal(f.dataset("misc/unfiltered").read(), big[:4, :5])
        with pytest.raises(KeyError):
            f.dataset("misc")
        with pytest.raises(KeyError):
            f.dataset("nope/Positions")


def test_rejects_what_it_does_not_support(tmp_path):
    p = tmp_path / "not.h5"
    p.write_bytes(b"hello world" * 100)
    with pytest.raises(HDF5FormatError, match="signature"):
        H5File(str(p))
    raw = bytearray(H5Writer().tobytes())
    raw[8] = 2  # superblock version 2 (libver='latest')
    p.write_bytes(bytes(raw))
    with pytest.raises(HDF5FormatError, match="superblock version 2"):
        H5File(str(p))


def _scipy_matlab_hdf5():
    try:
        import scipy.io
    except ImportError:
        return None
    hits = glob.glob(os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data",
                                  "testhdf5*.mat"))
    return hits[0] if hits else None


@pytest.mark.skipif(_scipy_matlab_hdf5() is None, reason="SciPy's MATLAB v7.3 test file not found")
def test_foreign_file_written_by_matlab():
    """A file this package did not write: MATLAB 7.3 = HDF5 behind a 512-byte user block.  The
    superblock, root symbol table, local heap and object headers must parse, and every dataset
    that uses supported features must read."""
    with H5File(_scipy_matlab_hdf5()) as f:
        names = f.listdir("/")
        assert names, "root group is empty"
        seen = 0
        for name in names:
            try:
                if f.is_dataset(name):
                    arr = f.dataset(name).read()
                    assert arr.size >= 1
                    seen += 1
                    if name == "testdouble":  # SciPy's own test expects 0, pi/4, ..., 2 pi
                        assert np.allclose(arr.ravel(), np.arange(9) * np.pi / 4, rtol=1e-15)
            except HDF5FormatError:
                pass  # MATLAB-specific types (references, etc.)
        assert seen >= 1


def test_hdf5_database_feeds_frames_and_imports_into_the_frame_store(tmp_path):
    """The adapter serves frame-major batches from the atom-major gzip layout (what the RDF
    batcher asks for), and import_hdf5 re-chunks a reference database into the native store."""
    from mdsuite_b200.database.simulation_database import Database, HDF5Database, import_hdf5
    rng = np.random.default_rng(3)
    na, cl = _positions(rng, 30, 75), _positions(rng, 26, 75)
    w = H5Writer()
    w.create_dataset("Na/Positions", na, chunks=(8, 10, 3), maxshape=(30, None, 3))
    w.create_dataset("Cl/Positions", cl, chunks=(13, 32, 2), maxshape=(26, None, 3), shuffle=True)
    w.create_dataset("Na/Velocities", na, chunks=(30, 75, 3))
    path = str(tmp_path / "database.hdf5")
    w.save(path)
    db = HDF5Database(path)
    assert db.check_existence("Na/Positions") and not db.check_existence("K/Positions")
    assert db.get_data_size("Cl/Positions") == (26, 75, 26 * 75 * 3 * 4)
    for frames in (np.arange(0, 20), np.arange(17, 43), np.array([3, 50, 51, 74]), np.arange(70, 75)):
        got = db.load_frames("Na/Positions", frames)
        assert got.dtype == np.float32 and np.array_equal(got, na[:, frames].transpose(1, 0, 2))
    out = np.empty((5, 4, 3), dtype=np.float32)
    db.load_frames("Cl/Positions", np.arange(60, 65), out=out, atom_selection=[0, 5, 6, 25])
    assert np.array_equal(out, cl[[0, 5, 6, 25], 60:65].transpose(1, 0, 2))
    ref_shaped = db.load_data(["Na/Positions"], (np.s_[:], np.s_[10:14]))
    assert np.array_equal(ref_shaped["Na/Positions"], na[:, 10:14])
    db.close()

    store = Database(str(tmp_path / "store"))
    done = import_hdf5(path, store, frames_per_block=16)
    assert done == {"Cl/Positions": (26, 75), "Na/Positions": (30, 75)}
    assert np.array_equal(store.load_frames("Na/Positions", np.arange(75)), na.transpose(1, 0, 2))
    assert np.array_equal(store.load_frames("Cl/Positions", np.arange(40, 60)),
                          cl[:, 40:60].transpose(1, 0, 2))


def test_rdf_on_an_existing_reference_style_project(tmp_path, monkeypatch):
    """An experiment directory that holds a reference-style database.hdf5 (and no native store) is
    read in place: Project -> Experiment -> run.RadialDistributionFunction gives the same g(r) as
    the same positions ingested natively."""
    import mdsuite_b200 as mds
    from mdsuite_b200 import synthetic
    from mdsuite_b200.calculators.radial_distribution_function import RadialDistributionFunction
    from test_reference_golden import _OracleHandle
    monkeypatch.setattr(RadialDistributionFunction, "handle_factory", _OracleHandle)
    sysm = synthetic.nacl_melt(2, 10, seed=9)  # 64 atoms, 10 configurations
    n_na = sysm.counts[0]

    native = mds.Project("native", storage_path=str(tmp_path))
    exp_n = native.add_experiment("e", units="metal")
    exp_n.add_positions(sysm.positions, sysm.species, sysm.box)
    want = exp_n.run.RadialDistributionFunction(number_of_configurations=10, number_of_bins=40,
                                                cutoff=6.0, plot=False)

    ref = mds.Project("reference_style", storage_path=str(tmp_path))
    exp_dir = tmp_path / "reference_style" / "e"
    exp_dir.mkdir(parents=True)
    w = H5Writer()
    pad = np.zeros((sysm.n_atoms, 6, 3), dtype=np.float32)  # h5py resizes in blocks: trailing zeros
    am = np.concatenate([sysm.positions.transpose(1, 0, 2), pad], axis=1)
    w.create_dataset("Na/Positions", am[:n_na], chunks=(8, 4, 3), maxshape=(n_na, None, 3))
    w.create_dataset("Cl/Positions", am[n_na:], chunks=(16, 8, 1), maxshape=(n_na, None, 3))
    w.save(str(exp_dir / "database.hdf5"))
    ref.database.add_experiment("e")
    for key, value in (("number_of_configurations", 10), ("box_array", [float(v) for v in sysm.box]),
                       ("number_of_atoms", sysm.n_atoms), ("version", 1)):
        ref.database.set_attribute("e", key, value)
    ref.database.set_species("e", {n: {"n_particles": c, "mass": None, "charge": None,
                                       "properties": ["Positions"]} for n, c in sysm.species.items()})
    exp_r = mds.Project("reference_style", storage_path=str(tmp_path)).experiments.e
    assert type(exp_r.database).__name__ == "HDF5Database"
    got = exp_r.run.RadialDistributionFunction(number_of_configurations=10, number_of_bins=40,
                                               cutoff=6.0, plot=False)
    for key in want.keys():
        assert np.array_equal(np.array(got[key]["y"])[1:], np.array(want[key]["y"])[1:])


def test_native_chunk_decoder_equals_the_python_reader(tmp_path):
    """mds_h5_read_chunks (threads + the library's own inflate + strided placement) against the
    pure-Python reader on every filter pipeline the writer can produce, for random boxes and for
    atom-major, frame-major and padded output layouts."""
    rng = np.random.default_rng(5)
    a = _positions(rng, 37, 120)
    b = (rng.random((23, 50, 3, 2)) * 9).astype(np.float64)   # rank 4, 8-byte elements
    c = rng.integers(-500, 500, (301,), dtype=np.int32)       # rank 1
    w = H5Writer()
    w.create_dataset("g/plain", a, chunks=(10, 16, 3), maxshape=(37, None, 3), gzip=4)
    w.create_dataset("g/shuffled", a, chunks=(16, 32, 1), maxshape=(37, None, 3), gzip=9, shuffle=True)
    w.create_dataset("g/one_frame_chunk", a, chunks=(8, 120, 3), gzip=1)
    w.create_dataset("g/unfiltered", a, chunks=(37, 7, 2), gzip=None)
    w.create_dataset("g/rank4", b, chunks=(5, 9, 2, 2), gzip=6, shuffle=True)
    w.create_dataset("g/rank1", c, chunks=(64,), gzip=6)
    path = str(tmp_path / "native.hdf5")
    w.save(path)
    with H5File(path) as f:
        for name, ref in (("g/plain", a), ("g/shuffled", a), ("g/one_frame_chunk", a),
                          ("g/unfiltered", a), ("g/rank4", b), ("g/rank1", c)):
            ds = f.dataset(name)
            for _ in range(12):
                lo = [int(rng.integers(0, n)) for n in ref.shape]
                hi = [int(rng.integers(l + 1, n + 1)) for l, n in zip(lo, ref.shape)]
                shape = [h - l for l, h in zip(lo, hi)]
                want = ref[tuple(slice(l, h) for l, h in zip(lo, hi))]
                # C-contiguous output
                out = np.zeros(shape, dtype=ref.dtype)
                strides = [int(np.prod(shape[d + 1:])) for d in range(len(shape))]
                assert ds.read_box_into(lo, hi, out, strides, threads=3)
                assert np.array_equal(out, want)
                assert np.array_equal(out, ds.read_hyperslab(tuple(slice(l, h) for l, h in zip(lo, hi))))
                if len(shape) >= 2:
                    # first two axes swapped (atom-major file -> frame-major batch), padded rows
                    perm = [1, 0] + list(range(2, len(shape)))
                    t_shape = [shape[p] for p in perm]
                    t_shape[1] += 3
                    buf = np.full(t_shape, -7, dtype=ref.dtype)
                    t_strides = [int(np.prod(t_shape[d + 1:])) for d in range(len(t_shape))]
                    ds_strides = [t_strides[1], t_strides[0]] + t_strides[2:]
                    assert ds.read_box_into(lo, hi, buf, ds_strides, threads=2)
                    assert np.array_equal(buf[:, :shape[0]], np.moveaxis(want, 0, 1))
                    assert (buf[:, shape[0]:] == -7).all()
    # the database adapter: native and Python paths give the same frames
    from mdsuite_b200.database.simulation_database import HDF5Database
    w2 = H5Writer()
    w2.create_dataset("Na/Positions", a, chunks=(10, 16, 3), maxshape=(37, None, 3), gzip=4)
    p2 = str(tmp_path / "database.hdf5")
    w2.save(p2)
    frames = np.array([3, 4, 5, 40, 41, 119])
    got = {}
    for native in (True, False):
        db = HDF5Database(p2)
        db.native = native
        got[native] = [db.load_frames("Na/Positions", np.arange(10, 50)),
                       db.load_frames("Na/Positions", frames, atom_selection=[0, 5, 36])]
        db.close()
    assert np.array_equal(got[True][0], a[:, 10:50].transpose(1, 0, 2))
    assert np.array_equal(got[True][1], a[[0, 5, 36]][:, frames].transpose(1, 0, 2))
    assert all(np.array_equal(x, y) for x, y in zip(got[True], got[False]))


def test_reference_written_file():
    """A database.hdf5 written by h5py / libhdf5 the way the reference writes it
    (tests/golden/make_hdf5_golden.py).  The build image has no h5py, so the file may not exist
    yet: the reader is then UNPINNED against libhdf5 and this test says so by skipping."""
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_database.hdf5")
    if not os.path.exists(golden):
        pytest.skip("tests/golden/reference_database.hdf5 has not been generated (needs h5py): "
                    "HDF5 reader unpinned against libhdf5")
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "make_hdf5_golden", os.path.join(os.path.dirname(golden), "make_hdf5_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from mdsuite_b200.database.simulation_database import HDF5Database
    arrays = mod.golden_arrays()
    with H5File(golden) as f:
        for name, arr in arrays.items():
            assert np.array_equal(f.dataset(name).read(), arr)
    for native in (True, False):
        db = HDF5Database(golden)
        db.native = native
        for name, arr in arrays.items():
            got = db.load_frames(name, np.arange(7, 131))
            assert np.array_equal(got, arr[:, 7:131].transpose(1, 0, 2))
        db.close()
