"""CPU tests of the host-side mirror of the reference interface: Project -> Experiment ->
run.RadialDistributionFunction -> Computation, with the GPU handle replaced by a test double that
answers from the oracle (tests are the only place allowed to do that)."""
import json
import os
import sqlite3

import numpy as np
import pytest

import mdsuite_b200 as mds
from mdsuite_b200 import synthetic
from mdsuite_b200.calculators.radial_distribution_function import RadialDistributionFunction
from mdsuite_b200.memory_management.memory_manager import MemoryManager
from oracle import c_oracle, rdf_oracle


class OracleHandle:
    """Stands in for _cuda.RDFHandle in host-logic tests."""
    instances = []

    def __init__(self, counts, nbins, cutoff, box, parity=True, device=0, path=None, **kw):
        self.args = (list(counts), nbins, cutoff, np.asarray(box), parity)
        self.n_pairs = len(counts) * (len(counts) + 1) // 2
        self.nbins = nbins
        self.device = device
        self.total = np.zeros((self.n_pairs, nbins), dtype=np.uint64)
        self.frames = 0
        self.evaluated = 0
        self.submits = 0
        OracleHandle.instances.append(self)

    def submit(self, positions, layout=0, pinned=False):
        pos = positions.numpy() if hasattr(positions, "numpy") else positions
        counts, nbins, cutoff, box, parity = self.args
        c, ev = c_oracle.rdf_counts_c(np.array(pos), counts, box, cutoff, nbins, skip_first=parity,
                                      layout=c_oracle.FRAME_MAJOR)
        self.total += c.astype(np.uint64)
        self.frames += pos.shape[0]
        self.evaluated += ev
        self.submits += 1

    def wait_host_copies(self):
        pass

    def counts(self):
        return self.total.copy()

    def stats(self):
        from types import SimpleNamespace
        return SimpleNamespace(pairs_evaluated=float(self.evaluated), frames_general_minimage=0)

    def close(self):
        pass


@pytest.fixture
def project(tmp_path, monkeypatch):
    OracleHandle.instances = []
    monkeypatch.setattr(RadialDistributionFunction, "handle_factory", OracleHandle)
    sysm = synthetic.nacl_melt(2, 12, seed=5)  # 64 atoms, 12 configurations
    dump = tmp_path / "nacl.lammpstraj"
    synthetic.write_lammps_dump(sysm, str(dump))
    p = mds.Project("proj", storage_path=str(tmp_path))
    p.add_experiment("NaCl", timestep=0.002, temperature=1400.0, units="metal",
                     simulation_data=str(dump))
    return p, sysm, tmp_path


def test_rdf_from_experiment_matches_oracle_normalisation(project):
    p, sysm, tmp = project
    comp = p.experiments.NaCl.run.RadialDistributionFunction(
        number_of_configurations=12, number_of_bins=50, cutoff=6.0, plot=True, frames_per_batch=5)
    assert comp.keys() == ["Na_Na", "Na_Cl", "Cl_Cl"]
    counts, _ = c_oracle.rdf_counts_c(sysm.positions, sysm.counts, sysm.box, 6.0, 50,
                                      layout=c_oracle.FRAME_MAJOR)
    want = rdf_oracle.normalise(counts, ["Na", "Cl"], sysm.counts, sysm.box, 6.0, 12,
                                length_unit=1e-10)
    for key in comp.keys():
        got_y, want_y = np.array(comp[key]["y"]), np.array(want[key]["y"])
        assert np.array_equal(np.isnan(got_y), np.isnan(want_y))
        ok = np.isfinite(want_y)
        assert np.allclose(got_y[ok], want_y[ok], rtol=1e-12)
        assert np.allclose(comp[key]["x"], want[key]["x"], rtol=1e-15)
    assert OracleHandle.instances[0].submits == 3  # 12 frames in batches of 5
    assert os.path.exists(tmp / "proj" / "NaCl" / "figures" / "Radial_Distribution_Function.html")
    assert comp.computation_parameter["number_of_bins"] == 50
    assert comp.computation_parameter["cutoff"] == 6.0
    assert comp.data_range == 1


def test_defaults_cache_and_version_invalidation(project):
    p, sysm, _ = project
    exp = p.experiments.NaCl
    c1 = exp.run.RadialDistributionFunction(number_of_configurations=6, plot=False)
    # user args are stored as given (None), check_input defaults do not leak into the cache key
    assert c1.computation_parameter["cutoff"] is None and c1.computation_parameter["stop"] is None
    n_bins_default = int((sysm.box[0] / 2 - 0.1) / 0.01)
    assert len(c1["Na_Cl"]["x"]) == n_bins_default
    n_handles = len(OracleHandle.instances)
    c2 = exp.run.RadialDistributionFunction(number_of_configurations=6, plot=False)
    assert len(OracleHandle.instances) == n_handles and c2.id == c1.id  # cache hit, nothing ran
    c3 = exp.run.RadialDistributionFunction(number_of_configurations=6, number_of_bins=40, plot=False)
    assert c3.id != c1.id and len(OracleHandle.instances) == n_handles + 1
    exp.add_positions(sysm.positions[:2], sysm.species, sysm.box, name="more")  # version += 1
    assert exp.number_of_configurations == 14
    c4 = exp.run.RadialDistributionFunction(number_of_configurations=6, plot=False)
    assert c4.id != c1.id


def test_project_run_returns_dict_and_sql_layout(project):
    p, sysm, tmp = project
    p.add_experiment("NaCl_b", units="real")
    p.experiments.NaCl_b.add_positions(sysm.positions[:4], sysm.species, sysm.box)
    out = p.run.RadialDistributionFunction(number_of_configurations=4, cutoff=5.0,
                                           number_of_bins=25, plot=False)
    assert sorted(out) == ["NaCl", "NaCl_b"] and out["NaCl_b"].keys() == ["Na_Na", "Na_Cl", "Cl_Cl"]
    con = sqlite3.connect(str(tmp / "proj" / "project.db"))
    tables = {r[0] for r in con.execute("SELECT name FROM sqlite_master WHERE type='table'")}
    assert {"projects", "experiments", "experiment_attributes", "experiment_species",
            "computations", "computation_attributes", "computation_results",
            "species_association"} <= tables
    rows = con.execute("SELECT data FROM computation_results").fetchall()
    d = json.loads(rows[0][0])
    assert set(d) == {"x", "y"} and len(d["x"]) == 25
    assoc = con.execute("SELECT count FROM species_association ORDER BY rowid").fetchall()
    assert [a[0] for a in assoc[:4]] == [2, 1, 1, 2]  # Na_Na -> one row count 2; Na_Cl -> two rows
    attr = dict(con.execute("SELECT name, data FROM computation_attributes WHERE computation_id=1"))
    assert json.loads(attr["number_of_configurations"]) == {"serialized_value": 4}
    assert json.loads(attr["atom_selection"]) == {"serialized_value": "slice(None, None, None)"}


def test_atom_selection_dict_and_species_subset(project):
    p, sysm, _ = project
    exp = p.experiments.NaCl
    sel = {"Na": [0, 1, 2, 5, 7], "Cl": [3, 4, 30]}
    comp = exp.run.RadialDistributionFunction(number_of_configurations=12, cutoff=6.0,
                                              number_of_bins=30, atom_selection=sel, plot=False)
    sub = np.concatenate([sysm.positions[:, sel["Na"]], sysm.positions[:, 32 + np.array(sel["Cl"])]], 1)
    counts, _ = c_oracle.rdf_counts_c(sub, [5, 3], sysm.box, 6.0, 30, layout=c_oracle.FRAME_MAJOR)
    want = rdf_oracle.normalise(counts, ["Na", "Cl"], [5, 3], sysm.box, 6.0, 12)
    ok = np.isfinite(np.array(want["Na_Cl"]["y"]))
    assert np.allclose(np.array(comp["Na_Cl"]["y"])[ok], np.array(want["Na_Cl"]["y"])[ok], rtol=1e-12)
    only_cl = exp.run.RadialDistributionFunction(number_of_configurations=12, cutoff=6.0,
                                                 number_of_bins=30, species=["Cl"], plot=False)
    assert only_cl.keys() == ["Cl_Cl"]


def test_too_many_configurations_is_a_clear_error(project):
    p, _, _ = project
    with pytest.raises(ValueError, match="exceeds"):
        p.experiments.NaCl.run.RadialDistributionFunction(number_of_configurations=500, plot=False)
    # nothing was persisted for the failed run (reference calculator.py:123-126)
    con = sqlite3.connect(p.database.path)
    assert con.execute("SELECT COUNT(*) FROM computations").fetchone()[0] == 0


def test_corrected_mode_ideal_gas_is_flat(tmp_path, monkeypatch):
    monkeypatch.setattr(RadialDistributionFunction, "handle_factory", OracleHandle)
    sysm = synthetic.ideal_gas(400, 10, seed=9, box_length=20.0)
    p = mds.Project("gas", storage_path=str(tmp_path))
    e = p.add_experiment("gas")
    e.add_positions(sysm.positions, sysm.species, sysm.box)
    comp = e.run.RadialDistributionFunction(number_of_configurations=10, cutoff=9.0,
                                            number_of_bins=18, plot=False, parity=False)
    y = np.array(comp["X_X"]["y"])[2:]
    # quirk Q2: the reference normalises bin k = [k dr, (k+1) dr) with a shell at
    # linspace(0, cutoff, B)[k] = k cutoff / (B-1); undo that to compare with the exact shell volume
    r = np.linspace(0, 9.0, 18)[2:]
    dr = 9.0 / 18
    lo = np.arange(18)[2:] * dr
    shell_exact = 4 * np.pi / 3 * ((lo + dr) ** 3 - lo ** 3)
    shell_ref = 4 * np.pi * r ** 2 * dr
    g = y * shell_ref / shell_exact * (400 / 399)
    assert np.all(np.abs(g - 1.0) < 0.08)


def test_memory_manager_batch_size_contract():
    class FakeDB:
        def get_data_size(self, path):
            return 1000, 100, 1000 * 100 * 12  # (n_particles, n_configs, n_bytes)

    mm = MemoryManager(data_path=["Na/Positions", "Cl/Positions"], database=FakeDB(),
                       memory_fraction=0.5, free_memory=10 ** 6)
    batch, n_batches, remainder = mm.get_batch_size()
    from mdsuite_b200.memory_management.memory_manager import GPU_LINEAR_SCALE
    per_conf = 2 * 12000 * GPU_LINEAR_SCALE
    assert batch == int(0.5 * 10 ** 6 / per_conf) and (n_batches, remainder) == divmod(100, batch)
    mm = MemoryManager(data_path=["Na/Positions"], database=FakeDB(), free_memory=10 ** 12)
    assert mm.get_batch_size() == (100, 1, 0)  # clipped to the number of configurations
    with pytest.raises(ValueError):
        MemoryManager(database=FakeDB(), free_memory=1).get_batch_size()


def test_molecules_true_reads_molecule_datasets(project, tmp_path):
    """molecules=True (reference :245-249, :701-705): species are the molecule names whose
    <molecule>/Positions datasets a MolecularMap run left in the store; same kernel, the molecule
    table supplies particles_list and the prefactor counts."""
    from mdsuite_b200.database.simulation_database import (PropertyInfo, SpeciesInfo,
                                                           TrajectoryChunkData)
    p, sysm, _ = project
    exp = p.experiments.NaCl
    n_mol = sysm.counts[0]
    mol = 0.5 * (sysm.positions[:, :n_mol] + sysm.positions[:, n_mol:])  # one "NaCl" site per ion pair
    info = SpeciesInfo("NaCl", n_mol, [PropertyInfo("Positions", 3)])
    exp.database.add_dataset("NaCl/Positions", n_mol, 3)
    chunk = TrajectoryChunkData([info], mol.shape[0])
    chunk.add_data(mol.astype(np.float32), 0, "NaCl", "Positions")
    exp.database.add_data(chunk)
    p.database.set_species("NaCl", {"NaCl": {"n_particles": n_mol, "mass": None, "charge": None,
                                              "properties": ["Positions"]}}, molecule=True)
    comp = exp.run.RadialDistributionFunction(number_of_configurations=12, number_of_bins=30,
                                              cutoff=6.0, plot=False, molecules=True)
    assert comp.keys() == ["NaCl_NaCl"]
    counts, _ = c_oracle.rdf_counts_c(mol.astype(np.float32), [n_mol], sysm.box, 6.0, 30,
                                      layout=c_oracle.FRAME_MAJOR)
    want = rdf_oracle.normalise(counts, ["NaCl"], [n_mol], sysm.box, 6.0, 12, length_unit=1e-10)
    assert np.allclose(np.array(comp["NaCl_NaCl"]["y"])[1:], np.array(want["NaCl_NaCl"]["y"])[1:],
                       rtol=1e-12)


def test_unit_systems_equal_the_reference():
    """tests/golden/units_golden.json is the reference's units_dict (make_units_golden.py)."""
    import dataclasses
    from mdsuite_b200.calculators import rdf_postprocessing as post
    from mdsuite_b200.utils import units
    with open(os.path.join(os.path.dirname(__file__), "golden", "units_golden.json")) as fh:
        want = json.load(fh)
    assert sorted(units.units_dict) == sorted(want["systems"])
    for name, ref in want["systems"].items():
        mine = dict(dataclasses.asdict(units.units_dict[name]), volume=units.units_dict[name].volume)
        assert mine == ref, name
    assert post.BOLTZMANN_SI == want["boltzmann_constant"]
    assert post.GOLDEN_RATIO == want["golden_ratio"]
