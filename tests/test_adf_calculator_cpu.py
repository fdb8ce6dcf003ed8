"""Host side of the ADF calculator (Project -> Experiment -> run.AngularDistributionFunction ->
Computation) with the GPU engine replaced by a test double that answers from the oracle, against
tests/golden/adf_golden.json — the output of the reference's own run_calculator: same batches, same
curves (float32 summation noise only), same axis and max_peak."""
import json
import os

import numpy as np
import pytest

import mdsuite_b200 as mds
from mdsuite_b200.calculators import angular_distribution_function as adf_calc
from mdsuite_b200.calculators.angular_distribution_function import AngularDistributionFunction
from mdsuite_b200.file_io.script_input import ScriptInput
from oracle import adf_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "adf_golden.json")) as _fh:
    GOLDEN = json.load(_fh)["cases"]


class OracleADFHandle:
    """Stands in for _cuda.adf.ADFHandle in host-logic tests."""
    calls = []

    def __init__(self, counts, nbins, cutoff, box, norm_power=4, **kw):
        self.cfg = (list(counts), int(nbins), float(cutoff), np.asarray(box, dtype=np.float32),
                    int(norm_power))

    def compute(self, positions):
        counts, nbins, cutoff, box, power = self.cfg
        OracleADFHandle.calls.append(positions.shape[0])
        return np.stack([adf_oracle.raw_histograms(f, counts, box, cutoff, nbins, power)
                         for f in positions])

    def close(self):
        pass


def positions_for(case):
    rng = np.random.Generator(np.random.Philox(case["seed"]))
    n = sum(case["counts"])
    box = np.asarray(case["box"], dtype=np.float64)
    pos = rng.random((case["frames"], n, 3)) * box
    if case.get("unwrapped"):
        pos += rng.integers(-2, 3, size=pos.shape) * box
    return pos.astype(np.float32)


def make_experiment(tmp_path, case):
    from mdsuite_b200.database.simulation_database import (PropertyInfo, SpeciesInfo,
                                                           TrajectoryChunkData, TrajectoryMetadata)
    pos = positions_for(case)
    species_list = [SpeciesInfo(name=n, n_particles=c, properties=[PropertyInfo("Positions", 3)])
                    for n, c in zip(case["names"], case["counts"])]
    md = TrajectoryMetadata(n_configurations=case["frames"], species_list=species_list,
                            box_l=list(case["box"]), sample_rate=1)
    chunk = TrajectoryChunkData(species_list, case["frames"])
    off = 0
    for n, c in zip(case["names"], case["counts"]):
        chunk.add_data(pos[:, off:off + c], 0, n, "Positions")
        off += c
    os.makedirs(str(tmp_path), exist_ok=True)
    project = mds.Project("adf_" + case["name"], storage_path=str(tmp_path))
    exp = project.add_experiment("run", timestep=0.002, temperature=300.0, units="metal",
                                 simulation_data=ScriptInput(data=chunk, metadata=md, name="mem"))
    return project, exp


@pytest.mark.parametrize("case", GOLDEN, ids=[c["name"] for c in GOLDEN])
def test_calculator_equals_the_reference_run(tmp_path, monkeypatch, case):
    monkeypatch.setattr(AngularDistributionFunction, "handle_factory", OracleADFHandle)
    OracleADFHandle.calls = []
    project, exp = make_experiment(tmp_path, case)
    comp = exp.run.AngularDistributionFunction(
        number_of_configurations=case["number_of_configurations"], cutoff=case["cutoff"],
        start=case.get("start", 1), number_of_bins=case["nbins"], norm_power=case["norm_power"],
        batches=case["n_batches"], plot=False)
    assert OracleADFHandle.calls == [len(b) for b in case["batches"]]
    assert comp.keys() == ["_".join(r["subjects"]) for r in case["results"]]
    for rec in case["results"]:
        got = comp["_".join(rec["subjects"])]
        angle = np.array([float.fromhex(v) for v in rec["angle"]])
        want = np.array([float.fromhex(v) for v in rec["adf"]], dtype=np.float32)
        assert np.array_equal(np.array(got["angle"]), angle)
        assert got["max_peak"] == rec["max_peak"]
        mine = np.array(got["adf"])
        if case["norm_power"] == 0:
            assert np.array_equal(mine.astype(np.float32), want)
        else:
            assert np.max(np.abs(mine - want)) <= 2e-6 * want.max()
    # cached: the second call with the same arguments does not compute again
    OracleADFHandle.calls = []
    again = exp.run.AngularDistributionFunction(
        number_of_configurations=case["number_of_configurations"], cutoff=case["cutoff"],
        start=case.get("start", 1), number_of_bins=case["nbins"], norm_power=case["norm_power"],
        batches=case["n_batches"], plot=False)
    assert OracleADFHandle.calls == [] and again.keys() == comp.keys()


def test_normalise_batch_is_numpy_density():
    rng = np.random.Generator(np.random.Philox(3))
    theta = rng.uniform(0, np.pi, 5000).astype(np.float32)
    w = rng.uniform(0.1, 2.0, 5000).astype(np.float32)
    nbins = 64
    raw = np.histogram(theta, bins=nbins, range=adf_calc.BIN_RANGE, weights=w.astype(np.float64))[0]
    want = np.histogram(theta, bins=nbins, range=adf_calc.BIN_RANGE, weights=w, density=True)[0]
    got = adf_calc.normalise_batch(raw, nbins)
    assert got.dtype == np.float32
    assert np.max(np.abs(got - want.astype(np.float32))) <= 1e-6 * want.max()
    empty = adf_calc.normalise_batch(np.zeros(nbins), nbins)
    assert np.all(np.isnan(empty))  # np.histogram(density=True) of nothing is 0/0, as in the reference


def test_no_gpu_means_an_error_not_a_fallback(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    project, exp = make_experiment(tmp_path, GOLDEN[0])
    with pytest.raises(RuntimeError, match="no CUDA device"):
        exp.run.AngularDistributionFunction(number_of_configurations=2, cutoff=3.0, plot=False)


def test_atom_selection_and_species_subset(tmp_path, monkeypatch):
    """A dict atom_selection picks atoms per species (reference :218-301); the engine sees only the
    selected atoms, species in the order of ``species=``."""
    monkeypatch.setattr(AngularDistributionFunction, "handle_factory", OracleADFHandle)
    case = GOLDEN[1]  # three species: O 6, H 9, C 5
    project, exp = make_experiment(tmp_path, case)
    selection = {"H": [0, 2, 4, 6, 8], "O": [1, 2, 3]}
    comp = exp.run.AngularDistributionFunction(
        number_of_configurations=2, cutoff=3.0, start=0, number_of_bins=24, norm_power=0,
        species=["H", "O"], atom_selection=selection, batches=1, plot=False)
    assert comp.keys() == ["H_H_H", "H_H_O", "H_O_O", "O_O_O"]
    pos = positions_for(case)
    frames = np.linspace(0, case["frames"] - 1, 2, dtype=int)
    sub = np.concatenate([pos[:, 6:15][:, selection["H"]], pos[:, 0:6][:, selection["O"]]], axis=1)
    want = adf_oracle.adf(sub, [5, 3], ["H", "O"], case["box"], 3.0, 24, 0, [list(frames)])
    for key in comp.keys():
        ref = want[key.replace("_", "-")]["adf"]
        got = np.array(comp[key]["adf"], dtype=np.float32)
        assert np.array_equal(np.nan_to_num(got, nan=-1.0), np.nan_to_num(ref, nan=-1.0))


def test_molecules_true_reads_molecule_datasets(tmp_path, monkeypatch):
    """molecules=True (reference calculators/angular_distribution_function.py:154, :234-235):
    species are the molecule names whose <molecule>/Positions datasets a MolecularMap run left in
    the store; same engine, the molecule table supplies the index ranges."""
    from mdsuite_b200.database.simulation_database import (PropertyInfo, SpeciesInfo,
                                                           TrajectoryChunkData)
    monkeypatch.setattr(AngularDistributionFunction, "handle_factory", OracleADFHandle)
    case = GOLDEN[1]  # O 6, H 9, C 5
    project, exp = make_experiment(tmp_path, case)
    pos = positions_for(case)
    # two kinds of "molecules": OH sites (midpoints of the first six O-H pairs) and the C atoms
    oh = (0.5 * (pos[:, 0:6] + pos[:, 6:12])).astype(np.float32)
    c = pos[:, 15:20]
    table = {}
    for name, arr in (("OH", oh), ("C1", c)):
        info = SpeciesInfo(name, arr.shape[1], [PropertyInfo("Positions", 3)])
        exp.database.add_dataset(f"{name}/Positions", arr.shape[1], 3)
        chunk = TrajectoryChunkData([info], arr.shape[0])
        chunk.add_data(arr, 0, name, "Positions")
        exp.database.add_data(chunk)
        table[name] = {"n_particles": arr.shape[1], "mass": None, "charge": None,
                       "properties": ["Positions"]}
    project.database.set_species("run", table, molecule=True)
    comp = exp.run.AngularDistributionFunction(
        number_of_configurations=2, cutoff=4.0, start=0, number_of_bins=24, norm_power=0,
        molecules=True, batches=1, plot=False)
    assert comp.keys() == ["OH_OH_OH", "OH_OH_C1", "OH_C1_C1", "C1_C1_C1"]
    frames = np.linspace(0, case["frames"] - 1, 2, dtype=int)
    sub = np.concatenate([oh, c], axis=1)
    want = adf_oracle.adf(sub, [6, 5], ["OH", "C1"], case["box"], 4.0, 24, 0, [list(frames)])
    for key in comp.keys():
        ref = want[key.replace("_", "-")]["adf"]
        got = np.array(comp[key]["adf"], dtype=np.float32)
        assert np.array_equal(np.nan_to_num(got, nan=-1.0), np.nan_to_num(ref, nan=-1.0))
