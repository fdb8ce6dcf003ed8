"""GPU end-to-end: LAMMPS dump -> Project/Experiment -> run.RadialDistributionFunction on the B200,
against the oracle's histograms and the reference normalisation (configs[0] of BASELINE.json)."""
import numpy as np
import pytest

import mdsuite_b200 as mds
from mdsuite_b200 import synthetic
from oracle import c_oracle, rdf_oracle

pytestmark = pytest.mark.gpu


def test_cfg1_nacl_1000_atoms_100_configurations_through_the_calculator(tmp_path):
    sysm = synthetic.nacl_melt(5, 100, seed=1)
    dump = tmp_path / "nacl_1000.lammpstraj"
    synthetic.write_lammps_dump(sysm, str(dump))
    project = mds.Project("cfg1", storage_path=str(tmp_path))
    project.add_experiment("NaCl", timestep=0.002, temperature=1400.0, units="metal",
                           simulation_data=str(dump))
    exp = project.experiments.NaCl
    calc_out = exp.run.RadialDistributionFunction(number_of_configurations=100, number_of_bins=500,
                                                  cutoff=10.0, plot=True, frames_per_batch=32)
    counts, evaluated = c_oracle.rdf_counts_c(sysm.positions, sysm.counts, sysm.box, 10.0, 500,
                                              layout=c_oracle.FRAME_MAJOR)
    want = rdf_oracle.normalise(counts, ["Na", "Cl"], sysm.counts, sysm.box, 10.0, 100)
    assert calc_out.keys() == ["Na_Na", "Na_Cl", "Cl_Cl"]
    for key in calc_out.keys():
        got_y, want_y = np.array(calc_out[key]["y"]), np.array(want[key]["y"])
        assert np.isnan(got_y[0]) or np.isinf(got_y[0])
        assert np.array_equal(np.isfinite(got_y), np.isfinite(want_y))
        # north star: g(r) within 1e-6 relative for indices >= 1 (integer counts are bit-exact, so
        # this is float64 normalisation noise only)
        assert np.allclose(got_y[1:], want_y[1:], rtol=1e-6, atol=0)
        assert np.allclose(calc_out[key]["x"], want[key]["x"], rtol=1e-15)
    # default arguments: cutoff = L/2 - 0.1, bins = int(cutoff / 0.01), every stored configuration
    second = exp.run.RadialDistributionFunction(number_of_configurations=50, plot=False)
    cutoff = sysm.box[0] / 2 - 0.1
    nbins = int(cutoff / 0.01)
    frames = np.linspace(0, 99, 50, dtype=int)
    counts, _ = c_oracle.rdf_counts_c(sysm.positions[frames], sysm.counts, sysm.box, cutoff, nbins,
                                      layout=c_oracle.FRAME_MAJOR)
    want = rdf_oracle.normalise(counts, ["Na", "Cl"], sysm.counts, sysm.box, cutoff, 50)
    assert np.allclose(np.array(second["Cl_Cl"]["y"])[1:], np.array(want["Cl_Cl"]["y"])[1:], rtol=1e-6)


def test_corrected_mode_ideal_gas_g_of_r_is_one(tmp_path):
    """cfg5 acceptance: uniform random points, parity=False, g(r) = 1 within Poisson noise."""
    sysm = synthetic.ideal_gas(20000, 8, seed=5, box_length=100.0)
    project = mds.Project("cfg5", storage_path=str(tmp_path))
    exp = project.add_experiment("gas")
    exp.add_positions(sysm.positions, sysm.species, sysm.box)
    nbins, cutoff = 200, 25.0
    comp = exp.run.RadialDistributionFunction(number_of_configurations=8, number_of_bins=nbins,
                                              cutoff=cutoff, plot=False, parity=False)
    y = np.array(comp["X_X"]["y"])
    k = np.arange(nbins)
    dr = cutoff / nbins
    shell_exact = 4 * np.pi / 3 * (((k + 1) * dr) ** 3 - (k * dr) ** 3)
    shell_ref = 4 * np.pi * np.linspace(0, cutoff, nbins) ** 2 * dr
    n = 20000
    g = y[5:] * shell_ref[5:] / shell_exact[5:] * (n / (n - 1))
    expected_pairs = 8 * n * (n - 1) / 2 * shell_exact[5:] / 100.0 ** 3
    sigma = 1.0 / np.sqrt(expected_pairs)
    assert np.all(np.abs(g - 1.0) < 5 * sigma + 1e-4)


def test_cell_list_path_through_the_calculator_matches_all_pairs(tmp_path):
    """A system with cutoff << box: the library picks the cell-list kernel on its own; the result
    must equal the forced all-pairs run bit for bit (same integer counts -> same g(r))."""
    sysm = synthetic.lj_fluid(20, 6, seed=8)  # 8000 atoms, L = 21.2
    project = mds.Project("cells", storage_path=str(tmp_path))
    exp = project.add_experiment("lj")
    exp.add_positions(sysm.positions, sysm.species, sysm.box)
    exp2 = project.add_experiment("lj_allpairs")  # separate experiment: the results cache keys on
    exp2.add_positions(sysm.positions, sysm.species, sysm.box)  # the reference's Args only
    auto = exp.run.RadialDistributionFunction(number_of_configurations=6, number_of_bins=300,
                                              cutoff=3.0, plot=False)
    forced = exp2.run.RadialDistributionFunction(number_of_configurations=6, number_of_bins=300,
                                                 cutoff=3.0, plot=False, path="allpairs")
    counts, _ = c_oracle.rdf_counts_c(sysm.positions, sysm.counts, sysm.box, 3.0, 300,
                                      layout=c_oracle.FRAME_MAJOR)
    want = rdf_oracle.normalise(counts, ["Ar"], sysm.counts, sysm.box, 3.0, 6)
    for out in (auto, forced):
        y = np.array(out["Ar_Ar"]["y"])
        assert np.allclose(y[1:], np.array(want["Ar_Ar"]["y"])[1:], rtol=1e-6, atol=0)
    assert np.array_equal(np.array(auto["Ar_Ar"]["y"])[1:], np.array(forced["Ar_Ar"]["y"])[1:])


def test_cfg3_scale_through_the_frame_store_and_batcher(tmp_path):
    """97k-atom LJ frames from the trajectory store through FrameBatcher (pinned ring, several
    batches) into the cell-list kernel; the forced all-pairs run on the same store must agree."""
    sysm = synthetic.lj_fluid(46, 6, seed=3)
    project = mds.Project("cfg3", storage_path=str(tmp_path))
    exp = project.add_experiment("lj", units="real")
    exp.add_positions(sysm.positions, sysm.species, sysm.box)
    exp2 = project.add_experiment("lj_allpairs", units="real")
    exp2.add_positions(sysm.positions[:2], sysm.species, sysm.box)
    cells = exp.run.RadialDistributionFunction(number_of_configurations=6, number_of_bins=300,
                                               cutoff=3.0, plot=False, frames_per_batch=4)
    first2 = exp.run.RadialDistributionFunction(number_of_configurations=2, stop=1, number_of_bins=300,
                                                cutoff=3.0, plot=False)
    allp = exp2.run.RadialDistributionFunction(number_of_configurations=2, number_of_bins=300,
                                               cutoff=3.0, plot=False, path="allpairs")
    assert np.array_equal(np.array(first2["Ar_Ar"]["y"])[1:], np.array(allp["Ar_Ar"]["y"])[1:])
    y = np.array(cells["Ar_Ar"]["y"])
    assert np.isfinite(y[1:]).all() and y[1:].max() > 1.0
