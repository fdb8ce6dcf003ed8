"""The three calculators that consume RDF output against golden vectors
produced by the REFERENCE'S OWN CODE (tests/golden/make_consumer_golden.py executes the reference's
CoordinationNumbers / PotentialOfMeanForce / KirkwoodBuffIntegral unmodified; they are NumPy/SciPy
only, so nothing is restated).  Series within 1e-12 relative; scalars likewise."""
import json
import os
import sys

import numpy as np
import pytest

import mdsuite_b200 as mds
from mdsuite_b200.calculators import rdf_postprocessing as post
from mdsuite_b200.database.project_database import Computation

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_consumer_golden  # noqa: E402  (only its seeded synthetic_rdf is used here)


def _cases():
    with open(os.path.join(HERE, "golden", "consumer_golden.json")) as fh:
        return json.load(fh)["cases"]


def _experiment(case, tmp_path):
    project = mds.Project("consumers", storage_path=str(tmp_path))
    exp = project.add_experiment("e", temperature=case["temperature"], units="real")
    n = sum(case["species"].values())
    exp.add_positions(np.zeros((1, n, 3), dtype=np.float32), case["species"], case["box"])
    return exp


def _rdf_computation(case):
    rdf = make_consumer_golden.synthetic_rdf(case)
    attrs = [("number_of_bins", {"serialized_value": case["nbins"]}),
             ("cutoff", {"serialized_value": case["cutoff_nm"] * 10.0}),
             ("number_of_configurations", {"serialized_value": 100})]
    results = []
    for key, data in rdf.items():
        a, b = key.split("_")
        results.append((data, [(a, 2)] if a == b else [(a, 1), (b, 1)]))
    return Computation(-1, "Radial_Distribution_Function", -1, attrs, results)


def _decode(v):
    return np.array([float.fromhex(x) for x in v]) if isinstance(v, list) else float.fromhex(v)


def _compare(comp, want_list):
    assert len(comp.keys()) == len(want_list)
    for want in want_list:
        got = comp["_".join(want["subjects"])]
        assert set(got) == set(want["data"])
        for k, v in want["data"].items():
            w, g = _decode(v), np.asarray(got[k], dtype=float)
            assert g.shape == np.shape(w), k
            fin = np.isfinite(w)
            assert np.array_equal(np.isfinite(g), fin), k
            assert np.allclose(g[fin] if g.ndim else g, w[fin] if g.ndim else w, rtol=1e-12, atol=0), k


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c["name"])
def test_coordination_numbers_equal_the_reference(case, tmp_path):
    exp = _experiment(case, tmp_path)
    comp = exp.run.CoordinationNumbers(rdf_data=_rdf_computation(case), plot=False,
                                       savgol_order=case["savgol"][0],
                                       savgol_window_length=case["savgol"][1],
                                       number_of_shells=case["shells"])
    _compare(comp, case["coordination_numbers"])


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c["name"])
def test_potential_of_mean_force_equals_the_reference(case, tmp_path):
    exp = _experiment(case, tmp_path)
    comp = exp.run.PotentialOfMeanForce(rdf_data=_rdf_computation(case), plot=False,
                                        savgol_order=case["savgol"][0],
                                        savgol_window_length=case["savgol"][1],
                                        number_of_shells=case["shells"])
    _compare(comp, case["potential_of_mean_force"])


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c["name"])
def test_kirkwood_buff_integral_equals_the_reference(case, tmp_path):
    exp = _experiment(case, tmp_path)
    comp = exp.run.KirkwoodBuffIntegral(rdf_data=_rdf_computation(case), plot=True,
                                        savgol_order=case["savgol"][0],
                                        savgol_window_length=case["savgol"][1])
    _compare(comp, case["kirkwood_buff"])
    assert os.path.exists(os.path.join(exp.figures_path, "Kirkwood-Buff_Integral.html"))


def test_ideal_gas_limits():
    """g(r) = 1: n(r) = 4/3 pi rho (r^3 - r_1^3), G(r) = 0, w(r) = 0."""
    x = np.linspace(0.0, 2.0, 401)
    radii, rdf = x[1:], np.ones(400)
    rho = 33.0
    cn = post.running_coordination(radii, rdf, rho)
    r = radii[1:]
    exact = 4 / 3 * np.pi * rho * (r[1:] ** 3 - r[0] ** 3)
    assert np.allclose(cn[100:], exact[100:], rtol=1e-4)  # trapezoid error matters only near r = 0
    assert np.allclose(post.kirkwood_buff_running_integral(radii, rdf, 2, 17), 0.0, atol=1e-12)
    assert np.all(post.potential_of_mean_force(rdf, 300.0) == 0.0)


def test_too_few_peaks_raises(tmp_path):
    case = dict(_cases()[0])
    exp = _experiment(case, tmp_path)
    from mdsuite_b200.calculators.coordination_number_calculation import CannotPerformThisAnalysis
    with pytest.raises(CannotPerformThisAnalysis):
        exp.run.CoordinationNumbers(rdf_data=_rdf_computation(case), plot=False, number_of_shells=40)
