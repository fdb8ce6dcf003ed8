"""Golden vectors produced by the REFERENCE'S OWN CODE (tests/golden/make_reference_golden.py runs
mdsuite/calculators/radial_distribution_function.py::run_calculator here, over a NumPy stand-in for
its TensorFlow leaf ops).  They pin the oracle's — and the product's — control flow, index masks
(quirk Q1), configuration sampling, defaults and normalisation to what the reference's Python does;
the float32 leaf arithmetic remains a restatement (see the generator's docstring).

CPU tests: the oracle (NumPy literal, NumPy blocked, C) and the product's host logic against the
fixtures.  GPU test: the CUDA path against the same histograms.
"""
import json
import os
import sys

import numpy as np
import pytest

import mdsuite_b200 as mds
from mdsuite_b200.calculators.radial_distribution_function import RadialDistributionFunction
from oracle import c_oracle, rdf_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402


def _cases():
    with open(os.path.join(HERE, "golden", "reference_golden.json")) as fh:
        return json.load(fh)["cases"]


def _inputs(case):
    r = case["resolved"]
    pos, box = make_golden.case_inputs({**case, "cutoff": r["cutoff"], "nbins": r["nbins"]})
    return pos, box, r


def _want_hist(case):
    return np.array([case["histograms"][k] for k in case["resolved"]["key_list"]], dtype=np.int64)


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c["name"])
def test_oracle_histograms_equal_the_reference_run(case):
    pos, box, r = _inputs(case)
    frames = pos[r["sample_configurations"]]
    want = _want_hist(case)
    mb = case["minibatch"] if case["minibatch"] > 0 else r["number_of_configurations"]
    literal = rdf_oracle.rdf_counts(frames.transpose(1, 0, 2), case["counts"], box, r["cutoff"],
                                    r["nbins"], minibatch=mb)
    assert np.array_equal(literal, want)
    blocked = rdf_oracle.rdf_counts_blocked(frames.transpose(1, 0, 2), case["counts"], box,
                                            r["cutoff"], r["nbins"], skip_first=True)
    assert np.array_equal(blocked, want)
    c, _ = c_oracle.rdf_counts_c(frames, case["counts"], box, r["cutoff"], r["nbins"],
                                 skip_first=True, layout=c_oracle.FRAME_MAJOR)
    assert np.array_equal(c, want)


def test_q1_first_atom_exclusion_is_what_the_reference_code_does():
    """Species sizes [5, 4], every pair inside the cutoff: the reference's own masks count
    C(4,2) / 4*3 / C(3,2) pairs per configuration, not C(5,2) / 5*4 / C(4,2)."""
    case = next(c for c in _cases() if c["name"] == "q1_demo_5_4")
    per_frame = (_want_hist(case).sum(1) / case["frames"]).tolist()
    assert per_frame == [6.0, 12.0, 3.0]


class _OracleHandle:
    """Test double for the GPU handle (host-logic test): answers from the C oracle."""

    def __init__(self, counts, nbins, cutoff, box, parity=True, device=0, path=None, **kw):
        self.cfg = (list(counts), nbins, cutoff, np.asarray(box), parity)
        self.total = None
        self.ev = 0

    def submit(self, positions, layout=0, pinned=False):
        pos = np.array(positions.numpy() if hasattr(positions, "numpy") else positions)
        counts, nbins, cutoff, box, parity = self.cfg
        c, ev = c_oracle.rdf_counts_c(pos, counts, box, cutoff, nbins, skip_first=parity,
                                      layout=c_oracle.FRAME_MAJOR)
        self.total = c.astype(np.uint64) if self.total is None else self.total + c.astype(np.uint64)
        self.ev += ev

    def wait_host_copies(self):
        pass

    def counts(self):
        return self.total.copy()

    def stats(self):
        from types import SimpleNamespace
        return SimpleNamespace(pairs_evaluated=float(self.ev), frames_general_minimage=0)

    def close(self):
        pass


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c["name"])
def test_product_host_logic_equals_the_reference_run(case, tmp_path, monkeypatch):
    """Defaults (cutoff, bins, stop), np.linspace sampling, key order, prefactor, ideal correction
    (all three branches), nm x axis: the calculator's queued x / y against the reference's."""
    monkeypatch.setattr(RadialDistributionFunction, "handle_factory", _OracleHandle)
    pos, box, r = _inputs(case)
    project = mds.Project("ref", storage_path=str(tmp_path))
    exp = project.add_experiment("e", units="real")
    exp.add_positions(pos, dict(zip(case["species"], case["counts"])), box)
    comp = exp.run.RadialDistributionFunction(
        number_of_bins=case["nbins"], cutoff=case["cutoff"], plot=False,
        number_of_configurations=case.get("number_of_configurations", case["frames"]),
        minibatch=case["minibatch"])
    assert comp.keys() == r["key_list"]
    for res in case["results"]:
        key = "_".join(res["subjects"])
        want_x = np.array([float.fromhex(v) for v in res["x"]])
        want_y = np.array([float.fromhex(v) for v in res["y"]])
        got_x, got_y = np.array(comp[key]["x"]), np.array(comp[key]["y"])
        assert np.array_equal(got_x, want_x)
        assert np.array_equal(np.isnan(got_y), np.isnan(want_y))
        assert np.array_equal(np.isinf(got_y), np.isinf(want_y))
        fin = np.isfinite(want_y)
        assert np.allclose(got_y[fin], want_y[fin], rtol=1e-12, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("case", _cases(), ids=lambda c: c["name"])
def test_cuda_histograms_equal_the_reference_run(case):
    from mdsuite_b200 import _cuda
    pos, box, r = _inputs(case)
    frames = np.ascontiguousarray(pos[r["sample_configurations"]])
    for path in ("allpairs", None):
        with _cuda.RDFHandle(case["counts"], r["nbins"], r["cutoff"], box, parity=True, path=path) as h:
            h.submit(frames)
            got = h.counts()
        assert np.array_equal(got.astype(np.int64), _want_hist(case))
