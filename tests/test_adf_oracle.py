"""The ADF restatement (oracle/adf_oracle.py) against tests/golden/adf_golden.json — the output of
the reference's own AngularDistributionFunction code run over a NumPy stand-in for its TensorFlow
leaf ops (tests/golden/make_adf_golden.py).  Same contracts on both sides, so: x axis and max_peak
exact, histograms to float32 summation noise (the reference sums weights in a different triple
order; tolerance 2e-6 of the curve's maximum), and exact for the count-only case (norm_power 0)."""
import json
import os

import numpy as np
import pytest

from oracle import adf_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "adf_golden.json")) as _fh:
    GOLDEN = json.load(_fh)["cases"]


def positions_for(case):
    rng = np.random.Generator(np.random.Philox(case["seed"]))
    n = sum(case["counts"])
    box = np.asarray(case["box"], dtype=np.float64)
    pos = rng.random((case["frames"], n, 3)) * box
    if case.get("unwrapped"):
        pos += rng.integers(-2, 3, size=pos.shape) * box
    return pos.astype(np.float32)


@pytest.mark.parametrize("case", GOLDEN, ids=[c["name"] for c in GOLDEN])
def test_restatement_equals_the_reference_run(case):
    pos = positions_for(case)
    got = adf_oracle.adf(pos, case["counts"], case["names"], case["box"], case["cutoff"],
                         case["nbins"], case["norm_power"], case["batches"])
    assert list(got) == ["-".join(r["subjects"]) for r in case["results"]]
    for rec in case["results"]:
        mine = got["-".join(rec["subjects"])]
        angle = np.array([float.fromhex(v) for v in rec["angle"]])
        want = np.array([float.fromhex(v) for v in rec["adf"]], dtype=np.float32)
        assert np.array_equal(mine["angle"], angle)
        assert mine["max_peak"] == rec["max_peak"]
        assert mine["adf"].dtype == np.float32
        if case["norm_power"] == 0:
            assert np.array_equal(mine["adf"], want)
        else:
            assert np.max(np.abs(mine["adf"].astype(np.float64) - want)) <= 2e-6 * want.max()


def test_sampled_configurations_follow_linspace():
    for case in GOLDEN:
        stop = case["frames"] - 1
        want = np.linspace(case.get("start", 1), stop, case["number_of_configurations"], dtype=int)
        assert case["sample_configurations"] == want.tolist()
        assert sum(case["batches"], []) == want.tolist()


def test_raw_histograms_are_counts_without_weights():
    case = GOLDEN[1]
    pos = positions_for(case)
    raw = adf_oracle.raw_histograms(pos[0], case["counts"], case["box"], case["cutoff"],
                                    case["nbins"], 0)
    assert raw.shape == (10, case["nbins"])
    assert np.array_equal(raw, np.round(raw)) and raw.sum() > 0
    triples = adf_oracle.frame_triples(pos[0], case["counts"], case["box"], case["cutoff"], 0)
    assert raw.sum() == sum(len(t) for t, _ in triples.values())


def test_float16_cutoff_decides_the_neighbours():
    """3.2 is not a float16 (it becomes 3.19921875); norms of 3.1985 and 3.1995 round to that same
    float16 and are therefore NOT below the cutoff, although they are below 3.2."""
    n = np.array([3.1975, 3.1985, 3.1995, 3.2005, 0.0, 1e-9, 7e4], dtype=np.float32)
    assert adf_oracle.neighbour_mask(n, 3.2).tolist() == [True, False, False, False, False, False,
                                                          False]
