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


def test_signed_short_minimum_image_identity():
    """The kernel's 3-op signed minimum image for |d| <= 1.25 L,
        a = |d|; b = L - a; r = d if a <= b else (-b if d > 0 else b),
    equals the reference's d - rint(d / L) * L bit for bit in float32 — checked here in NumPy on
    random boxes, on d spread over [-1.25 L, 1.25 L] and on every float within 64 ulps of the
    places where rint switches (+-L/2) and where the result changes sign (+-L)."""
    rng = np.random.Generator(np.random.Philox(2024))
    boxes = np.concatenate([rng.uniform(0.5, 300.0, 60), [1.0, 2.0, 7.0, 10.0, 49.1, 75.6, 215.3]])
    for L in boxes.astype(np.float32):
        spread = rng.uniform(-1.25, 1.25, 20000).astype(np.float32) * L
        spread = spread[np.abs(spread) <= np.float32(1.25) * L]
        near = []
        for centre in (0.5 * float(L), -0.5 * float(L), float(L), -float(L), 0.0):
            x = np.float32(centre)
            ups, downs = [x], [x]
            for _ in range(64):
                ups.append(np.nextafter(ups[-1], np.float32(np.inf)))
                downs.append(np.nextafter(downs[-1], np.float32(-np.inf)))
            near += ups + downs
        d = np.concatenate([spread, np.array(near, dtype=np.float32)])
        general = d - np.rint(d / L) * L
        a = np.abs(d)
        b = L - a
        short = np.where(a <= b, d, np.where(d > 0, -b, b)).astype(np.float32)
        same = (general.view(np.uint32) == short.view(np.uint32)) | ((general == 0) & (short == 0))
        assert same.all(), (float(L), d[~same][:5], general[~same][:5], short[~same][:5])
