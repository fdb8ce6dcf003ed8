"""CPU tests of the oracle (test infrastructure): self-consistency, golden vectors, quirks."""
import json
import os
import sys

import numpy as np
import pytest

from oracle import c_oracle, rdf_oracle, rdf_torch_cpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402


def golden_cases():
    with open(os.path.join(HERE, "golden", "rdf_golden.json")) as fh:
        return json.load(fh)["cases"]


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_oracles_reproduce_golden_vectors(case):
    pos, box = make_golden.case_inputs(case)
    import hashlib
    assert hashlib.sha256(pos.tobytes()).hexdigest() == case["positions_sha"]
    want = np.array(case["counts_parity"])
    blocked = rdf_oracle.rdf_counts_blocked(pos.transpose(1, 0, 2), case["counts"], box,
                                            case["cutoff"], case["nbins"])
    assert np.array_equal(blocked, want)
    c, _ = c_oracle.rdf_counts_c(pos, case["counts"], box, case["cutoff"], case["nbins"],
                                 layout=c_oracle.FRAME_MAJOR)
    assert np.array_equal(c, want)
    t, _ = rdf_torch_cpu.rdf_counts(pos.transpose(1, 0, 2), case["counts"], box, case["cutoff"],
                                    case["nbins"], minibatch=11)
    assert np.array_equal(t, want)
    cc, _ = c_oracle.rdf_counts_c(pos, case["counts"], box, case["cutoff"], case["nbins"],
                                  skip_first=False, layout=c_oracle.FRAME_MAJOR)
    assert np.array_equal(cc, np.array(case["counts_corrected"]))


def test_first_atom_quirk_pair_counts():
    """Q1 (reference :637-638): species sizes [5, 4] count 6 / 12 / 3 pairs, not 10 / 20 / 6."""
    rng = np.random.default_rng(0)
    pos = (rng.random((9, 3, 3)) * 10).astype(np.float32)
    box = np.array([10.0, 10.0, 10.0])
    c = rdf_oracle.rdf_counts(pos, [5, 4], box, 20.0, 50)  # cutoff > any distance
    assert (c.sum(1) // 3).tolist() == [6, 12, 3]
    full = rdf_oracle.rdf_counts_blocked(pos, [5, 4], box, 20.0, 50, skip_first=False)
    assert (full.sum(1) // 3).tolist() == [10, 20, 6]


@pytest.mark.parametrize("minibatch", [1, 2, 5, 9, 100])
def test_minibatch_size_does_not_change_result(minibatch):
    rng = np.random.default_rng(1)
    pos = (rng.random((23, 2, 3)) * 6).astype(np.float32)
    box = np.array([6.0, 6.5, 7.0])
    ref = rdf_oracle.rdf_counts(pos, [10, 13], box, 2.9, 40)
    assert np.array_equal(rdf_oracle.rdf_counts(pos, [10, 13], box, 2.9, 40, minibatch=minibatch), ref)


def test_partial_triu_indices_match_definition():
    idx = rdf_oracle.get_partial_triu_indices(7, 3, 2)  # rows 2..4 of a 7-atom system
    want = [(r, c) for r in range(3) for c in range(7) if c > r + 2]
    assert idx.dtype == np.int32 and idx.shape == (2, len(want))
    assert [tuple(x) for x in idx.T.tolist()] == want


def test_fast_minimum_image_is_bit_identical_to_rint_form():
    """The 2-op minimum image the CUDA path uses on wrapped frames."""
    rng = np.random.default_rng(2)
    for L in (np.float32(31.5), np.float32(75.6), np.float32(49.11), np.float32(1.0), np.float32(215.3)):
        d = ((rng.random(2_000_000) * 2.5 - 1.25) * L).astype(np.float32)
        # adversarial: +-64 ulps around +-L/2 and +-L
        for c in (L / 2, -L / 2, L, -L):
            x = np.full(129, c, dtype=np.float32)
            for k in range(1, 65):
                x[k] = np.nextafter(x[k - 1], np.float32(np.inf))
                x[64 + k] = np.nextafter(x[64 + k - 1] if k > 1 else c, np.float32(-np.inf))
            d = np.concatenate([d, x])
        ref = rdf_oracle.apply_minimum_image(d, np.array(L, dtype=np.float32))
        a = np.abs(d)
        fast = np.minimum(a, (L - a).astype(np.float32))
        assert np.array_equal(np.abs(ref), np.abs(fast))
        assert np.array_equal(ref * ref, fast * fast)


def test_threshold_table_reproduces_sqrt_and_histogram():
    """bin(s) from the s-space thresholds == the reference chain sqrt -> cutoff -> histogram."""
    rng = np.random.default_rng(3)
    for cutoff, nbins in ((10.0, 500), (37.699999999999996, 3769), (3.0, 300), (49.9, 5000)):
        edges, s_cut = rdf_oracle.s_threshold_table(cutoff, nbins)
        s = (rng.random(300_000) * np.float32(cutoff) ** 2 * 1.05).astype(np.float32)
        # plus values straddling every edge
        near = np.concatenate([np.nextafter(edges[1:], np.float32(-np.inf)), edges[1:],
                               np.nextafter(edges[1:], np.float32(np.inf))]).astype(np.float32)
        s = np.concatenate([s, near[np.isfinite(near)]])
        d = np.sqrt(s)
        keep_ref = d < np.float32(cutoff)
        assert np.array_equal(keep_ref, s < s_cut)
        bins_ref = rdf_oracle.bin_of_distance(d[keep_ref], cutoff, nbins)
        bins_tab = np.searchsorted(edges[1:nbins], s[keep_ref], side="right")
        assert np.array_equal(bins_ref, bins_tab)


def test_kernel_vs_pydoc_histogram_formula_differ_only_at_edges():
    rng = np.random.default_rng(4)
    d = (rng.random(2_000_000) * 10).astype(np.float32)
    a = rdf_oracle.tf_histogram_fixed_width(d, [0, 10.0], 500)
    b = rdf_oracle.tf_histogram_fixed_width_pydoc(d, [0, 10.0], 500)
    assert a.sum() == b.sum() == d.size
    assert np.abs(a - b).sum() <= 2 * 200  # a handful of edge cases at most (~1e-5)


def test_normalisation_quirks():
    """Q2: x = linspace(0, cutoff, B) in nm, y[0] is nan/inf; prefactor uses full species counts."""
    counts = np.ones((3, 100), dtype=np.int64)
    counts[:, 0] = 0
    out = rdf_oracle.normalise(counts, ["Na", "Cl"], [500, 500], np.array([31.5] * 3), 10.0, 100)
    assert list(out) == ["Na_Na", "Na_Cl", "Cl_Cl"]
    x, y = np.array(out["Na_Cl"]["x"]), np.array(out["Na_Cl"]["y"])
    assert np.allclose(x, 0.1 * np.linspace(0, 10.0, 100)) and np.isnan(y[0])
    # g = counts * scale / (F * rho_b * 4 pi r^2 dr * n_a)
    r = np.linspace(0, 10.0, 100)[1:]
    want = 1.0 / (100 * (500 / 31.5 ** 3) * 4 * np.pi * r ** 2 * (10.0 / 100) * 500)
    assert np.allclose(y[1:], want, rtol=1e-12)
    assert np.allclose(np.array(out["Na_Na"]["y"])[1:], 2 * want, rtol=1e-12)


def test_ideal_correction_branches():
    box0 = 10.0
    i1 = rdf_oracle.ideal_correction(4.9, 50, box0)
    assert np.allclose(i1, 4 * np.pi * np.linspace(0, 4.9, 50) ** 2 * (4.9 / 50))
    i2 = rdf_oracle.ideal_correction(6.5, 50, box0)
    r = np.linspace(0, 6.5, 50)
    mid = (r > 5.0) & (r < np.sqrt(2) * 5.0)
    assert np.allclose(i2[mid], (2 * np.pi * r * (3 - 4 * r))[mid] * (6.5 / 50))
    i3 = rdf_oracle.ideal_correction(8.0, 50, box0)
    assert len(i3) == 50 and np.isfinite(i3[-1])


def test_check_input_defaults():
    d = rdf_oracle.check_input_defaults(np.array([75.6, 75.6, 75.6]), 1000,
                                        number_of_configurations=1000)
    assert d["stop"] == 999 and abs(d["cutoff"] - 37.7) < 1e-9 and d["number_of_bins"] == 3769
    assert d["minibatch"] == 1000 and d["sample_configurations"].tolist() == list(range(1000))
    d = rdf_oracle.check_input_defaults(np.array([31.5] * 3), 100, number_of_configurations=7,
                                        start=10, stop=40)
    assert d["sample_configurations"].tolist() == [10, 15, 20, 25, 30, 35, 40]
