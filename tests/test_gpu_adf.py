"""GPU parity of the ADF path, through the C ABI (include/mds_adf.h): per-configuration weight sums
against the NumPy oracle.  With norm_power = 0 the sums are counts and must be IDENTICAL (same
neighbours under the float16 cutoff, same bin for every triple); with weights the kernel adds
float32 terms in another order than NumPy: tolerance 2e-6 of the row maximum, written below."""
import json
import os

import numpy as np
import pytest

from oracle import adf_oracle

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
WEIGHT_TOL = 2e-6


def _adf():
    from mdsuite_b200._cuda import adf as cuda_adf
    return cuda_adf


def uniform(seed, counts, frames, box, unwrapped=False):
    rng = np.random.Generator(np.random.Philox(seed))
    box = np.asarray(box, dtype=np.float64)
    pos = rng.random((frames, sum(counts), 3)) * box
    if unwrapped:
        pos += rng.integers(-3, 4, size=pos.shape) * box
    return pos.astype(np.float32)


def oracle_raw(pos, counts, box, cutoff, nbins, power):
    return np.stack([adf_oracle.raw_histograms(f, counts, box, cutoff, nbins, power) for f in pos])


def check(pos, counts, box, cutoff, nbins, power, **kw):
    with _adf().ADFHandle(counts, nbins, cutoff, box, norm_power=power, **kw) as h:
        got = h.compute(pos)
        st = h.stats()
    want = oracle_raw(pos, counts, box, cutoff, nbins, power)
    assert got.shape == want.shape
    if power == 0:
        assert np.array_equal(got, want), f"{int(np.sum(got != want))} bins differ"
        assert st.triples == want.sum()
    else:
        scale = np.maximum(want.max(axis=-1, keepdims=True), 1e-300)
        assert np.max(np.abs(got - want) / scale) <= WEIGHT_TOL
    return got, st


CASES = [
    dict(id="two_species", counts=[40, 30], box=[9.0, 9.0, 9.0], cutoff=3.5, nbins=50, frames=3),
    dict(id="three_species_noncubic", counts=[25, 31, 17], box=[7.0, 9.5, 8.25], cutoff=3.0,
         nbins=36, frames=2),
    dict(id="single_species_cutoff_not_float16", counts=[97], box=[8.0, 8.0, 8.0], cutoff=3.2,
         nbins=100, frames=4),
    dict(id="unwrapped", counts=[33, 20], box=[8.0, 8.0, 8.0], cutoff=3.2, nbins=64, frames=3,
         unwrapped=True),
    dict(id="ragged_species", counts=[1, 0, 50, 2], box=[7.0, 7.0, 7.0], cutoff=2.9, nbins=40,
         frames=2),
    dict(id="cutoff_beyond_half_box", counts=[20, 12], box=[5.0, 5.0, 5.0], cutoff=4.0, nbins=30,
         frames=2),
    dict(id="many_bins", counts=[60], box=[9.0, 9.0, 9.0], cutoff=3.4, nbins=3000, frames=2),
    dict(id="one_bin", counts=[30, 30], box=[9.0, 9.0, 9.0], cutoff=3.4, nbins=1, frames=1),
    dict(id="more_atoms_than_an_item", counts=[150, 143], box=[14.0, 14.0, 14.0], cutoff=3.3,
         nbins=90, frames=2),
]


@pytest.mark.parametrize("power", [0, 4])
@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_weight_sums_equal_the_oracle(case, power):
    pos = uniform(11 + len(case["id"]), case["counts"], case["frames"], case["box"],
                  case.get("unwrapped", False))
    check(pos, case["counts"], case["box"], case["cutoff"], case["nbins"], power)


@pytest.mark.parametrize("power", [1, 2, 7])
def test_other_powers(power):
    pos = uniform(5, [30, 20], 2, [8.0, 8.0, 8.0])
    check(pos, [30, 20], [8.0, 8.0, 8.0], 3.3, 45, power)


def test_reference_run_cases_through_the_calculator_arithmetic():
    """The positions of the reference-run fixtures: counts identical to the restatement that the
    fixtures pin."""
    with open(os.path.join(HERE, "golden", "adf_golden.json")) as fh:
        cases = json.load(fh)["cases"]
    for case in cases:
        rng = np.random.Generator(np.random.Philox(case["seed"]))
        box = np.asarray(case["box"], dtype=np.float64)
        pos = rng.random((case["frames"], sum(case["counts"]), 3)) * box
        if case.get("unwrapped"):
            pos += rng.integers(-2, 3, size=pos.shape) * box
        check(pos.astype(np.float32), case["counts"], case["box"], case["cutoff"], case["nbins"], 0)


def test_coincident_nan_and_far_atoms():
    """Two atoms on the same point (norm 0: float16 zero, not neighbours), a NaN atom and an atom at
    1e30 take part in no triple; everything else is untouched."""
    counts, box = [30, 24], [8.0, 8.0, 8.0]
    pos = uniform(21, counts, 2, box)
    pos[0, 5] = pos[0, 4]
    pos[0, 40] = np.nan
    pos[1, 7] = [1e30, 0.0, 0.0]
    pos[1, 8] = [np.inf, 1.0, 2.0]
    with np.errstate(all="ignore"):
        check(pos, counts, box, 3.2, 48, 0)


def test_frames_are_independent_and_batched():
    counts, box = [48], [8.0, 8.0, 8.0]
    pos = uniform(31, counts, 9, box)
    got, _ = check(pos, counts, box, 3.1, 32, 0, max_frames_in_flight=4)
    with _adf().ADFHandle(counts, 32, 3.1, box, norm_power=0) as h:
        single = np.stack([h.compute(pos[k:k + 1])[0] for k in range(len(pos))])
        assert h.stats().frames == len(pos)
    assert np.array_equal(single, got)


def test_device_entry_point_equals_host_entry_point():
    import torch
    counts, box = [35, 29], [8.0, 8.0, 8.0]
    pos = uniform(41, counts, 3, box)
    with _adf().ADFHandle(counts, 40, 3.3, box, norm_power=4) as h:
        host = h.compute(pos)
        dev = h.compute_device(torch.from_numpy(pos).cuda())
        torch.cuda.synchronize()
    assert np.array_equal(dev.cpu().numpy(), host) or \
        np.max(np.abs(dev.cpu().numpy() - host)) <= WEIGHT_TOL * host.max()


def test_too_many_neighbours_is_an_error_not_a_wrong_answer():
    cuda_adf, cuda = _adf(), __import__("mdsuite_b200._cuda", fromlist=["x"])
    counts, box = [1300], [6.0, 6.0, 6.0]
    pos = uniform(51, counts, 1, box)
    with cuda_adf.ADFHandle(counts, 20, 5.9, box, norm_power=0) as h:
        with pytest.raises(cuda.MdsCudaError, match="neighbours"):
            h.compute(pos)


def test_neighbour_capacity_grows_on_demand(monkeypatch):
    """Started with room for 32 neighbours per atom, the library doubles the capacity and repeats
    the batch until the lists fit; the result is the same as with ample room."""
    counts, box = [90, 70], [8.0, 8.0, 8.0]
    pos = uniform(53, counts, 3, box)
    monkeypatch.setenv("MDS_ADF_NEIGHBOUR_CAPACITY", "32")
    got, st = check(pos, counts, box, 3.9, 48, 0)
    assert st.max_neighbours > 64 and st.frames == 3
    monkeypatch.setenv("MDS_ADF_NEIGHBOUR_CAPACITY", "512")
    with _adf().ADFHandle(counts, 48, 3.9, box, norm_power=0) as h:
        assert np.array_equal(h.compute(pos), got)


def test_dense_liquid_at_scale_properties():
    """4 096 atoms: totals against the triple count from an independent neighbour count, symmetry
    of the same-species rows (every unordered pair is two triples: even counts)."""
    counts, box = [2048, 2048], [36.0, 36.0, 36.0]
    pos = uniform(61, counts, 2, box)
    with _adf().ADFHandle(counts, 180, 4.0, box, norm_power=0) as h:
        got = h.compute(pos)
        st = h.stats()
    s_lo, s_hi = _adf().neighbour_window(4.0)
    species = np.repeat([0, 1], counts)
    for f in range(2):
        r = adf_oracle.min_image_vectors(pos[f], box)
        sq = r * r
        s = (sq[..., 0] + sq[..., 1]) + sq[..., 2]
        nb = (s >= s_lo) & (s < s_hi)
        n0 = (nb & (species[None, :] == 0)).sum(1)
        n1 = (nb & (species[None, :] == 1)).sum(1)
        c0, c1 = species == 0, species == 1
        want = [np.sum(n0[c0] * (n0[c0] - 1)), np.sum(n0[c0] * n1[c0]),
                np.sum(n1[c0] * (n1[c0] - 1)), np.sum(n1[c1] * (n1[c1] - 1))]
        assert got[f].sum(axis=1).tolist() == [float(w) for w in want]
    assert np.all(got[:, [0, 2, 3]] % 2 == 0)
    assert st.triples == got.sum() and st.max_neighbours < 64


# ---- cell-list neighbour search ---------------------------------------------------------------

def _compute_with_path(monkeypatch, path, pos, counts, box, cutoff, nbins, power):
    if path is None:
        monkeypatch.delenv("MDS_ADF_PATH", raising=False)
    else:
        monkeypatch.setenv("MDS_ADF_PATH", path)
    with _adf().ADFHandle(counts, nbins, cutoff, box, norm_power=power) as h:
        out = h.compute(pos)
        return out, h.stats()


CELL_CASES = [
    dict(id="cubic_two_species", counts=[300, 280], box=[16.0, 16.0, 16.0], cutoff=3.3, nbins=90),
    dict(id="noncubic_three_species", counts=[200, 150, 90], box=[11.0, 17.5, 13.0], cutoff=3.1,
         nbins=64),
    dict(id="exactly_three_cells", counts=[400], box=[10.0, 10.0, 10.0], cutoff=3.3, nbins=50),
    dict(id="unwrapped_within_8_boxes", counts=[260, 100], box=[14.0, 14.0, 14.0], cutoff=3.0,
         nbins=72, unwrapped=True),
    # coordinates in [-L/2, L/2): the other window in which the short minimum image is valid
    dict(id="centred_box", counts=[280, 150], box=[13.0, 15.0, 14.0], cutoff=3.2, nbins=80,
         shift=-0.5),
    # one atom a little outside [0, L): neither window holds, the general minimum image is used
    dict(id="one_atom_outside_both_windows", counts=[300], box=[12.0, 12.0, 12.0], cutoff=3.0,
         nbins=60, stray=True),
]


@pytest.mark.parametrize("case", CELL_CASES, ids=[c["id"] for c in CELL_CASES])
def test_cell_search_equals_all_atoms_search_and_oracle(monkeypatch, case):
    pos = uniform(71 + len(case["id"]), case["counts"], 2, case["box"], case.get("unwrapped", False))
    if case.get("shift"):
        pos = (pos + np.float32(case["shift"]) * np.asarray(case["box"], dtype=np.float32)).astype(np.float32)
    if case.get("stray"):
        pos[:, 7, 1] += np.float32(1.3 * case["box"][1])
    # atoms exactly on cell and box boundaries
    pos[0, 0] = [0.0, 0.0, 0.0]
    pos[0, 1] = [case["box"][0], case["box"][1], case["box"][2]]
    pos[0, 2] = [case["box"][0] / 3, case["box"][1] / 3, case["box"][2] / 3]
    pos[0, 3] = np.nextafter(np.float32(case["box"][0]), np.float32(0))
    args = (pos, case["counts"], case["box"], case["cutoff"], case["nbins"])
    cells, st_c = _compute_with_path(monkeypatch, "cells", *args, 0)
    allatoms, st_a = _compute_with_path(monkeypatch, "allatoms", *args, 0)
    auto, st_auto = _compute_with_path(monkeypatch, None, *args, 0)
    assert min(st_c.cells) >= 3 and st_a.cells == (0, 0, 0) and st_auto.cells == st_c.cells
    # with exactly three cells per dimension the 27-cell stencil is the whole box
    assert st_c.pair_tests <= st_a.pair_tests == 2 * sum(case["counts"]) ** 2
    want = oracle_raw(*args, 0)
    assert np.array_equal(allatoms, want)
    assert np.array_equal(cells, want)
    assert np.array_equal(auto, want)
    weighted, _ = _compute_with_path(monkeypatch, "cells", *args, 4)
    with np.errstate(all="ignore"):
        want_w = oracle_raw(*args, 4)
        # the atom one float below the box edge is ~1e-6 from the atom at the origin: its weights
        # overflow to inf on both sides
        assert np.array_equal(np.isinf(weighted), np.isinf(want_w))
        finite = np.where(np.isinf(want_w), 0.0, want_w)
        scale = np.maximum(finite.max(axis=-1, keepdims=True), 1e-300)
        diff = np.where(np.isinf(want_w), 0.0, np.abs(weighted - want_w))
    assert np.max(diff / scale) <= WEIGHT_TOL


@pytest.mark.parametrize("capacity", [None, "32"])
def test_cell_search_hands_far_configurations_to_the_all_atoms_search(monkeypatch, capacity):
    if capacity:  # too small for 37 neighbours: the batch is repeated with twice the room
        monkeypatch.setenv("MDS_ADF_NEIGHBOUR_CAPACITY", capacity)
    counts, box = [300, 212], [15.0, 15.0, 15.0]
    pos = uniform(81, counts, 4, box)
    pos[1, 17, 0] += 9 * 15.0       # beyond 8 box lengths: same atom under the minimum image
    pos[3, 400] = [np.inf, 1.0, 2.0]
    with np.errstate(all="ignore"):
        want = oracle_raw(pos, counts, box, 3.2, 60, 0)
    got, st = _compute_with_path(monkeypatch, "cells", pos, counts, box, 3.2, 60, 0)
    assert st.frames_far == 2
    assert np.array_equal(got, want)


def test_cell_path_is_refused_for_small_boxes(monkeypatch):
    cuda = __import__("mdsuite_b200._cuda", fromlist=["x"])
    monkeypatch.setenv("MDS_ADF_PATH", "cells")
    with pytest.raises(cuda.MdsCudaError, match="3.006"):
        _adf().ADFHandle([300], 10, 3.4, [10.0, 10.0, 10.0])


def test_clustered_atoms_in_one_cell(monkeypatch):
    """300 atoms in one cell corner and a dilute background: long neighbour lists next to empty
    cells."""
    rng = np.random.Generator(np.random.Philox(91))
    box = [20.0, 20.0, 20.0]
    cluster = rng.random((1, 120, 3)) * 3.0 + 1.0
    rest = rng.random((1, 300, 3)) * 20.0
    pos = np.concatenate([cluster, rest], axis=1).astype(np.float32)
    counts = [200, 220]
    want = oracle_raw(pos, counts, box, 3.0, 40, 0)
    got, st = _compute_with_path(monkeypatch, "cells", pos, counts, box, 3.0, 40, 0)
    assert np.array_equal(got, want) and st.max_neighbours > 100


# ---- the calculator call, end to end on the GPU -------------------------------------------------

def test_calculator_call_equals_the_reference_run(tmp_path):
    """experiment.run.AngularDistributionFunction(...) with the real engine against the output of
    the reference's own run_calculator (tests/golden/adf_golden.json)."""
    from tests.test_adf_calculator_cpu import GOLDEN, make_experiment
    for case in GOLDEN:
        project, exp = make_experiment(tmp_path / case["name"], case)
        comp = exp.run.AngularDistributionFunction(
            number_of_configurations=case["number_of_configurations"], cutoff=case["cutoff"],
            start=case.get("start", 1), number_of_bins=case["nbins"],
            norm_power=case["norm_power"], batches=case["n_batches"], plot=False)
        assert comp.keys() == ["_".join(r["subjects"]) for r in case["results"]]
        for rec in case["results"]:
            got = comp["_".join(rec["subjects"])]
            want = np.array([float.fromhex(v) for v in rec["adf"]], dtype=np.float32)
            assert got["max_peak"] == rec["max_peak"]
            assert np.array_equal(np.array(got["angle"]),
                                  np.array([float.fromhex(v) for v in rec["angle"]]))
            mine = np.array(got["adf"])
            if case["norm_power"] == 0:
                assert np.array_equal(mine.astype(np.float32), want)
            else:
                assert np.max(np.abs(mine - want)) <= WEIGHT_TOL * want.max()


def test_calculator_on_a_melt_from_a_dump(tmp_path):
    """LAMMPS dump -> project -> ADF on the cell path; curves against the oracle's."""
    import mdsuite_b200 as mds
    from mdsuite_b200 import synthetic
    sysm = synthetic.nacl_melt(4, 4, seed=12)  # 512 atoms
    dump = tmp_path / "melt.lammpstraj"
    synthetic.write_lammps_dump(sysm, str(dump))
    project = mds.Project("adf_melt", storage_path=str(tmp_path))
    exp = project.add_experiment("NaCl", timestep=0.002, temperature=1400.0, units="metal",
                                 simulation_data=str(dump))
    comp = exp.run.AngularDistributionFunction(number_of_configurations=3, cutoff=4.5, start=0,
                                               number_of_bins=90, batches=3, plot=True)
    want = adf_oracle.adf(sysm.positions, sysm.counts, ["Na", "Cl"], sysm.box, 4.5, 90, 4,
                          [[0], [1], [3]])
    assert comp.keys() == ["Na_Na_Na", "Na_Na_Cl", "Na_Cl_Cl", "Cl_Cl_Cl"]
    for key in comp.keys():
        ref = want[key.replace("_", "-")]
        assert comp[key]["max_peak"] == ref["max_peak"]
        assert np.max(np.abs(np.array(comp[key]["adf"]) - ref["adf"])) <= WEIGHT_TOL * ref["adf"].max()


# ---- random shapes ------------------------------------------------------------------------------

# MDS_ADF_FUZZ_EXTRA=n adds n more seeds (run once with 400 on a B200 after the last kernel change)
@pytest.mark.parametrize("seed", range(24 + int(os.environ.get("MDS_ADF_FUZZ_EXTRA", "0"))))
def test_random_shapes_counts_are_identical(seed, monkeypatch):
    """Random species tables, boxes, cutoffs, bin counts and coordinate ranges; whichever search the
    library picks, and the other one when both apply."""
    rng = np.random.Generator(np.random.Philox(1000 + seed))
    n_species = int(rng.integers(1, 5))
    counts = [int(c) for c in rng.integers(0, 140, n_species)]
    if sum(counts) < 3:
        counts[0] += 40
    box = [float(b) for b in rng.uniform(7.0, 13.0, 3)]
    cutoff = float(rng.uniform(1.5, 3.4))
    nbins = int(rng.choice([1, 9, 50, 181, 500, 1024]))
    pos = uniform(2000 + seed, counts, int(rng.integers(1, 4)), box, unwrapped=bool(seed % 3 == 0))
    if seed % 4 == 1:
        pos -= np.float32(0.5) * np.asarray(box, dtype=np.float32)
    want = oracle_raw(pos, counts, box, cutoff, nbins, 0)
    for path in (None, "allatoms") + (("cells",) if min(box) >= 3.02 * cutoff else ()):
        got, st = _compute_with_path(monkeypatch, path, pos, counts, box, cutoff, nbins, 0)
        assert np.array_equal(got, want), (seed, path, counts, box, cutoff, nbins)
        assert st.triples == want.sum()


# ---- histograms that do not fit in shared memory ------------------------------------------------

def test_many_species_use_the_global_histogram():
    """6 species = 56 combinations x 500 bins x 8 bytes do not fit beside the neighbour lists: the
    sums go straight to the configuration's global histogram."""
    counts = [30, 25, 20, 28, 22, 26]
    box = [9.0, 9.0, 9.0]
    pos = uniform(301, counts, 2, box)
    got, st = check(pos, counts, box, 3.0, 500, 0)
    assert got.shape[1] == 56
    check(pos, counts, box, 3.0, 500, 4)


def test_global_histogram_equals_shared_histogram(monkeypatch):
    counts, box = [300, 260], [15.0, 15.0, 15.0]
    pos = uniform(302, counts, 3, box)
    shared, _ = check(pos, counts, box, 3.2, 120, 0)
    monkeypatch.setenv("MDS_ADF_HIST", "global")
    glob, _ = check(pos, counts, box, 3.2, 120, 0)
    assert np.array_equal(shared, glob)


def test_very_fine_bins_walk_the_table():
    """20 000 bins are narrower than the arccosine guess is accurate: the bin is found by walking
    the threshold table, not by one step."""
    counts, box = [120], [9.0, 9.0, 9.0]
    pos = uniform(303, counts, 2, box)
    check(pos, counts, box, 3.3, 20000, 0)


def test_cfg2_size_totals_against_blockwise_neighbour_counts():
    """13 824-atom melt (the cfg2 shape): per-combination totals of the triple counts against
    n_B (n_B - 1) / n_B n_C sums from neighbour counts computed blockwise in NumPy with the same
    float32 arithmetic and the same squared-distance window."""
    from mdsuite_b200 import synthetic
    sysm = synthetic.nacl_melt(12, 2, seed=2)
    cutoff = 6.0
    with _adf().ADFHandle(sysm.counts, 500, cutoff, sysm.box, norm_power=0) as h:
        got = h.compute(sysm.positions)
        st = h.stats()
    assert st.cells == (12, 12, 12) and st.frames_far == 0
    s_lo, s_hi = _adf().neighbour_window(cutoff)
    box = np.asarray(sysm.box, dtype=np.float32)
    species = np.repeat([0, 1], sysm.counts)
    for f in range(2):
        pos = sysm.positions[f]
        n0 = np.zeros(len(pos), dtype=np.int64)
        n1 = np.zeros(len(pos), dtype=np.int64)
        for lo in range(0, len(pos), 1024):
            r = pos[lo:lo + 1024, None, :] - pos[None, :, :]
            r = r - np.rint(r / box) * box
            sq = r * r
            s = (sq[..., 0] + sq[..., 1]) + sq[..., 2]
            nb = (s >= s_lo) & (s < s_hi)
            n0[lo:lo + 1024] = (nb & (species[None, :] == 0)).sum(1)
            n1[lo:lo + 1024] = (nb & (species[None, :] == 1)).sum(1)
        c0, c1 = species == 0, species == 1
        want = [np.sum(n0[c0] * (n0[c0] - 1)), np.sum(n0[c0] * n1[c0]),
                np.sum(n1[c0] * (n1[c0] - 1)), np.sum(n1[c1] * (n1[c1] - 1))]
        assert got[f].sum(axis=1).tolist() == [float(w) for w in want]
    assert st.triples == got.sum()


def test_handle_reuse_with_growing_and_empty_batches():
    """One handle, calls of 1, 6, 0 and 3 configurations: the per-configuration device buffers grow
    on demand, an empty call returns an empty result, nothing leaks between calls."""
    counts, box = [260, 200], [14.0, 14.0, 14.0]
    pos = uniform(401, counts, 6, box)
    want = oracle_raw(pos, counts, box, 3.1, 64, 0)
    with _adf().ADFHandle(counts, 64, 3.1, box, norm_power=0) as h:
        assert np.array_equal(h.compute(pos[:1]), want[:1])
        assert np.array_equal(h.compute(pos), want)
        empty = h.compute(pos[:0])
        assert empty.shape == (0, 4, 64)
        assert np.array_equal(h.compute(pos[2:5]), want[2:5])
        assert h.stats().frames == 10
