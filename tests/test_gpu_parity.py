"""GPU parity: the CUDA path (through the C ABI) against the oracle, bit-exact integer histograms.

The oracle restates the reference's op sequence (oracle/rdf_oracle.py, oracle/rdf_oracle.c);
tolerance for the integer histograms is ZERO.
"""
import os

import numpy as np
import pytest

from mdsuite_b200 import _cuda, synthetic
from oracle import c_oracle, rdf_oracle

pytestmark = pytest.mark.gpu


def gpu_counts(pos_fm, counts, box, cutoff, nbins, parity=True, layout=_cuda.LAYOUT_FRAME_MAJOR,
               device_input=False, max_frames=0, n_submits=1):
    import torch
    with _cuda.RDFHandle(counts, nbins, cutoff, box, parity=parity,
                         max_frames_in_flight=max_frames) as h:
        arr = pos_fm if layout == _cuda.LAYOUT_FRAME_MAJOR else np.ascontiguousarray(
            pos_fm.transpose(1, 0, 2))
        chunks = np.array_split(np.arange(pos_fm.shape[0]), n_submits)
        for ch in chunks:
            if len(ch) == 0:
                continue
            if layout == _cuda.LAYOUT_FRAME_MAJOR:
                part = np.ascontiguousarray(arr[ch[0]:ch[-1] + 1])
            else:
                part = np.ascontiguousarray(arr[:, ch[0]:ch[-1] + 1])
            if device_input:
                h.submit_device(torch.from_numpy(part).cuda(), layout)
            else:
                h.submit(part, layout)
        out = h.counts()
        st = h.stats()
    return out, st


def oracle_counts(pos_fm, counts, box, cutoff, nbins, parity=True):
    out, ev = c_oracle.rdf_counts_c(pos_fm, counts, box, cutoff, nbins, skip_first=parity,
                                    layout=c_oracle.FRAME_MAJOR)
    return out.astype(np.uint64), ev


def assert_same(got, want, st=None, ev=None):
    diff = np.argwhere(got != want)
    assert diff.size == 0, f"{len(diff)} bins differ, first: {diff[:5].tolist()} " \
                           f"got {got[tuple(diff[0])]} want {want[tuple(diff[0])]}"
    if st is not None and ev is not None:
        assert st.pairs_allpairs_equiv == ev
        if st.path == _cuda.PATH_ALLPAIRS:
            assert st.pairs_evaluated == ev
        assert st.pairs_binned == float(want.sum())


def test_tiny_two_species_vs_literal_numpy_oracle():
    rng = np.random.default_rng(0)
    counts, box = [5, 4], np.array([10.0, 10.0, 10.0])
    pos = (rng.random((3, 9, 3)) * 10).astype(np.float32)
    for parity in (True, False):
        got, _ = gpu_counts(pos, counts, box, 4.9, 50, parity=parity)
        want = rdf_oracle.rdf_counts_blocked(pos.transpose(1, 0, 2), counts, box, 4.9, 50,
                                             skip_first=parity)
        assert_same(got, want.astype(np.uint64))
    lit = rdf_oracle.rdf_counts(pos.transpose(1, 0, 2), counts, box, 4.9, 50, minibatch=3)
    got, _ = gpu_counts(pos, counts, box, 4.9, 50, parity=True)
    assert_same(got, lit.astype(np.uint64))
    assert got.sum(1).tolist() == [sum(r) for r in lit.tolist()]


@pytest.mark.parametrize("parity", [True, False])
def test_cfg1_shape_nacl_1000(parity):
    sysm = synthetic.nacl_melt(5, 4, seed=1)
    got, st = gpu_counts(sysm.positions, sysm.counts, sysm.box, 10.0, 500, parity=parity)
    want, ev = oracle_counts(sysm.positions, sysm.counts, sysm.box, 10.0, 500, parity=parity)
    assert_same(got, want, st, ev)
    assert st.frames_general_minimage == 0


def test_three_species_noncubic_cutoff_beyond_half_box():
    rng = np.random.default_rng(3)
    counts, box = [300, 1, 211], np.array([17.3, 21.0, 19.5])
    pos = (rng.random((5, 512, 3)) * box).astype(np.float32)
    for cutoff, nbins in ((8.0, 333), (14.0, 1000)):
        got, st = gpu_counts(pos, counts, box, cutoff, nbins)
        want, ev = oracle_counts(pos, counts, box, cutoff, nbins)
        assert_same(got, want, st, ev)


def test_unwrapped_coordinates_use_general_minimum_image():
    rng = np.random.default_rng(4)
    counts, box = [200, 150], np.array([12.0, 12.0, 12.0])
    pos = ((rng.random((4, 350, 3)) * 5 - 2) * 12.0).astype(np.float32)  # spans [-2L, 3L)
    pos[1] = (rng.random((350, 3)) * 12.0).astype(np.float32)  # one wrapped frame in between
    got, st = gpu_counts(pos, counts, box, 5.9, 590)
    want, ev = oracle_counts(pos, counts, box, 5.9, 590)
    assert_same(got, want, st, ev)
    assert st.frames_general_minimage == 3


def test_centered_box_coordinates_stay_on_fast_path():
    rng = np.random.default_rng(5)
    counts, box = [400], np.array([20.0, 20.0, 20.0])
    pos = ((rng.random((3, 400, 3)) - 0.5) * 20.0).astype(np.float32)  # LAMMPS-style [-L/2, L/2)
    got, st = gpu_counts(pos, counts, box, 9.9, 990)
    want, ev = oracle_counts(pos, counts, box, 9.9, 990)
    assert_same(got, want, st, ev)
    assert st.frames_general_minimage == 0


def test_layouts_and_input_residency_agree():
    sysm = synthetic.nacl_melt(4, 6, seed=7)  # 512 atoms
    want, ev = oracle_counts(sysm.positions, sysm.counts, sysm.box, 9.0, 450)
    for layout in (_cuda.LAYOUT_FRAME_MAJOR, _cuda.LAYOUT_ATOM_MAJOR):
        for device_input in (False, True):
            for max_frames, n_submits in ((0, 1), (2, 1), (4, 3)):
                got, st = gpu_counts(sysm.positions, sysm.counts, sysm.box, 9.0, 450,
                                     layout=layout, device_input=device_input,
                                     max_frames=max_frames, n_submits=n_submits)
                assert_same(got, want, st, ev)


def test_distances_on_bin_edges():
    """Atoms on a line with spacing = bin width: every distance sits on (or one ulp around) a bin
    edge, the only place where the threshold table could disagree with sqrt + double divide."""
    nbins, cutoff, L = 400, 8.0, 40.0
    step = cutoff / nbins
    n = 1200
    base = np.zeros((n, 3), dtype=np.float32)
    base[:, 0] = (np.arange(n) * np.float32(step)).astype(np.float32) % np.float32(L)
    frames = []
    for k, eps in enumerate((0, 1, -1, 2, -2)):
        f = base.copy()
        f[:, 0] = np.nextafter(f[:, 0], np.float32(np.inf if eps > 0 else -np.inf)) if eps else f[:, 0]
        if abs(eps) == 2:
            f[::2, 0] = np.nextafter(f[::2, 0], np.float32(np.inf if eps > 0 else -np.inf))
        f[:, 1] = np.float32(0.25 * k)
        frames.append(np.clip(f, 0, np.nextafter(np.float32(L), np.float32(0))))
    pos = np.stack(frames).astype(np.float32)
    box = np.array([L, L, L])
    got, st = gpu_counts(pos, [n], box, cutoff, nbins, parity=False)
    want, ev = oracle_counts(pos, [n], box, cutoff, nbins, parity=False)
    assert_same(got, want, st, ev)
    # the two histogram formulas the survey discusses differ on such inputs; report, don't assert
    alt = rdf_oracle.rdf_counts_blocked(pos[:1].transpose(1, 0, 2), [n], box, cutoff, nbins,
                                        skip_first=False,
                                        histogram=rdf_oracle.tf_histogram_fixed_width_pydoc)
    prim = rdf_oracle.rdf_counts_blocked(pos[:1].transpose(1, 0, 2), [n], box, cutoff, nbins,
                                         skip_first=False)
    print("bins where kernel-formula and pydoc-formula differ:", int((alt != prim).sum()))


def test_pairs_around_half_box_and_cutoff():
    """Coordinates differences within a few ulps of +-L/2 (minimum-image switch) and distances within
    a few ulps of the cutoff."""
    rng = np.random.default_rng(11)
    L = np.float32(23.7)
    half = np.float32(L / 2)
    n = 600
    pos = np.zeros((2, n, 3), dtype=np.float32)
    a = (rng.random(n // 2) * 5).astype(np.float32)
    for f in range(2):
        off = half
        for _ in range(f * 3):
            off = np.nextafter(off, np.float32(np.inf))
        pos[f, :n // 2, 0] = a
        b = a + off
        jitter = rng.integers(-3, 4, size=n // 2)
        for k in range(n // 2):
            v = b[k]
            for _ in range(abs(int(jitter[k]))):
                v = np.nextafter(v, np.float32(np.inf if jitter[k] > 0 else -np.inf))
            b[k] = v
        pos[f, n // 2:, 0] = b
        pos[f, :, 1] = (rng.random(n) * 0.01).astype(np.float32)
    box = np.array([L, L, L], dtype=np.float64)
    cutoff = float(half) - 0.1
    got, st = gpu_counts(pos, [n // 2, n // 2], box, cutoff, 1175)
    want, ev = oracle_counts(pos, [n // 2, n // 2], box, cutoff, 1175)
    assert_same(got, want, st, ev)


def test_nan_and_empty_inputs():
    rng = np.random.default_rng(12)
    counts, box = [64, 64], np.array([9.0, 9.0, 9.0])
    pos = (rng.random((2, 128, 3)) * 9).astype(np.float32)
    pos[0, 5, 1] = np.nan
    pos[1, 100, :] = np.nan
    got, st = gpu_counts(pos, counts, box, 4.4, 100)
    with np.errstate(invalid="ignore"):
        want, ev = oracle_counts(pos, counts, box, 4.4, 100)
    assert_same(got, want)
    # zero frames, species with one atom (no pairs in parity mode), single atom system
    with _cuda.RDFHandle([1, 1], 10, 1.0, box) as h:
        h.submit(np.zeros((0, 2, 3), dtype=np.float32))
        h.submit(np.zeros((3, 2, 3), dtype=np.float32))
        assert h.counts().sum() == 0
    with _cuda.RDFHandle([2], 10, 1.0, box, parity=False) as h:
        h.submit(np.zeros((3, 2, 3), dtype=np.float32))
        c = h.counts()
        assert c[0, 0] == 3 and c.sum() == 3  # coincident atoms: d = 0 -> bin 0


def test_reset_and_accumulate():
    sysm = synthetic.nacl_melt(3, 4, seed=9)  # 216 atoms
    want, _ = oracle_counts(sysm.positions, sysm.counts, sysm.box, 9.0, 90)
    with _cuda.RDFHandle(sysm.counts, 90, 9.0, sysm.box) as h:
        h.submit(sysm.positions[:2].copy())
        h.submit(sysm.positions[2:].copy())
        assert_same(h.counts(), want)
        h.reset()
        assert h.counts().sum() == 0
        h.submit(sysm.positions)
        assert_same(h.counts(), want)
        lib_edges, s_cut = h.threshold_table()
    edges, sc = rdf_oracle.s_threshold_table(9.0, 90)
    assert (lib_edges == edges).all() and s_cut == sc


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("dense", [0, 1])
def test_every_kernel_variant(variant, dense, monkeypatch):
    monkeypatch.setenv("MDS_RDF_VARIANT", str(variant))
    monkeypatch.setenv("MDS_RDF_DENSE", str(dense))
    sysm = synthetic.nacl_melt(6, 3, seed=21)  # 1728 atoms
    got, st = gpu_counts(sysm.positions, sysm.counts, sysm.box, 12.0, 1200)
    want, ev = oracle_counts(sysm.positions, sysm.counts, sysm.box, 12.0, 1200)
    assert st.variant == variant
    assert_same(got, want, st, ev)


def test_cfg2_shape_nacl_13824_default_cutoff_and_bins():
    sysm = synthetic.nacl_melt(12, 2, seed=2)
    cutoff = sysm.box[0] / 2 - 0.1
    nbins = int(cutoff / 0.01)
    assert (sysm.n_atoms, nbins) == (13824, 3769)  # int(37.699999999999996 / 0.01), like the reference
    got, st = gpu_counts(sysm.positions, sysm.counts, sysm.box, cutoff, nbins)
    want, ev = oracle_counts(sysm.positions, sysm.counts, sysm.box, cutoff, nbins)
    assert_same(got, want, st, ev)
    got, st = gpu_counts(sysm.positions, sysm.counts, sysm.box, 10.0, 500)
    want, ev = oracle_counts(sysm.positions, sysm.counts, sysm.box, 10.0, 500)
    assert_same(got, want, st, ev)


def test_many_species_multi_pass_and_large_bins():
    """5 species x 5000 bins does not fit one CTA's shared memory: pair groups / global fallback."""
    rng = np.random.default_rng(13)
    counts, box = [120, 90, 60, 30, 100], np.array([30.0, 30.0, 30.0])
    pos = (rng.random((3, 400, 3)) * 30).astype(np.float32)
    got, st = gpu_counts(pos, counts, box, 14.9, 5000)
    want, ev = oracle_counts(pos, counts, box, 14.9, 5000)
    assert_same(got, want, st, ev)
    got, st = gpu_counts(pos, counts, box, 14.9, 60000)
    want, ev = oracle_counts(pos, counts, box, 14.9, 60000)
    assert_same(got, want, st, ev)


def _golden_cases():
    import json
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "rdf_golden.json")) as fh:
        return json.load(fh)["cases"]


@pytest.mark.parametrize("case", _golden_cases(), ids=lambda c: c["name"])
def test_committed_golden_vectors(case):
    """The fixtures under tests/golden/ (self-generated from the literal NumPy restatement)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden
    pos, box = make_golden.case_inputs(case)
    got, _ = gpu_counts(pos, case["counts"], box, case["cutoff"], case["nbins"], parity=True)
    assert_same(got, np.array(case["counts_parity"], dtype=np.uint64))
    got, _ = gpu_counts(pos, case["counts"], box, case["cutoff"], case["nbins"], parity=False,
                        layout=_cuda.LAYOUT_ATOM_MAJOR)
    assert_same(got, np.array(case["counts_corrected"], dtype=np.uint64))


@pytest.mark.parametrize("path", ["allpairs", "celllist", "celllist-one-kernel-build"])
def test_item_shards_add_up_to_the_full_histogram(path, monkeypatch):
    """mds_rdf_set_item_shard: strong scaling inside frames (fewer frames than GPUs).  The shards
    of one frame, summed, equal the unsharded histogram and the oracle's; so do their pair counts.
    (Third case: the slab of z layers of a shard as the one-kernel build of the cell lists cuts it.)"""
    if path.endswith("one-kernel-build"):
        monkeypatch.setenv("MDS_RDF_CELL_BUILD", "fused")
        path = "celllist"
    sysm = synthetic.lj_fluid(14, 2, seed=41) if path == "celllist" else synthetic.nacl_melt(6, 2, seed=41)
    cutoff, nbins = (2.9, 145) if path == "celllist" else (9.0, 300)
    want, ev = oracle_counts(sysm.positions, sysm.counts, sysm.box, cutoff, nbins)
    for count in (1, 3, 8):
        total = np.zeros_like(want)
        evaluated = equiv = 0.0
        with _cuda.RDFHandle(sysm.counts, nbins, cutoff, sysm.box, path=path) as h:
            for index in range(count):
                h.reset()
                h.set_item_shard(index, count)
                h.submit(sysm.positions)
                part = h.counts()
                st = h.stats()
                assert st.item_shards == count
                if count > 1:
                    assert 0 < part.sum() < want.sum()
                total += part
                evaluated += st.pairs_evaluated
                equiv += st.pairs_allpairs_equiv
        assert_same(total, want)
        assert abs(equiv - ev) < 1e-6 * ev
        if path == "allpairs":
            assert evaluated == ev
    with _cuda.RDFHandle(sysm.counts, nbins, cutoff, sysm.box, path=path) as h:
        with pytest.raises(_cuda.MdsCudaError, match="shard 3 of 3"):
            h.set_item_shard(3, 3)


def test_host_submit_paths_agree_and_reuse_their_buffers():
    """The three ways host frames reach the device give the oracle's histogram: page-locked memory
    (copied straight out of the caller's buffer), pageable memory one configuration per call (the
    reference's batching: coalesced in the library's pinned ring) and pageable (N, C, 3) blocks (sent
    as they are).  A pageable buffer may be overwritten as soon as the call returns; reset discards
    frames that are still waiting in the ring; tickets order the reuse of pinned buffers."""
    import torch
    sysm = synthetic.nacl_melt(3, 23, seed=77)  # 216 atoms, 23 frames
    cutoff, nbins = 8.0, 160
    want, _ = oracle_counts(sysm.positions, sysm.counts, sysm.box, cutoff, nbins)
    with _cuda.RDFHandle(sysm.counts, nbins, cutoff, sysm.box) as h:
        # pageable, one configuration per call, the SAME array reused for every call
        scratch = np.empty((sysm.n_atoms, 1, 3), dtype=np.float32)
        for f in range(sysm.n_frames):
            scratch[:, 0, :] = sysm.positions[f]
            h.submit(scratch, _cuda.LAYOUT_ATOM_MAJOR)
            scratch[...] = np.nan  # the library has its own copy by now
        assert_same(h.counts(), want)
        # reset drops what is still waiting in the ring
        h.reset()
        h.submit(sysm.positions[:5])
        h.reset()
        h.submit(np.ascontiguousarray(sysm.positions.transpose(1, 0, 2)), _cuda.LAYOUT_ATOM_MAJOR)
        assert_same(h.counts(), want)
        # page-locked ring of two buffers, refilled under tickets
        h.reset()
        ring = [torch.empty((4, sysm.n_atoms, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
        tickets = [None, None]
        assert h.host_ticket() >= 0
        for k, f0 in enumerate(range(0, sysm.n_frames, 4)):
            n = min(4, sysm.n_frames - f0)
            if tickets[k & 1] is not None:
                h.wait_host_ticket(tickets[k & 1])
            ring[k & 1][:n] = torch.from_numpy(sysm.positions[f0:f0 + n])
            h.submit(ring[k & 1][:n], pinned=True)
            tickets[k & 1] = h.host_ticket()
        assert tickets[0] != tickets[1]
        assert_same(h.counts(), want)
        with pytest.raises(_cuda.MdsCudaError, match="ticket"):
            h.wait_host_ticket(h.host_ticket() + 1)
