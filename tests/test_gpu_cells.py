"""GPU parity of the cell-list path (north-star kernel (b)): same integer histograms as the oracle's
all-pairs restatement, bit for bit — a cell list may only change WHICH pairs are distance-tested,
never a count.  Tolerance for the histograms is ZERO.

Small and medium cases are checked against the C oracle (oracle/rdf_oracle.c); at the full
BASELINE sizes (cfg3 100k atoms, cfg4 ~1M atoms) the oracle's O(N^2) loop is too slow for a test,
so one frame of cfg3 goes against the oracle and the large cases are checked through a
size-independent property: the cell-list kernel and the (independently written, oracle-checked)
all-pairs kernel must produce identical histograms.
"""
import os

import numpy as np
import pytest

from mdsuite_b200 import _cuda, synthetic
from oracle import c_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["1|fused", "2|fused", "2,1,2|fused", "4,1,1|fused", "1|split", "2,1,2|split",
                        "4,1,1|eager", "2"],
                ids=["K1", "K2", "K212", "K411", "K1-split-build", "K212-split-build",
                     "K411-eager-pack", "K2-auto-build"])
def cell_k(request, monkeypatch):
    """Stencil half-widths of the cell kernel (cells per cutoff length), forced through the
    library's tuning variable so that every case runs on the classical grid, on the half-width
    grid and on a mixed one.  The cell lists of these small cases are built by the one-kernel build
    (forced: by default only batches of at least one wave of its CTAs take it); "split-build":
    by pack / scan / scatter; "eager-pack": the one-kernel build also writes the all-pairs kernel's
    copy of every frame (default: only flagged frames are packed, afterwards); "auto-build": the
    library's own choice."""
    spec, _, build = request.param.partition("|")
    monkeypatch.setenv("MDS_RDF_CELL_K", spec)
    if build:
        monkeypatch.setenv("MDS_RDF_CELL_BUILD", build)
    k = [int(v) for v in spec.split(",")]
    return tuple(k * 3 if len(k) == 1 else k)


def expected_grid(box, cutoff, k):
    """Cells per dimension as mds_rdf_create picks them: k cells span >= 1.001 x cutoff."""
    return tuple(int(min(128 * k[d], np.floor(float(np.float32(box[d])) /
                                               (float(np.float32(cutoff)) * 1.001 / k[d]))))
                 for d in range(3))


def run(pos, counts, box, cutoff, nbins, path, parity=True, layout=_cuda.LAYOUT_FRAME_MAJOR,
        device_input=False, max_frames=0):
    import torch
    with _cuda.RDFHandle(counts, nbins, cutoff, box, parity=parity, path=path,
                         max_frames_in_flight=max_frames) as h:
        arr = pos if layout == _cuda.LAYOUT_FRAME_MAJOR else np.ascontiguousarray(
            pos.transpose(1, 0, 2))
        if device_input:
            h.submit_device(torch.from_numpy(arr).cuda(), layout)
        else:
            h.submit(arr, layout)
        out = h.counts()
        st = h.stats()
    if st.path == _cuda.PATH_CELLLIST and st.frames:
        # which build ran: one kernel for frame-major input whose (species, cell) table fits in
        # shared memory (rdf_cells.cu BUILD_MAX_KEYS), else pack / scan / scatter
        fits = len(counts) * int(np.prod(st.cells)) <= 48 * 1024
        mode = os.environ.get("MDS_RDF_CELL_BUILD")
        if mode == "split" or not fits:
            assert st.cell_build == 0
        elif mode in ("fused", "eager") and layout == _cuda.LAYOUT_FRAME_MAJOR:
            assert st.cell_build == 1
        # (atom-major input takes the one-kernel build only where a block of it is one frame;
        # without the variable the library decides by the size of the batch)
    return out, st


def oracle(pos, counts, box, cutoff, nbins, parity=True):
    with np.errstate(invalid="ignore"):
        out, ev = c_oracle.rdf_counts_c(pos, counts, box, cutoff, nbins, skip_first=parity,
                                        layout=c_oracle.FRAME_MAJOR)
    return out.astype(np.uint64), ev


def assert_same(got, want):
    diff = np.argwhere(got != want)
    assert diff.size == 0, f"{len(diff)} bins differ, first: {diff[:5].tolist()} " \
                           f"got {got[tuple(diff[0])]} want {want[tuple(diff[0])]}"


def cell_index(pos, box, n):
    """NumPy twin of cell_coord in rdf_cells.cu (exact IEEE float32 ops)."""
    out = []
    for k in range(3):
        u = (pos[..., k] / np.float32(box[k])).astype(np.float32)
        u = (u - np.floor(u)).astype(np.float32)
        c = np.trunc((u * np.float32(n[k])).astype(np.float32)).astype(np.int64)
        out.append(np.clip(c, 0, n[k] - 1))
    return (out[2] * n[1] + out[1]) * n[0] + out[0]


def expected_candidates(pos, counts, box, n, parity=True, k=(1, 1, 1)):
    """Pairs the cell kernel distance-tests: same-species pairs inside a cell + all pairs of
    atoms in distinct cells at most k cells apart, per species pair, summed over frames."""
    nx, ny, nz = n
    nc = nx * ny * nz
    first = 1 if parity else 0
    total = 0
    offs = np.concatenate([[0], np.cumsum(counts)])
    for f in range(pos.shape[0]):
        occ = []
        for s in range(len(counts)):
            p = pos[f, offs[s] + first:offs[s + 1]]
            c = cell_index(p, box, n) if len(p) else np.zeros(0, dtype=np.int64)
            occ.append(np.bincount(c, minlength=nc).reshape(nz, ny, nx).astype(np.int64))
        for a in range(len(counts)):
            for b in range(a, len(counts)):
                neigh = np.zeros_like(occ[b])
                for dz in range(-k[2], k[2] + 1):
                    for dy in range(-k[1], k[1] + 1):
                        for dx in range(-k[0], k[0] + 1):
                            neigh += np.roll(occ[b], (dz, dy, dx), axis=(0, 1, 2))
                if a == b:
                    # ordered neighbours incl. self: (sum n_c * neigh_c - sum n_c) / 2
                    total += int(((occ[a] * neigh).sum() - occ[a].sum()) // 2)
                else:
                    total += int((occ[a] * neigh).sum())
    return total


@pytest.mark.parametrize("parity", [True, False])
def test_lj_4096_vs_oracle(parity, cell_k):
    sysm = synthetic.lj_fluid(16, 3, seed=3)  # L = 16.93, cutoff 3 -> 5 (K = 1) / 11 (K = 2) cells
    got, st = run(sysm.positions, sysm.counts, sysm.box, 3.0, 300, "celllist", parity=parity)
    want, _ = oracle(sysm.positions, sysm.counts, sysm.box, 3.0, 300, parity=parity)
    assert st.path == _cuda.PATH_CELLLIST and st.cell_stencil == cell_k
    assert st.cells == expected_grid(sysm.box, 3.0, cell_k)
    assert st.cells == tuple({1: 5, 2: 11, 4: 22}[k] for k in cell_k)
    assert_same(got, want)
    assert st.pairs_binned == float(want.sum())
    assert st.pairs_evaluated == expected_candidates(sysm.positions, sysm.counts, sysm.box,
                                                     st.cells, parity, cell_k)
    assert st.pairs_allpairs_equiv == sysm.pairs_per_frame(parity) * 3
    assert st.frames_general_minimage == 0 and st.frames_far == 0


def test_single_pair_launches_with_separate_tables_vs_oracle(monkeypatch):
    """A launch with one species pair keeps threshold table and histogram interleaved in one
    shared-memory array (one address per pair in the binning sequence); MDS_RDF_CELL_TWO_TABLES
    switches such launches back to the two arrays multi-pair launches use.  Same counts."""
    sysm = synthetic.lj_fluid(16, 3, seed=32)
    want, _ = oracle(sysm.positions, sysm.counts, sysm.box, 2.5, 250)
    inter, _ = run(sysm.positions, sysm.counts, sysm.box, 2.5, 250, "celllist")
    monkeypatch.setenv("MDS_RDF_CELL_TWO_TABLES", "1")
    monkeypatch.setenv("MDS_RDF_FLUSH_THRESHOLD", "8192")
    two, _ = run(sysm.positions, sysm.counts, sysm.box, 2.5, 250, "celllist")
    monkeypatch.delenv("MDS_RDF_CELL_TWO_TABLES")
    flushed, _ = run(sysm.positions, sysm.counts, sysm.box, 2.5, 250, "celllist")  # interleaved + flushes
    assert_same(inter, want)
    assert_same(two, want)
    assert_same(flushed, want)


def test_water_three_pairs_vs_oracle(cell_k):
    sysm = synthetic.spc_water(12, 2, seed=4)  # 5184 atoms, L = 37.26, cutoff 8 -> 4 / 9 cells
    got, st = run(sysm.positions, sysm.counts, sysm.box, 8.0, 800, "celllist")
    want, _ = oracle(sysm.positions, sysm.counts, sysm.box, 8.0, 800)
    assert st.cells == expected_grid(sysm.box, 8.0, cell_k)
    assert st.cells == tuple({1: 4, 2: 9, 4: 18}[k] for k in cell_k)
    assert_same(got, want)
    assert st.pairs_evaluated == expected_candidates(sysm.positions, sysm.counts, sysm.box,
                                                     st.cells, True, cell_k)


def test_noncubic_three_species_minimum_grid(cell_k):
    """3 x 4 x 5 cutoff lengths (3 is the minimum: 3 or 6 cells, the stencil covers every cell of
    the dimension and the x window always wraps), a species with one atom (no pairs in parity
    mode) and uneven species sizes."""
    rng = np.random.default_rng(31)
    counts, box = [1500, 1, 900], np.array([9.1, 12.2, 15.3])
    pos = (rng.random((4, sum(counts), 3)) * box).astype(np.float32)
    for parity in (True, False):
        got, st = run(pos, counts, box, 3.0, 250, "celllist", parity=parity)
        want, _ = oracle(pos, counts, box, 3.0, 250, parity=parity)
        assert st.cells == tuple(n * k for n, k in zip((3, 4, 5), cell_k))
        assert_same(got, want)
        assert st.pairs_evaluated == expected_candidates(pos, counts, box, st.cells, parity, cell_k)


def test_unwrapped_centered_and_far_frames(cell_k):
    rng = np.random.default_rng(32)
    counts, box = [1300, 1100], np.array([14.0, 14.0, 14.0])
    n = sum(counts)
    pos = np.empty((5, n, 3), dtype=np.float32)
    pos[0] = (rng.random((n, 3)) * 14.0)                      # wrapped
    pos[1] = ((rng.random((n, 3)) * 5 - 2) * 14.0)            # [-2L, 3L): general minimum image
    pos[2] = ((rng.random((n, 3)) - 0.5) * 14.0)              # [-L/2, L/2): fast path, window B
    pos[3] = ((rng.random((n, 3)) * 40 - 20) * 14.0)          # +-20 L: handed to all-pairs
    pos[4] = pos[0]
    pos[4, 7, 2] = np.float32(1e30)                           # one atom absurdly far out
    got, st = run(pos, counts, box, 4.0, 400, "celllist")
    want, _ = oracle(pos, counts, box, 4.0, 400)
    assert_same(got, want)
    assert st.frames_far == 2
    assert st.frames_general_minimage == 3  # frames 1, 3, 4 leave both windows


def test_nan_inf_coordinates(cell_k):
    rng = np.random.default_rng(33)
    counts, box = [1200, 1200], np.array([12.0, 12.0, 12.0])
    pos = (rng.random((3, 2400, 3)) * 12).astype(np.float32)
    pos[0, 5, 1] = np.nan
    pos[1, 1500, :] = np.nan
    pos[2, 9, 0] = np.inf
    got, st = run(pos, counts, box, 3.5, 100, "celllist")
    want, _ = oracle(pos, counts, box, 3.5, 100)
    assert_same(got, want)


def test_clustered_atoms_row_passes_and_histogram_flush(monkeypatch, cell_k):
    """Everything in a few cells: cells with dozens to hundreds of atoms (segments that have to be
    cut down to single cells and pencil subsets to fit the staging buffers), most cells empty; a
    tiny flush threshold makes warps empty the shared histogram while others count."""
    monkeypatch.setenv("MDS_RDF_FLUSH_THRESHOLD", "4096")
    rng = np.random.default_rng(34)
    counts, box = [700, 500], np.array([30.0, 30.0, 30.0])
    pos = (rng.normal(0.0, 1.2, (3, 1200, 3)) + np.array([10.0, 29.5, 0.3])).astype(np.float32)
    pos = (pos - np.floor(pos / 30.0) * 30.0).astype(np.float32)
    pos = np.where(pos >= 30.0, 0.0, pos).astype(np.float32)
    got, st = run(pos, counts, box, 2.9, 290, "celllist")
    want, _ = oracle(pos, counts, box, 2.9, 290)
    assert st.cells == tuple({1: 10, 2: 20, 4: 41}[k] for k in cell_k)
    assert_same(got, want)
    if st.frames_far == 0:  # (a grid whose fullest cell cannot be staged hands the frame over)
        assert st.pairs_evaluated == expected_candidates(pos, counts, box, st.cells, True, cell_k)


@pytest.mark.parametrize("cap", [256, 1536])
def test_crowded_frames_fall_back_to_all_pairs(monkeypatch, cap):
    """A frame with more atoms in one cell than the pencil kernel can stage goes to the all-pairs
    kernel as a whole (counted in frames_far); its neighbours in the batch stay on the cell path."""
    monkeypatch.setenv("MDS_RDF_CELL_CAP", str(cap))
    rng = np.random.default_rng(38)
    counts, box = [1500, 900], np.array([24.0, 24.0, 24.0])
    pos = (rng.random((3, 2400, 3)) * 24.0).astype(np.float32)
    pos[1, :1500] = (12.0 + rng.random((1500, 3)) * 0.5).astype(np.float32)  # one tiny blob
    got, st = run(pos, counts, box, 3.0, 150, "celllist")
    want, _ = oracle(pos, counts, box, 3.0, 150)
    assert_same(got, want)
    assert st.frames_far == 1


def test_layouts_residency_and_batching_agree(cell_k):
    sysm = synthetic.lj_fluid(12, 7, seed=35)  # 1728 atoms, L = 12.7
    want, _ = oracle(sysm.positions, sysm.counts, sysm.box, 2.5, 125)
    for layout in (_cuda.LAYOUT_FRAME_MAJOR, _cuda.LAYOUT_ATOM_MAJOR):
        for device_input in (False, True):
            for max_frames in (0, 1, 3):
                got, st = run(sysm.positions, sysm.counts, sysm.box, 2.5, 125, "celllist",
                              layout=layout, device_input=device_input, max_frames=max_frames)
                assert st.path == _cuda.PATH_CELLLIST
                assert_same(got, want)


def test_many_species_pair_groups_and_global_histogram(cell_k):
    rng = np.random.default_rng(36)
    counts, box = [500, 400, 300, 200, 600], np.array([20.0, 20.0, 20.0])
    pos = (rng.random((2, 2000, 3)) * 20).astype(np.float32)
    for nbins in (5000, 60000):  # pair groups; then bins in global memory
        got, st = run(pos, counts, box, 4.9, nbins, "celllist")
        want, _ = oracle(pos, counts, box, 4.9, nbins)
        assert_same(got, want)


def test_path_selection_and_forced_path_errors():
    rng = np.random.default_rng(37)
    box = np.array([40.0, 40.0, 40.0])
    pos = (rng.random((1, 3000, 3)) * 40).astype(np.float32)
    _, st = run(pos, [3000], box, 5.0, 50, None)      # 7^3 cells: auto picks the cell list
    # 3000 atoms in 7^3 cutoff cells: 1.1 per half-width cell -> the classical grid
    assert st.path == _cuda.PATH_CELLLIST and st.cells == (7, 7, 7) and st.cell_stencil == (1, 1, 1)
    _, st = run(pos, [3000], box, 12.0, 50, None)     # 3^3 cells: not worth it
    assert st.path == _cuda.PATH_ALLPAIRS and st.cells == (0, 0, 0)
    _, st = run(pos, [3000], box, 12.0, 50, "celllist")
    assert st.path == _cuda.PATH_CELLLIST and st.cells == (6, 6, 6) and st.cell_stencil == (2, 2, 2)
    with pytest.raises(_cuda.MdsCudaError, match="cell-list path needs"):
        run(pos, [3000], box, 14.0, 50, "celllist")   # 40 / 14 < 3 cells
    a, _ = run(pos, [3000], box, 12.0, 120, "celllist")
    b, _ = run(pos, [3000], box, 12.0, 120, "allpairs")
    assert_same(a, b)


def test_cfg5_ideal_gas_bins_sweep_cells_equals_allpairs():
    sysm = synthetic.ideal_gas(50_000, 2, seed=5)
    for nbins in (100, 1000, 5000):
        a, sa = run(sysm.positions, sysm.counts, sysm.box, 10.0, nbins, "celllist")
        b, sb = run(sysm.positions, sysm.counts, sysm.box, 10.0, nbins, "allpairs")
        assert sa.cells == (19, 19, 19) and sa.cell_stencil == (2, 2, 2)
        assert_same(a, b)
        assert sa.pairs_binned == sb.pairs_binned
    want, _ = oracle(sysm.positions[:1], sysm.counts, sysm.box, 10.0, 1000)
    got, _ = run(sysm.positions[:1], sysm.counts, sysm.box, 10.0, 1000, "celllist")
    assert_same(got, want)


def test_cfg3_lj_100k_one_frame_vs_oracle_and_allpairs():
    sysm = synthetic.lj_fluid(46, 3, seed=3)  # 97 336 atoms (46^3), cfg3's density and cutoff
    got, st = run(sysm.positions, sysm.counts, sysm.box, 3.0, 300, None)
    # 25 atoms per cutoff-sized cell: 3 cells per cutoff along x, cutoff-sized cells in y and z
    assert st.path == _cuda.PATH_CELLLIST and st.cells == (48, 16, 16) and st.cell_stencil == (3, 1, 1)
    ap, _ = run(sysm.positions, sysm.counts, sysm.box, 3.0, 300, "allpairs")
    assert_same(got, ap)
    want, _ = oracle(sysm.positions[:1], sysm.counts, sysm.box, 3.0, 300)
    one, _ = run(sysm.positions[:1], sysm.counts, sysm.box, 3.0, 300, "celllist")
    assert_same(one, want)


def test_long_batches_take_the_one_kernel_build_with_several_frames_per_cta():
    """The library's own choice: a batch of at least one wave of build CTAs (one per SM) goes through
    the one-kernel build, each CTA taking one frame after another (its shared-memory table is
    cleared in between); shorter batches through pack / scan / scatter.  Same counts either way."""
    sysm = synthetic.lj_fluid(12, 7, seed=36)  # 1728 atoms
    want, _ = oracle(sysm.positions, sysm.counts, sysm.box, 2.5, 125)
    reps = 60                                   # 420 frames: one batch, ~3 frames per CTA
    pos = np.tile(sysm.positions, (reps, 1, 1))
    pos[211, 5, 1] -= 8.5 * sysm.box[1]         # one far-out frame in the middle (frame 211 = base 1)
    far, _ = oracle(pos[211:212], sysm.counts, sysm.box, 2.5, 125)
    near, _ = oracle(sysm.positions[1:2], sysm.counts, sysm.box, 2.5, 125)
    got, st = run(pos, sysm.counts, sysm.box, 2.5, 125, "celllist", device_input=True)
    assert st.cell_build == 1 and st.frames_far == 1
    assert_same(got, want * reps - near + far)
    short, st = run(pos[:70], sysm.counts, sysm.box, 2.5, 125, "celllist", device_input=True)
    assert st.cell_build == 0
    assert_same(short, want * 10)


@pytest.mark.parametrize("mode", ["fused", "eager"])
def test_cfg3_lj_100k_one_kernel_build_equals_split_build(monkeypatch, mode):
    """The two ways of building the cell lists at cfg3's size (12 288 keys, 8 atoms per cell): same
    histograms, same candidate-pair count; one far-out atom sends its frame to the all-pairs kernel
    through the copy that is packed afterwards (default) or by the build itself ("eager")."""
    sysm = synthetic.lj_fluid(46, 3, seed=3)
    pos = sysm.positions.copy()
    pos[1, 17, 0] += 9.0 * sysm.box[0]  # frame 1: beyond 8 box lengths -> MDS_FRAME_FAR
    monkeypatch.setenv("MDS_RDF_CELL_BUILD", "split")
    want, st_split = run(pos, sysm.counts, sysm.box, 3.0, 300, "celllist")
    monkeypatch.setenv("MDS_RDF_CELL_BUILD", mode)
    got, st = run(pos, sysm.counts, sysm.box, 3.0, 300, "celllist")
    assert st.cell_build == 1 and st_split.cell_build == 0
    assert st.frames_far == 1 and st_split.frames_far == 1
    assert st.pairs_evaluated == st_split.pairs_evaluated
    assert_same(got, want)


def test_cfg4_water_1m_atoms_cells_equals_allpairs():
    sysm = synthetic.spc_water(69, 1, seed=4)  # 985 527 atoms, L = 214.3
    got, st = run(sysm.positions, sysm.counts, sysm.box, 8.0, 800, None)
    assert st.path == _cuda.PATH_CELLLIST and st.cells == (107, 26, 26) and st.cell_stencil == (4, 1, 1)
    ap, sa = run(sysm.positions, sysm.counts, sysm.box, 8.0, 800, "allpairs")
    assert_same(got, ap)
    assert st.pairs_allpairs_equiv == sa.pairs_evaluated == sysm.pairs_per_frame(True)
    assert st.pairs_evaluated < 0.01 * sa.pairs_evaluated
