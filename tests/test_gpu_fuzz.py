"""Seeded fuzzing of both pair kernels against the C oracle: random species tables (including empty
and single-atom species), non-cubic boxes, cutoffs from tiny to beyond half the box, bin counts
from 1 to thousands, wrapped / centred / unwrapped coordinates, both layouts, both parity modes.
Integer histograms must be identical (tolerance zero)."""
import os

import numpy as np
import pytest

from mdsuite_b200 import _cuda
from oracle import c_oracle

pytestmark = pytest.mark.gpu
EXTRA = int(os.environ.get("MDS_FUZZ_EXTRA", "0"))  # more seeds for an occasional long hunt


def _case(rng, cells: bool):
    n_species = int(rng.integers(1, 5))
    if cells:
        counts = [int(rng.integers(0, 1500)) for _ in range(n_species)]
        counts[int(rng.integers(0, n_species))] += 1100  # the auto rule needs >= 1024 atoms
        box = rng.uniform(18.0, 40.0, 3)
        cutoff = float(rng.uniform(1.5, box.min() / 4.2))
    else:
        counts = [int(rng.integers(0, 400)) for _ in range(n_species)]
        if sum(counts) < 2:
            counts[0] = 5
        box = rng.uniform(6.0, 25.0, 3)
        cutoff = float(rng.uniform(0.3, 0.9 * box.max()))
    nbins = int(rng.choice([1, 2, 7, 50, 333, 1000, 4097]))
    frames = int(rng.integers(1, 5))
    n = sum(counts)
    mode = int(rng.integers(0, 4))
    u = rng.random((frames, n, 3))
    if mode == 0:
        pos = u * box                      # wrapped
    elif mode == 1:
        pos = (u - 0.5) * box              # centred
    elif mode == 2:
        pos = (u * 6 - 3) * box            # unwrapped: general minimum image
    else:
        pos = u * box
        pos[0] = (u[0] * 3 - 1) * box      # mixed: only the first frame unwrapped
    return dict(counts=counts, box=box, cutoff=cutoff, nbins=nbins, pos=pos.astype(np.float32),
                parity=bool(rng.integers(0, 2)), atom_major=bool(rng.integers(0, 2)),
                max_frames=int(rng.choice([0, 1, 2])))


def _run(case, path):
    pos = case["pos"]
    with _cuda.RDFHandle(case["counts"], case["nbins"], case["cutoff"], case["box"],
                         parity=case["parity"], path=path,
                         max_frames_in_flight=case["max_frames"]) as h:
        if case["atom_major"]:
            h.submit(np.ascontiguousarray(pos.transpose(1, 0, 2)), _cuda.LAYOUT_ATOM_MAJOR)
        else:
            h.submit(pos)
        got = h.counts()
        st = h.stats()
    want, ev = c_oracle.rdf_counts_c(pos, case["counts"], case["box"], case["cutoff"], case["nbins"],
                                     skip_first=case["parity"], layout=c_oracle.FRAME_MAJOR)
    return got.astype(np.int64), want, st, ev


@pytest.mark.parametrize("seed", range(40 + EXTRA))
def test_allpairs_kernel_random_shapes(seed):
    case = _case(np.random.default_rng(1000 + seed), cells=False)
    got, want, st, ev = _run(case, "allpairs")
    assert np.array_equal(got, want), {k: v for k, v in case.items() if k != "pos"}
    assert st.pairs_evaluated == ev


@pytest.mark.parametrize("seed", range(30 + EXTRA))
def test_cell_list_kernel_random_shapes(seed):
    case = _case(np.random.default_rng(2000 + seed), cells=True)
    got, want, st, ev = _run(case, "celllist")
    assert st.path == _cuda.PATH_CELLLIST
    assert np.array_equal(got, want), {k: v for k, v in case.items() if k != "pos"}
    assert st.pairs_allpairs_equiv == ev
