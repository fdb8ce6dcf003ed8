"""
TEST INFRASTRUCTURE / CPU BASELINE ONLY — op-by-op torch-CPU restatement of the reference's RDF
batch loop, used by bench.py as the "reference-equivalent CPU" arm (TensorFlow is not importable
in this image).  Parity: see oracle/rdf_oracle.py (control flow pinned, leaf arithmetic restated).

It keeps the reference's cost structure, not just its results: one materialised tensor per
TensorFlow op, the (2, P) index tensor rebuilt for every atom-minibatch of every batch
(mdsuite/calculators/radial_distribution_function.py:441, utils/linalg.py:116-121), one full
masking pass per species pair (:637-640), and torch's intra-op thread pool standing in for
TensorFlow's.  Results are bit-identical to oracle/rdf_oracle.py (tests/test_oracle.py).
"""
from __future__ import annotations

import itertools
from typing import Sequence

import numpy as np
import torch


def get_partial_triu_indices(n_atoms: int, m_atoms: int, idx: int) -> torch.Tensor:
    """utils/linalg.py:102-122 — ones -> band_part -> ~ -> where -> int32 -> transpose."""
    bool_mat = torch.ones((m_atoms, n_atoms), dtype=torch.bool)
    in_band = torch.tril(bool_mat, diagonal=idx)  # band_part(x, -1, idx)
    indices = torch.nonzero(~in_band)  # tf.where
    return indices.to(torch.int32).T.contiguous()


def get_dij(indices, positions_tensor, atoms, box_array) -> torch.Tensor:
    """calculators/radial_distribution_function.py:647-689 (+ utils/linalg.py:99)."""
    cols = indices[1].long()
    rows = indices[0].long()
    _positions = positions_tensor.index_select(0, cols)
    atoms_position = atoms.index_select(0, rows)
    r_ij = _positions - atoms_position
    q = r_ij / box_array
    n = torch.round(q)  # half to even, like tf.math.rint
    r_ij = r_ij - n * box_array
    sq = r_ij * r_ij
    s = (sq[..., 0] + sq[..., 1]) + sq[..., 2]  # Eigen's inner-dim reducer order
    # torch.sqrt(float32) on CPU is NOT correctly rounded (≈0.7 % of results are 1 ulp off,
    # measured here); TensorFlow/Eigen use sqrtps, which is.  NumPy's sqrt is exact too.
    return torch.from_numpy(np.sqrt(s.numpy()))


def histogram_fixed_width(values: torch.Tensor, lo: float, hi: float, nbins: int) -> torch.Tensor:
    """TF CPU HistogramFixedWidth kernel formula (see oracle/rdf_oracle.py)."""
    lo32, hi32 = np.float32(lo), np.float32(hi)
    step = float(np.float64(np.float32(hi32 - lo32)) / np.float64(nbins))
    shifted = (torch.clamp_min(values, float(lo32)) - float(lo32)).to(torch.float64)
    # a scalar divisor would be turned into a multiplication by 1/step; keep the true division
    idx = torch.clamp_max(shifted / torch.full_like(shifted, step), float(nbins - 1)).to(torch.int64)
    return torch.bincount(idx, minlength=nbins)


def bin_minibatch(start, stop, indices_t, d_ij, cutoff: float, nbins: int) -> torch.Tensor:
    """calculators/radial_distribution_function.py:616-645."""
    mask_1 = (indices_t[:, 0] > start[0]) & (indices_t[:, 0] < stop[0])
    mask_2 = (indices_t[:, 1] > start[1]) & (indices_t[:, 1] < stop[1])
    values_species = d_ij[mask_1 & mask_2]
    values = values_species[values_species < float(np.float32(cutoff))]
    return histogram_fixed_width(values, 0.0, cutoff, nbins)


def rdf_counts(positions_tensor, particles_list: Sequence[int], box_array, cutoff: float,
               number_of_bins: int, minibatch: int = -1) -> np.ndarray:
    """One batch of HOT LOOP 1 (:846-885).  positions_tensor: (N, C, 3) float32 (reference layout).
    Returns (n_pairs, B) int64 and the number of (row, col, frame) triples evaluated."""
    pos = torch.as_tensor(np.ascontiguousarray(positions_tensor, dtype=np.float32))
    box = torch.as_tensor(np.asarray(box_array, dtype=np.float32))
    n_atoms, n_conf = pos.shape[0], pos.shape[1]
    if minibatch in (-1, None):
        minibatch = n_atoms
    pairs = list(itertools.combinations_with_replacement(range(len(particles_list)), 2))
    rdf = torch.zeros((len(pairs), number_of_bins), dtype=torch.int64)
    evaluated = 0
    start_batch = 0
    for lo in range(0, n_atoms, minibatch):
        atoms = pos[lo:lo + minibatch]
        m = atoms.shape[0]
        indices = get_partial_triu_indices(n_atoms, m, start_batch)
        d_ij = get_dij(indices, pos, atoms, box)
        evaluated += indices.shape[1] * n_conf
        indices_t = indices.T
        for p, (a, b) in enumerate(pairs):
            start_ = (sum(particles_list[:a]) - start_batch, sum(particles_list[:b]))
            stop_ = (start_[0] + particles_list[a], start_[1] + particles_list[b])
            rdf[p] += bin_minibatch(start_, stop_, indices_t, d_ij, cutoff, number_of_bins)
        start_batch += m
    return rdf.numpy(), evaluated
