"""
TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of the MDSuite RDF hot path.

    PARITY PINNED FOR CONTROL FLOW AND NORMALISATION, UNPINNED FOR LEAF ARITHMETIC.
    The reference (zincware/MDSuite 0.2.0) delegates its arithmetic to TensorFlow, which is
    neither vendored under /root/reference nor importable in this image, and the reference
    holds no golden vectors for this path (its only RDF test downloads JSON from GitHub and
    compares at decimal=1).  What can be done here is done: tests/golden/make_reference_golden.py
    executes the reference's own run_calculator (calculators/radial_distribution_function.py,
    utils/linalg.py, utils/meta_functions.py, loaded from /root/reference unmodified) over a
    NumPy stand-in for its ~20 TensorFlow leaf ops, and commits the histograms and the
    normalised x / y as tests/golden/reference_golden.json.  This file, the C twin and the
    product's host logic are checked against those (tests/test_reference_golden.py): defaults,
    sampling, index tensors, the strict-'>' species masks (quirk Q1), accumulation, prefactor,
    ideal-gas correction, x axis are the reference's.  The float32 arithmetic of the leaf ops
    (subtract, divide, rint, multiply, norm, less, histogram_fixed_width) is TensorFlow's and
    remains a restatement of SURVEY.md §8c — in the shim exactly as here.
    tests/golden/rdf_golden.json (tests/golden/make_golden.py) is SELF-GENERATED from this file
    and kept because its cases are shared with the reference-run fixtures.

Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of
bench.py may import this module.  The product (mdsuite_b200) never does.

Every function cites the reference file:line it follows (paths relative to
/root/reference/mdsuite/).  Arithmetic contract: float32, every op rounded
separately (NumPy never contracts to FMA), the histogram bin index in float64 the way
TensorFlow's CPU HistogramFixedWidth kernel computes it.
"""
from __future__ import annotations

import itertools
import warnings
from typing import Dict, List, Sequence, Tuple

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------
# utils/linalg.py
# --------------------------------------------------------------------------------------
def get_partial_triu_indices(n_atoms: int, m_atoms: int, idx: int) -> np.ndarray:
    """utils/linalg.py:102-122.

    ``~band_part(ones((m, N)), -1, idx)`` keeps (row, col) with ``col - row > idx``;
    ``tf.where`` lists them row-major; cast to int32 and transposed to (2, P).
    """
    rows = np.arange(m_atoms, dtype=np.int64)[:, None]
    cols = np.arange(n_atoms, dtype=np.int64)[None, :]
    in_band = (cols - rows) <= idx  # num_lower=-1 keeps the whole lower part
    indices = np.argwhere(~in_band)  # row-major like tf.where
    return indices.astype(np.int32).T


def apply_minimum_image(r_ij: np.ndarray, box_array: np.ndarray) -> np.ndarray:
    """utils/linalg.py:84-99 — ``r - rint(r / box) * box``; four separately rounded f32 ops."""
    assert r_ij.dtype == F32 and box_array.dtype == F32
    q = r_ij / box_array  # RealDiv
    n = np.rint(q)  # Rint, half-to-even
    nb = n * box_array  # Mul
    return r_ij - nb  # Sub


def apply_system_cutoff(tensor: np.ndarray, cutoff: np.float32) -> np.ndarray:
    """utils/linalg.py:125-136 — keep ``d < cutoff`` (strict, f32), flattened."""
    assert tensor.dtype == F32
    return tensor[tensor < F32(cutoff)]


# --------------------------------------------------------------------------------------
# TensorFlow library ops that the path calls (restated; TF source is not under /root/reference)
# --------------------------------------------------------------------------------------
def tf_norm_last_axis(r_ij: np.ndarray) -> np.ndarray:
    """``tf.linalg.norm(r, axis=-1)`` = sqrt(reduce_sum(r * r, -1)).

    Products are materialised, then Eigen's inner-most-dim reducer adds the three scalars
    left to right starting from 0: ((0 + x²) + y²) + z².  sqrt is correctly rounded.
    """
    sq = r_ij * r_ij
    s = (sq[..., 0] + sq[..., 1]) + sq[..., 2]
    return np.sqrt(s)


def tf_histogram_fixed_width(values: np.ndarray, value_range, nbins: int) -> np.ndarray:
    """TF CPU kernel ``HistogramFixedWidthFunctor<CPUDevice,float,int32>``
    (tensorflow/core/kernels/histogram_op.cc, recalled from the TF 2.x source):

        step = double(hi - lo) / double(nbins)                (hi - lo subtracted in float32)
        bin  = int32(min(double(max(v, lo) - lo) / step, double(nbins - 1)))

    Returns int32 counts like the reference (Q5).
    """
    values = np.asarray(values, dtype=F32).ravel()
    lo, hi = F32(value_range[0]), F32(value_range[1])
    step = np.float64(F32(hi - lo)) / np.float64(nbins)
    shifted = (np.maximum(values, lo) - lo).astype(np.float64)
    idx = np.minimum(shifted / step, np.float64(nbins - 1)).astype(np.int32)
    return np.bincount(idx, minlength=nbins).astype(np.int32)


def tf_histogram_fixed_width_pydoc(values: np.ndarray, value_range, nbins: int) -> np.ndarray:
    """Secondary formula (the one the Python docstring of tf.histogram_fixed_width describes):
    ``floor(nbins * (v - lo) / (hi - lo))`` in float32, clipped to [0, nbins-1].
    Differs from the kernel formula only for values within ~1 ulp of a bin edge.
    """
    values = np.asarray(values, dtype=F32).ravel()
    lo, hi = F32(value_range[0]), F32(value_range[1])
    scaled = (values - lo) / F32(hi - lo)
    idx = np.floor(F32(nbins) * scaled).astype(np.int64)
    idx = np.clip(idx, 0, nbins - 1)
    return np.bincount(idx, minlength=nbins).astype(np.int32)


# --------------------------------------------------------------------------------------
# calculators/radial_distribution_function.py — hot loop
# --------------------------------------------------------------------------------------
def get_dij(indices, positions_tensor, atoms, box_array) -> np.ndarray:
    """calculators/radial_distribution_function.py:647-689.

    positions_tensor (N, C, 3) f32, atoms (m, C, 3) f32, indices (2, P) int32 → d_ij (P, C) f32.
    """
    _positions = positions_tensor[indices[1]]  # tf.gather(axis=0)         :667
    atoms_position = atoms[indices[0]]  #                                   :672
    r_ij = _positions - atoms_position  #                                   :676
    if box_array is not None:
        r_ij = apply_minimum_image(r_ij, box_array)  #                      :682
    return tf_norm_last_axis(r_ij)  #                                       :686


def bin_minibatch(start, stop, indices, d_ij, bin_range, number_of_bins, cutoff,
                  histogram=tf_histogram_fixed_width) -> np.ndarray:
    """calculators/radial_distribution_function.py:616-645 — strict ``>`` lower bounds (Q1)."""
    mask_1 = (indices[:, 0] > start[0]) & (indices[:, 0] < stop[0])  # :637
    mask_2 = (indices[:, 1] > start[1]) & (indices[:, 1] < stop[1])  # :638
    values_species = d_ij[mask_1 & mask_2]  # boolean_mask(axis=0)      :640
    values = apply_system_cutoff(values_species, cutoff)  #              :641
    return histogram(values, bin_range, number_of_bins)  #               :642-644


def species_pairs(n_species: int) -> List[Tuple[int, int]]:
    """calculators/radial_distribution_function.py:272-275, :502 — combinations_with_replacement."""
    return list(itertools.combinations_with_replacement(range(n_species), 2))


def compute_species_values(indices, start_batch, d_ij, particles_list, number_of_bins, cutoff,
                           histogram=tf_histogram_fixed_width) -> np.ndarray:
    """calculators/radial_distribution_function.py:470-524 → (n_pairs, B) int32."""
    indices_t = indices.T  # :499
    bin_range = np.array([0, cutoff], dtype=F32)  # :520 (self.bin_range = [0, cutoff], :260)
    out = []
    for a, b in species_pairs(len(particles_list)):
        start_ = np.array([sum(particles_list[:a]) - start_batch, sum(particles_list[:b])])
        stop_ = start_ + np.array([particles_list[a], particles_list[b]])
        out.append(bin_minibatch(start_, stop_, indices_t, d_ij, bin_range, number_of_bins,
                                 F32(cutoff), histogram=histogram))
    return np.stack(out).astype(np.int32)


def run_minibatch_loop(atoms, stop, n_atoms, minibatch_start, positions_tensor, box_array,
                       particles_list, number_of_bins, cutoff, histogram=tf_histogram_fixed_width):
    """calculators/radial_distribution_function.py:422-468."""
    atoms_per_batch = atoms.shape[0]
    stop = stop + atoms_per_batch
    indices = get_partial_triu_indices(n_atoms, atoms_per_batch, minibatch_start)
    d_ij = get_dij(indices, positions_tensor, atoms, box_array.astype(F32))
    minibatch_rdf = compute_species_values(indices, minibatch_start, d_ij, particles_list,
                                           number_of_bins, cutoff, histogram=histogram)
    return minibatch_rdf, stop, stop


def rdf_counts(positions_tensor: np.ndarray, particles_list: Sequence[int], box_array,
               cutoff: float, number_of_bins: int, minibatch: int = -1,
               histogram=tf_histogram_fixed_width, wrap_int32: bool = True) -> np.ndarray:
    """Body of HOT LOOP 1 for one batch (calculators/radial_distribution_function.py:846-885).

    positions_tensor : (N, C, 3) float32 — species concatenated on axis 0, configurations on
                       axis 1 (the layout ``_format_data`` :535-563 produces).
    Returns (n_pairs, B) int32 counts (int32 adds wrap like the reference, Q5) — or int64
    when ``wrap_int32`` is False.
    """
    positions_tensor = np.ascontiguousarray(positions_tensor, dtype=F32)
    box_array = np.asarray(box_array, dtype=F32)
    n_atoms = positions_tensor.shape[0]
    if minibatch in (-1, None):
        minibatch = n_atoms
    n_pairs = len(species_pairs(len(particles_list)))
    acc_dtype = np.int32 if wrap_int32 else np.int64
    rdf = np.zeros((n_pairs, number_of_bins), dtype=acc_dtype)
    minibatch_start, stop = 0, 0
    for lo in range(0, n_atoms, minibatch):  # Dataset.from_tensor_slices(...).batch(m)  :855,:868
        atoms = positions_tensor[lo:lo + minibatch]
        mb, minibatch_start, stop = run_minibatch_loop(
            atoms, stop, n_atoms, minibatch_start, positions_tensor, box_array,
            list(particles_list), number_of_bins, cutoff, histogram=histogram)
        with np.errstate(over="ignore"):
            rdf = (rdf + mb.astype(acc_dtype)).astype(acc_dtype)  # combine_dictionaries :601-614
    return rdf


# --------------------------------------------------------------------------------------
# Same arithmetic, no index materialisation: row-chunked broadcasting.  Used for inputs
# where the literal restatement above would need GBs.  Validated against rdf_counts in
# tests/test_oracle.py.
# --------------------------------------------------------------------------------------
def rdf_counts_blocked(positions_tensor: np.ndarray, particles_list: Sequence[int], box_array,
                       cutoff: float, number_of_bins: int, skip_first: bool = True,
                       row_chunk: int = 256, histogram=tf_histogram_fixed_width,
                       return_s_near_edges: bool = False) -> np.ndarray:
    """(n_pairs, B) int64 counts; ``skip_first`` reproduces quirk Q1 (the atom at each species'
    offset never contributes as row or column); ``skip_first=False`` is the corrected mode."""
    pos = np.ascontiguousarray(positions_tensor, dtype=F32)
    box = np.asarray(box_array, dtype=F32)
    cut = F32(cutoff)
    n_atoms = pos.shape[0]
    offs = np.concatenate([[0], np.cumsum(particles_list)]).astype(np.int64)
    assert offs[-1] == n_atoms
    pairs = species_pairs(len(particles_list))
    out = np.zeros((len(pairs), number_of_bins), dtype=np.int64)
    first = 1 if skip_first else 0
    bin_range = np.array([0, cut], dtype=F32)
    for p, (a, b) in enumerate(pairs):
        r0, r1 = offs[a] + first, offs[a + 1]
        c0, c1 = offs[b] + first, offs[b + 1]
        for lo in range(r0, r1, row_chunk):
            hi = min(lo + row_chunk, r1)
            rows = np.arange(lo, hi)
            cs = max(c0, lo + 1) if a == b else c0
            if cs >= c1:
                continue
            cols = np.arange(cs, c1)
            r_ij = pos[None, cols] - pos[rows, None]  # (rows, cols, C, 3) = col - row
            r_ij = apply_minimum_image(r_ij, box)
            d = tf_norm_last_axis(r_ij)  # (rows, cols, C)
            if a == b:
                tri = cols[None, :] > rows[:, None]
                d = d[tri]
            vals = apply_system_cutoff(d, cut)
            out[p] += histogram(vals, bin_range, number_of_bins).astype(np.int64)
    return out


# --------------------------------------------------------------------------------------
# Normalisation (float64 NumPy in the reference)
# --------------------------------------------------------------------------------------
def split_array(data: np.ndarray, condition: np.ndarray) -> list:
    """utils/meta_functions.py:468-490."""
    initial_split = [data[condition], data[~condition]]
    if len(initial_split[1]) == 0:
        return [data[condition]]
    return list(initial_split)


def ideal_correction(cutoff: float, number_of_bins: int, box0: float) -> np.ndarray:
    """calculators/radial_distribution_function.py:719-826 (quirks Q2, Q8 reproduced)."""

    def _spherical_symmetry(data):
        return 4 * np.pi * (data ** 2)

    def _correction_1(data):
        return 2 * np.pi * data * (3 - 4 * data)

    def _correction_2(data):
        arctan_1 = np.arctan(np.sqrt(4 * (data ** 2) - 2))
        arctan_2 = (8 * data * np.arctan((2 * data * (4 * (data ** 2) - 3))
                                         / (np.sqrt(4 * (data ** 2) - 2) * (4 * (data ** 2) + 1))))
        return 2 * data * (3 * np.pi - 12 * arctan_1 + arctan_2)

    def _piecewise(data):
        lower_bound = box0 / 2
        middle_bound = np.sqrt(2) * box0 / 2
        split_1 = list(split_array(data, data <= lower_bound))
        if len(split_1) == 1:
            return _spherical_symmetry(split_1[0])
        split_2 = list(split_array(split_1[1], split_1[1] < middle_bound))
        if len(split_2) == 1:
            return np.concatenate((_spherical_symmetry(split_1[0]), _correction_1(split_2[0])))
        return np.concatenate((_spherical_symmetry(split_1[0]), _correction_1(split_2[0]),
                               _correction_2(split_2[1])))

    bin_width = cutoff / number_of_bins
    bin_edges = np.linspace(0.0, cutoff, number_of_bins)
    return _piecewise(np.array(bin_edges)) * bin_width


def calculate_prefactor(same_species: bool, n_species_0: int, n_species_1: int, volume: float,
                        number_of_configurations: int, ideal: np.ndarray) -> np.ndarray:
    """calculators/radial_distribution_function.py:299-345."""
    species_scale_factor = 2 if same_species else 1
    rho = n_species_1 / volume
    numerator = species_scale_factor
    denominator = number_of_configurations * rho * ideal * n_species_0
    with np.errstate(divide="ignore"):
        return numerator / denominator


def normalise(counts: np.ndarray, species: Sequence[str], particles_list: Sequence[int],
              box_array, cutoff: float, number_of_configurations: int,
              length_unit: float = 1e-10) -> Dict[str, Dict[str, list]]:
    """calculators/radial_distribution_function.py:347-393 → {"A_B": {"x": [...], "y": [...]}}."""
    number_of_bins = counts.shape[1]
    volume = float(box_array[0] * box_array[1] * box_array[2])
    ideal = ideal_correction(cutoff, number_of_bins, float(box_array[0]))
    out = {}
    for p, (a, b) in enumerate(species_pairs(len(species))):
        prefactor = calculate_prefactor(species[a] == species[b], particles_list[a],
                                        particles_list[b], volume, number_of_configurations, ideal)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            with np.errstate(invalid="ignore"):
                y = np.array(counts[p], dtype=float) * prefactor
        x = (length_unit / 1e-9) * np.linspace(0.0, cutoff, number_of_bins)
        out[f"{species[a]}_{species[b]}"] = {"x": x.tolist(), "y": y.tolist()}
    return out


# --------------------------------------------------------------------------------------
# Argument defaults / sampling
# --------------------------------------------------------------------------------------
def check_input_defaults(box_array, n_configurations_in_db: int, *, number_of_bins=None,
                         cutoff=None, start=0, stop=None, number_of_configurations=500,
                         minibatch=-1):
    """calculators/radial_distribution_function.py:215-279."""
    if stop is None:
        stop = n_configurations_in_db - 1
    if cutoff is None:
        cutoff = box_array[0] / 2 - 0.1
    if number_of_configurations == -1:
        number_of_configurations = n_configurations_in_db - 1
    if minibatch == -1:
        minibatch = number_of_configurations
    if number_of_bins is None:
        number_of_bins = int(cutoff / 0.01)
    sample_configurations = np.linspace(start, stop, number_of_configurations, dtype=int)
    return dict(number_of_bins=number_of_bins, cutoff=cutoff, start=start, stop=stop,
                number_of_configurations=number_of_configurations, minibatch=minibatch,
                sample_configurations=sample_configurations)


# --------------------------------------------------------------------------------------
# Squared-distance threshold table: an independent derivation of what the CUDA path uses,
# so tests can check the library's table entry by entry.
# --------------------------------------------------------------------------------------
def bin_of_distance(d: np.ndarray, cutoff: float, nbins: int) -> np.ndarray:
    """Bin index the TF kernel assigns to float32 distance(s) ``d`` (no cutoff test)."""
    d = np.asarray(d, dtype=F32)
    step = np.float64(F32(F32(cutoff) - F32(0))) / np.float64(nbins)
    return np.minimum(np.maximum(d, F32(0)).astype(np.float64) / step,
                      np.float64(nbins - 1)).astype(np.int32)


def s_threshold_table(cutoff: float, nbins: int) -> Tuple[np.ndarray, np.float32]:
    """edges[k] (k = 0..nbins) = smallest float32 s with bin(sqrt_rn(s)) >= k, edges[0] = 0,
    edges[nbins] = s_cut = smallest float32 s with sqrt_rn(s) >= float32(cutoff).
    Brute force by walking ulps around the analytic guess."""
    cut = F32(cutoff)
    step = np.float64(cut) / np.float64(nbins)

    def smallest_s_with_sqrt_ge(dk: np.float32) -> np.float32:
        s = F32(np.float64(dk) * np.float64(dk))
        while np.sqrt(s) >= dk and s > 0:
            s = np.nextafter(s, F32(-np.inf), dtype=F32)
        while np.sqrt(s) < dk:
            s = np.nextafter(s, F32(np.inf), dtype=F32)
        return s

    edges = np.zeros(nbins + 1, dtype=F32)
    for k in range(1, nbins):
        d = F32(k * step)
        while bin_of_distance(d, cutoff, nbins) >= k and d > 0:
            d = np.nextafter(d, F32(-np.inf), dtype=F32)
        while bin_of_distance(d, cutoff, nbins) < k:
            d = np.nextafter(d, F32(np.inf), dtype=F32)
        edges[k] = smallest_s_with_sqrt_ge(d)
    s_cut = smallest_s_with_sqrt_ge(cut)
    edges[nbins] = s_cut
    return edges, s_cut
