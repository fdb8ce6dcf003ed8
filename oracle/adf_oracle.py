"""
TEST INFRASTRUCTURE ONLY — CPU restatement (NumPy) of the reference's angular distribution function,
the checker for the CUDA path behind include/mds_adf.h.  Never imported by the product.

    PARITY PINNED FOR CONTROL FLOW, UNPINNED FOR LEAF ARITHMETIC: tests/golden/adf_golden.json is
    made by executing the reference's own ``_build_histograms`` / ``_compute_adfs`` /
    ``get_neighbour_list`` / ``get_triplets`` / ``get_angles`` over a NumPy stand-in for the
    TensorFlow leaf ops (tests/golden/make_adf_golden.py); this restatement is held to it in
    tests/test_adf_oracle.py.  TensorFlow itself is not installed here, so the float32 results of
    its norm / division / acos / pow kernels are restated by the contracts below, not observed.

What the reference computes (calculators/angular_distribution_function.py:143-609,
utils/neighbour_list.py:36-177, utils/linalg.py:31-81), per batch of configurations:

  r[i][j]   = minimum image of pos[i] - pos[j]           float32, r -= rint(r / L) * L   (:63-78)
  n[i][j]   = sqrt((x*x + y*y) + z*z)                     float32 (tf.norm)
  neighbour = float16(n) < float16(r_cut), with float16(n) == 0 read as "not a neighbour"
              (get_triplets :147-149: the norm is CAST TO float16 before the comparison)
  triples   = every ORDERED (i, j, k), j != k, j and k neighbours of i                  (:150-170)
  kept for the species combination (A, B, C), A <= B <= C in storage order
  (itertools.combinations_with_replacement), iff species(i) = A, species(j) = B, species(k) = C
  (_compute_angles :341-372) — so B-A-B angles around the LATER species are never counted
  cos       = clip((ux1*ux2 + uy1*uy2) + uz1*uz2, -1, 1),  u = r / n per component  (linalg :31-50)
  theta     = acos(cos)                                   float32
  weight    = 1 / (n[i][j] * n[i][k]) ** norm_power       float32
  histogram = np.histogram(theta, bins, range=[0, 3.15], weights=weight, density=True) as float32,
              summed over the batches                                                   (:374-405)
  x axis    = linspace(0, 3.15 * 180 / 3.14159, bins), max_peak = x[argmax(histogram)] (:407-441)

Contracts for the leaf arithmetic (each op rounded to float32 separately, no FMA):
  theta  = float32(acos_float64(cos))       (the correctly rounded float32 arccosine)
  weight = float32(1) / float32_pow(p, norm_power), 1 for norm_power = 0
np.histogram is NumPy's own code here as in the reference (bin of a float32 angle found in float64
against linspace(0, 3.15, bins + 1); weights summed per 65 536-element block in float64, blocks
accumulated in float32).
"""
from __future__ import annotations

import itertools
from typing import Dict, List, Sequence, Tuple

import numpy as np

BIN_RANGE = (0.0, 3.15)  # "from 0 to a chemists pi" (:215)


def min_image_vectors(pos: np.ndarray, box: np.ndarray) -> np.ndarray:
    """r[i][j] = minimg(pos[i] - pos[j]), float32 (N, N, 3)."""
    pos = np.asarray(pos, dtype=np.float32)
    box = np.asarray(box, dtype=np.float32)
    with np.errstate(invalid="ignore"):
        r = pos[:, None, :] - pos[None, :, :]
        r = r - np.rint(r / box) * box
    return r


def norms(r: np.ndarray) -> np.ndarray:
    sq = r * r
    return np.sqrt((sq[..., 0] + sq[..., 1]) + sq[..., 2])


def neighbour_mask(n: np.ndarray, cutoff: float) -> np.ndarray:
    with np.errstate(over="ignore"):
        n16 = n.astype(np.float16)
    cut16 = np.float16(cutoff)
    return (n16 != 0) & (n16 < cut16)


def species_combinations(n_species: int) -> List[Tuple[int, int, int]]:
    return list(itertools.combinations_with_replacement(range(n_species), 3))


def frame_triples(pos: np.ndarray, counts: Sequence[int], box, cutoff: float, norm_power: int
                  ) -> Dict[Tuple[int, int, int], Tuple[np.ndarray, np.ndarray]]:
    """{(A, B, C): (theta float32, weight float32)} for one configuration, triples in (i, j, k)
    lexicographic order."""
    r = min_image_vectors(pos, box)
    n = norms(r)
    mask = neighbour_mask(n, cutoff)
    species = np.repeat(np.arange(len(counts)), counts)
    out = {c: ([], []) for c in species_combinations(len(counts))}
    one = np.float32(1.0)
    for i in range(len(species)):
        nb = np.nonzero(mask[i])[0]
        if len(nb) < 2:
            continue
        with np.errstate(invalid="ignore", divide="ignore"):
            u = r[i, nb] / n[i, nb][:, None]
        prod = u[:, None, :] * u[None, :, :]
        cos = np.clip((prod[..., 0] + prod[..., 1]) + prod[..., 2], np.float32(-1), np.float32(1))
        theta = np.arccos(cos.astype(np.float64)).astype(np.float32)
        p = n[i, nb][:, None] * n[i, nb][None, :]
        if norm_power == 0:
            w = np.ones_like(p)
        else:
            with np.errstate(over="ignore", divide="ignore"):
                w = one / np.power(p, np.float32(norm_power))
        sp_nb = species[nb]
        for (a, b, c) in out:
            if species[i] != a:
                continue
            sel = (sp_nb[:, None] == b) & (sp_nb[None, :] == c)
            np.fill_diagonal(sel, False)
            if sel.any():
                out[(a, b, c)][0].append(theta[sel])
                out[(a, b, c)][1].append(w[sel])
    res = {}
    for key, (t, w) in out.items():
        res[key] = (np.concatenate(t) if t else np.zeros(0, np.float32),
                    np.concatenate(w) if w else np.zeros(0, np.float32))
    return res


def raw_histograms(pos: np.ndarray, counts: Sequence[int], box, cutoff: float, nbins: int,
                   norm_power: int) -> np.ndarray:
    """Sum of the weights per bin, (n_combinations, nbins) float64, for one configuration — what
    the C ABI returns per configuration (integers when norm_power = 0)."""
    triples = frame_triples(pos, counts, box, cutoff, norm_power)
    out = np.zeros((len(triples), nbins), dtype=np.float64)
    for c, (theta, w) in enumerate(triples.values()):
        if len(theta):
            out[c] = np.histogram(theta, bins=nbins, range=BIN_RANGE,
                                  weights=w.astype(np.float64))[0]
    return out


def adf(positions: np.ndarray, counts: Sequence[int], names: Sequence[str], box, cutoff: float,
        nbins: int, norm_power: int, batches: Sequence[Sequence[int]]) -> Dict[str, dict]:
    """The reference's result per species combination for configurations split into ``batches``
    (lists of frame indices): {"A-B-C": {"angle": [...], "adf": float32 [...], "max_peak": x}}."""
    combos = species_combinations(len(counts))
    total = {c: None for c in combos}
    for frames in batches:
        per_frame = [frame_triples(positions[f], counts, box, cutoff, norm_power) for f in frames]
        for c in combos:
            theta = np.concatenate([t[c][0] for t in per_frame])
            w = np.concatenate([t[c][1] for t in per_frame])
            with np.errstate(invalid="ignore", divide="ignore"):
                hist = np.histogram(theta, bins=nbins, range=BIN_RANGE, weights=w, density=True)[0]
            hist = hist.astype(np.float32)
            total[c] = hist if total[c] is None else total[c] + hist
    axis = np.linspace(BIN_RANGE[0] * (180 / 3.14159), BIN_RANGE[1] * (180 / 3.14159), nbins)
    out = {}
    for c in combos:
        name = "-".join(names[s] for s in c)
        out[name] = {"angle": axis, "adf": total[c], "max_peak": float(axis[int(np.argmax(total[c]))])}
    return out
