"""
Pure functions behind the three calculators that consume RDF output:
coordination numbers, potential of mean force, Kirkwood-Buff integrals.  Cheap NumPy / SciPy
post-processing on the ``x`` / ``y`` lists a RadialDistributionFunction computation stores.

Conventions shared with the reference (all three calculators do this): index 0 of the stored RDF is
dropped (``y[0]`` is nan/inf, quirk Q2), so ``radii = x[1:]`` and ``rdf = y[1:]``; series results
are reported on ``radii[1:]``.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
from scipy.integrate import cumulative_trapezoid
from scipy.signal import find_peaks, savgol_filter

GOLDEN_RATIO = 1.618033988749895  # mdsuite/utils/units.py:42
BOLTZMANN_SI = 1.380649e-23  # J / K, mdsuite/utils/units.py:32
POMF_TO_EV = 6.242e8  # the factor the reference applies (potential_of_mean_force.py:200) as is


def usable_rdf(result: dict) -> Tuple[np.ndarray, np.ndarray]:
    """(radii, rdf) with the singular first point removed."""
    return (np.asarray(result["x"], dtype=float)[1:], np.asarray(result["y"], dtype=float)[1:])


def running_coordination(radii: np.ndarray, rdf: np.ndarray, density: float) -> np.ndarray:
    """4 pi rho * cumulative trapezoid of r^2 g(r) over radii[1:]
    (coordination_number_calculation.py:59-82)."""
    r, g = radii[1:], rdf[1:]
    return 4.0 * np.pi * density * cumulative_trapezoid(y=r ** 2 * g, x=r)


def kirkwood_buff_running_integral(radii: np.ndarray, rdf: np.ndarray, order: int,
                                   window_length: int) -> np.ndarray:
    """4 pi * cumulative trapezoid of (smoothed g - 1) r^2 over radii[1:]
    (kirkwood_buff_integrals.py:160-185)."""
    smooth = savgol_filter(rdf, window_length, order)
    r = radii[1:]
    return 4.0 * np.pi * cumulative_trapezoid(y=(smooth[1:] - 1.0) * r ** 2, x=r)


def potential_of_mean_force(rdf: np.ndarray, temperature: float) -> np.ndarray:
    """-k_B T ln g(r), scaled as the reference does (potential_of_mean_force.py:180-200)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return -1.0 * BOLTZMANN_SI * temperature * np.log(rdf) * POMF_TO_EV


def smoothed_peaks(values: np.ndarray, order: int, window_length: int, height=None) -> np.ndarray:
    """Indices of the local maxima of the Savitzky-Golay smoothed curve."""
    smooth = savgol_filter(values, window_length, order)
    return find_peaks(smooth, height=height)[0]


def _nearest_grid_value(grid: np.ndarray, value: float) -> float:
    """First grid value at minimal distance (mdsuite/utils/meta_functions.py:358-373)."""
    return float(grid[int(np.argmin(np.abs(grid - value)))])


def grid_golden_section(grid: np.ndarray, values: np.ndarray, bound_1: float, bound_2: float,
                        tol: float = 1e-5) -> Tuple[float, float]:
    """Golden-section bracketing of a minimum of ``values`` sampled on ``grid``
    (mdsuite/utils/meta_functions.py:376-437, written as a loop).

    Probe points are snapped to the nearest grid value; the interval is replaced by [a, d] when
    f(c) < f(d), else by [c, b], and its nominal width shrinks by 1/phi each round until it is
    below ``tol``.  Returns the final (a, b), which are grid values once a round has run.
    """
    inv_phi, inv_phi2 = 1.0 / GOLDEN_RATIO, 1.0 / GOLDEN_RATIO ** 2
    a, b = (min(bound_1, bound_2), max(bound_1, bound_2))
    width = b - a
    c = d = fc = fd = None

    def value_at(x):
        return values[np.where(grid == x)]

    while width > tol:
        a, b = (min(a, b), max(a, b))  # the reference re-orders the bounds on every round
        if c is None:
            c = _nearest_grid_value(grid, a + inv_phi2 * width)
            fc = value_at(c)
        if d is None:
            d = _nearest_grid_value(grid, a + inv_phi * width)
            fd = value_at(d)
        if fc < fd:
            b, d, fd, c, fc = d, c, fc, None, None
        else:
            a, c, fc, d, fd = c, d, fd, None, None
        width *= inv_phi
    return min(a, b), max(a, b)


def shell_brackets(grid: np.ndarray, values: np.ndarray, peaks: np.ndarray, n_shells: int) -> dict:
    """{shell: [i_lo, i_hi]} grid indices bracketing the minimum between consecutive peaks
    (coordination_number_calculation.py:266-292, potential_of_mean_force.py:262-293)."""
    out = {}
    for k in range(n_shells):
        lo, hi = grid_golden_section(grid, values, grid[peaks[k + 1]], grid[peaks[k]])
        out[k] = np.array([np.where(grid == lo)[0][0], np.where(grid == hi)[0][0]], dtype=int)
    return out


def bracket_mean_and_error(values: np.ndarray, bracket) -> Tuple[float, float]:
    """Mean of the two bracket values and their standard deviation / sqrt(2)."""
    pair = [values[bracket[0]], values[bracket[1]]]
    return float(np.mean(pair)), float(np.std(pair) / np.sqrt(2))
