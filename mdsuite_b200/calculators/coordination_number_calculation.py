"""
CoordinationNumbers — consumer of RDF output (reference
mdsuite/calculators/coordination_number_calculation.py:157-359): running coordination number
n(r) = 4 pi rho int r^2 g(r) dr per species pair and its value at the minimum after each of the
first ``number_of_shells`` peaks.  The math lives in ``rdf_postprocessing``.
"""
from __future__ import annotations

import logging
from dataclasses import dataclass

import numpy as np

from mdsuite_b200.calculators import rdf_postprocessing as post
from mdsuite_b200.calculators.calculator import Calculator, call
from mdsuite_b200.database.project_database import Computation

log = logging.getLogger(__name__)


class CannotPerformThisAnalysis(Exception):
    """mdsuite/utils/exceptions.py — raised when the RDF has too few peaks."""


@dataclass
class Args:
    savgol_order: int
    savgol_window_length: int
    number_of_bins: int
    number_of_configurations: int
    cutoff: float
    number_of_shells: int


class CoordinationNumbers(Calculator):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.x_label = r"$$\text{r} /  nm$$"
        self.y_label = "CN"
        self.analysis_name = "Coordination_Numbers"
        self.result_keys = []
        self.result_series_keys = ["r", "cn"]
        self.rdf_data = None
        self.volume = None

    @call
    def __call__(self, rdf_data: Computation = None, plot: bool = True, savgol_order: int = 2,
                 savgol_window_length: int = 17, number_of_shells: int = 1):
        self.rdf_data = rdf_data if isinstance(rdf_data, Computation) else \
            self.experiment.run.RadialDistributionFunction(plot=False)
        param = self.rdf_data.computation_parameter
        self.args = Args(savgol_order=savgol_order, savgol_window_length=savgol_window_length,
                         number_of_bins=param["number_of_bins"], cutoff=param["cutoff"],
                         number_of_configurations=param["number_of_configurations"],
                         number_of_shells=number_of_shells)
        self.result_keys = []
        for shell in range(1, number_of_shells + 1):
            self.result_keys += [f"CN_{shell}", f"CN_{shell}_error"]
        # box volume in nm^3 (:208-218)
        self.volume = self.experiment.volume * self.experiment.units.volume / 1e-9 ** 3
        self.plot = plot

    def _density(self, pair_key: str) -> float:
        """Number density of the FIRST species of the pair in nm^-3 (:220-225, :341)."""
        first = pair_key.split("_")[0]
        return self.experiment.species[first].n_particles / self.volume

    def run_calculator(self):
        for pair_key, result in self.rdf_data.data_dict.items():
            radii, rdf = post.usable_rdf(result)
            running = post.running_coordination(radii, rdf, self._density(pair_key))
            peaks = post.smoothed_peaks(rdf, self.args.savgol_order, self.args.savgol_window_length,
                                        height=1.0)
            if len(peaks) < self.args.number_of_shells + 1:
                raise CannotPerformThisAnalysis(
                    "Not enough peaks were detecting in the RDF to perform the desired analysis. "
                    "Try reducing the number of desired peaks or improving the quality of the RDF "
                    "provided.")
            shells = post.shell_brackets(radii, rdf, peaks, self.args.number_of_shells)
            data = {"r": radii[1:].tolist(), "cn": running.tolist()}
            for shell, bracket in shells.items():
                mean, err = post.bracket_mean_and_error(running, bracket)
                data[f"CN_{shell + 1}"] = mean
                data[f"CN_{shell + 1}_error"] = err
            self.queue_data(data=data, subjects=pair_key.split("_"))
