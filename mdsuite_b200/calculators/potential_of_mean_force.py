"""
PotentialOfMeanForce — consumer of RDF output (reference
mdsuite/calculators/potential_of_mean_force.py:131-347): w(r) = -k_B T ln g(r) per species pair and
its value at the minimum between the first peaks.  The math lives in ``rdf_postprocessing``.
"""
from __future__ import annotations

from dataclasses import dataclass

from mdsuite_b200.calculators import rdf_postprocessing as post
from mdsuite_b200.calculators.calculator import Calculator, call
from mdsuite_b200.database.project_database import Computation


@dataclass
class Args:
    savgol_order: int
    savgol_window_length: int
    number_of_bins: int
    number_of_configurations: int
    cutoff: float
    number_of_shells: int


class PotentialOfMeanForce(Calculator):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.x_label = r"$$r / nm$$"
        self.y_label = r"$$w^{(2)}(r)$$"
        self.analysis_name = "Potential_of_Mean_Force"
        self.result_keys = []
        self.result_series_keys = ["r", "pomf"]
        self.rdf_data = None

    @call
    def __call__(self, rdf_data: Computation = None, plot=True, savgol_order: int = 2,
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
            self.result_keys += [f"POMF_{shell}", f"POMF_{shell}_error"]
        self.plot = plot

    def run_calculator(self):
        for pair_key, result in self.rdf_data.data_dict.items():
            radii, rdf = post.usable_rdf(result)
            pomf = post.potential_of_mean_force(rdf, self.experiment.temperature)
            peaks = post.smoothed_peaks(pomf, self.args.savgol_order, self.args.savgol_window_length)
            if len(peaks) < self.args.number_of_shells + 1:
                raise ValueError(
                    "Not enough peaks were detecting in the RDF to perform the desired analysis. "
                    "Try reducing the number of desired peaks or improving the quality of the RDF "
                    "provided.")
            shells = post.shell_brackets(radii, pomf, peaks, self.args.number_of_shells)
            # the series keeps the reference's lengths: r on radii[1:], pomf on radii (:333-336)
            data = {"r": radii[1:].tolist(), "pomf": pomf.tolist()}
            for shell, bracket in shells.items():
                mean, err = post.bracket_mean_and_error(pomf, bracket)
                data[f"POMF_{shell + 1}"] = mean
                data[f"POMF_{shell + 1}_error"] = err
            self.queue_data(data=data, subjects=pair_key.split("_"))
