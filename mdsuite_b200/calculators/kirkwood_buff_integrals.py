"""
KirkwoodBuffIntegral — consumer of RDF output (reference
mdsuite/calculators/kirkwood_buff_integrals.py:116-197): running integral
G(r) = 4 pi int (g(r') - 1) r'^2 dr' of the Savitzky-Golay smoothed RDF per species pair.
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


class KirkwoodBuffIntegral(Calculator):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.x_label = r"$$ \text{r} / nm$$"
        self.y_label = r"$$\text{G}(\mathbf{r})$$"
        self.analysis_name = "Kirkwood-Buff_Integral"
        self.result_series_keys = ["r", "kb_integral"]
        self.rdf_data = None

    @call
    def __call__(self, rdf_data: Computation = None, plot=True, savgol_order: int = 2,
                 savgol_window_length: int = 17):
        self.rdf_data = rdf_data if isinstance(rdf_data, Computation) else \
            self.experiment.run.RadialDistributionFunction(plot=False)
        param = self.rdf_data.computation_parameter
        self.args = Args(savgol_order=savgol_order, savgol_window_length=savgol_window_length,
                         number_of_bins=param["number_of_bins"], cutoff=param["cutoff"],
                         number_of_configurations=param["number_of_configurations"])
        self.plot = plot

    def run_calculator(self):
        for pair_key, result in self.rdf_data.data_dict.items():
            radii, rdf = post.usable_rdf(result)
            integral = post.kirkwood_buff_running_integral(radii, rdf, self.args.savgol_order,
                                                           self.args.savgol_window_length)
            self.queue_data(data={"r": radii[1:].tolist(), "kb_integral": integral.tolist()},
                            subjects=pair_key.split("_"))
