"""``experiment.run`` / ``project.run`` (reference mdsuite/experiment/run.py:58-243): one property
per calculator; only the RDF is in this package's scope."""
from __future__ import annotations

from mdsuite_b200.calculators.radial_distribution_function import RadialDistributionFunction


class RunComputation:
    def __init__(self, experiment=None, experiments=None):
        self.experiment = experiment
        self.experiments = experiments
        self.kwargs = {"experiment": experiment, "experiments": experiments}

    @property
    def RadialDistributionFunction(self) -> RadialDistributionFunction:
        """reference run.py:228-230."""
        return RadialDistributionFunction(**self.kwargs)
