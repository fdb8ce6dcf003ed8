"""``experiment.run`` / ``project.run`` (reference mdsuite/experiment/run.py:58-243): one property
per calculator; in this package's scope are the RDF, the calculators that consume
its output (§8f rank 3) and the angular distribution function (§8f rank 4)."""
from __future__ import annotations

from mdsuite_b200.calculators.angular_distribution_function import AngularDistributionFunction
from mdsuite_b200.calculators.coordination_number_calculation import CoordinationNumbers
from mdsuite_b200.calculators.kirkwood_buff_integrals import KirkwoodBuffIntegral
from mdsuite_b200.calculators.potential_of_mean_force import PotentialOfMeanForce
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

    @property
    def AngularDistributionFunction(self) -> AngularDistributionFunction:
        """reference run.py:164-166."""
        return AngularDistributionFunction(**self.kwargs)

    @property
    def CoordinationNumbers(self) -> CoordinationNumbers:
        """reference run.py:168-170."""
        return CoordinationNumbers(**self.kwargs)

    @property
    def KirkwoodBuffIntegral(self) -> KirkwoodBuffIntegral:
        """reference run.py:216-218."""
        return KirkwoodBuffIntegral(**self.kwargs)

    @property
    def PotentialOfMeanForce(self) -> PotentialOfMeanForce:
        """reference run.py:224-226."""
        return PotentialOfMeanForce(**self.kwargs)
