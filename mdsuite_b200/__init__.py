"""
mdsuite_b200 — B200-native radial distribution function for MDSuite-style trajectory analysis.

Scope: the minimum-image pair-distance histogram behind
``Experiment.run.RadialDistributionFunction(...)`` of zincware/MDSuite, rebuilt as hand-written
sm_100a CUDA behind a C ABI (include/mds_rdf.h), plus the host-side mirror of the reference's
calculator interface (Project -> Experiment -> run.RadialDistributionFunction -> Computation).
Nothing here falls back to a CPU implementation of the hot path.
"""
import logging
import sys

__version__ = "0.1.0"

# stdout handler at INFO like the reference (mdsuite/__init__.py:53-63)
_logger = logging.getLogger(__name__)
if not _logger.handlers:
    _h = logging.StreamHandler(sys.stdout)
    _h.setLevel(logging.INFO)
    _h.setFormatter(logging.Formatter("%(asctime)s %(levelname)s: %(message)s"))
    _logger.addHandler(_h)
    _logger.setLevel(logging.INFO)

from mdsuite_b200.utils.config import config  # noqa: E402,F401
from mdsuite_b200.project.project import Project  # noqa: E402,F401
from mdsuite_b200.experiment.experiment import Experiment  # noqa: E402,F401
