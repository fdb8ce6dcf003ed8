"""
mdsuite_b200 — B200-native radial distribution function for MDSuite-style trajectory analysis.

Scope: the minimum-image pair-distance histogram behind
``Experiment.run.RadialDistributionFunction(...)`` of zincware/MDSuite, rebuilt as hand-written
sm_100a CUDA behind a C ABI (include/mds_rdf.h), plus the host-side mirror of the reference's
calculator interface.  Nothing here falls back to a CPU implementation.
"""
__version__ = "0.1.0"
