"""End-to-end example on one B200: write a small synthetic NaCl melt as a LAMMPS dump, ingest it
through the native reader, run the RDF calculator (same call as in MDSuite), then the calculators
that consume its output.

    python examples/rdf_nacl.py [workdir]
"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mdsuite_b200 as mds  # noqa: E402
from mdsuite_b200 import synthetic  # noqa: E402

workdir = sys.argv[1] if len(sys.argv) > 1 else tempfile.mkdtemp(prefix="mds_b200_")
sysm = synthetic.nacl_melt(8, 50, seed=1)  # 4,096 ions, 50 configurations
dump = os.path.join(workdir, "nacl.lammpstraj")
synthetic.write_lammps_dump(sysm, dump)

project = mds.Project("example", storage_path=workdir)
project.add_experiment("NaCl_1400K", timestep=0.002, temperature=1400.0, units="metal",
                       simulation_data=dump)
exp = project.experiments.NaCl_1400K

rdf = exp.run.RadialDistributionFunction(number_of_configurations=50, number_of_bins=500,
                                         cutoff=12.0, plot=True)
for key in rdf.keys():
    y = np.array(rdf[key]["y"])[1:]
    x = np.array(rdf[key]["x"])[1:]
    print(f"g_{key}(r): first peak {y.max():.2f} at r = {x[np.argmax(y)]:.3f} nm")

cn = exp.run.CoordinationNumbers(rdf_data=rdf, plot=False, number_of_shells=1)
print("Na-Cl first-shell coordination number:", round(cn["Na_Cl"]["CN_1"], 2))
kb = exp.run.KirkwoodBuffIntegral(rdf_data=rdf, plot=False)
print("Na-Cl Kirkwood-Buff integral at the cutoff:", round(kb["Na_Cl"]["kb_integral"][-1], 4), "nm^3")
print("results and figures under", os.path.join(workdir, "example"))
