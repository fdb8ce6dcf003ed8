"""LAMMPS dump -> project -> angular distribution function on one GPU.

    python examples/adf_nacl.py [--cells 6] [--frames 20]

Writes a synthetic NaCl melt as a LAMMPS dump, loads it into a project and runs
``experiment.run.AngularDistributionFunction`` (the reference's call, same arguments); prints the
position of the main peak of every species combination.
"""
import argparse
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=6)
    ap.add_argument("--frames", type=int, default=20)
    args = ap.parse_args()
    import mdsuite_b200 as mds
    from mdsuite_b200 import synthetic

    sysm = synthetic.nacl_melt(args.cells, args.frames, seed=1)
    with tempfile.TemporaryDirectory() as tmp:
        dump = os.path.join(tmp, "nacl.lammpstraj")
        synthetic.write_lammps_dump(sysm, dump)
        project = mds.Project("adf_example", storage_path=tmp)
        exp = project.add_experiment("NaCl", timestep=0.002, temperature=1400.0, units="metal",
                                     simulation_data=dump)
        adf = exp.run.AngularDistributionFunction(number_of_configurations=args.frames - 1,
                                                  cutoff=4.5, number_of_bins=180, plot=False)
        for key in adf.keys():
            print(f"{key:10s} main peak at {adf[key]['max_peak']:6.1f} degrees")


if __name__ == "__main__":
    main()
