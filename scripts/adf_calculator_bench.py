"""The user-level ADF call on one GPU: Project -> Experiment -> run.AngularDistributionFunction over a
cfg2-shaped NaCl melt in the frame store, timed end to end (store read, H2D, kernels, D2H,
normalisation, SQL).  python scripts/adf_calculator_bench.py [--cells 12] [--frames 256]"""
import argparse
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=12)
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--batches", type=int, default=4)
    ap.add_argument("--cutoff", type=float, default=6.0)
    args = ap.parse_args()
    import mdsuite_b200 as mds
    from mdsuite_b200 import synthetic

    sysm = synthetic.nacl_melt(args.cells, 8, seed=2)
    import numpy as np
    pos = np.tile(sysm.positions, (args.frames // 8, 1, 1))
    with tempfile.TemporaryDirectory() as tmp:
        project = mds.Project("adf_bench", storage_path=tmp)
        exp = project.add_experiment("NaCl", timestep=0.002, temperature=1400.0, units="metal")
        exp.add_positions(pos, sysm.species, sysm.box)
        out = []
        for rep in range(3):
            t0 = time.perf_counter()
            comp = exp.run.AngularDistributionFunction(
                number_of_configurations=args.frames - rep, cutoff=args.cutoff, start=0,
                number_of_bins=500, batches=args.batches, plot=False)
            out.append(time.perf_counter() - t0)
        print(json.dumps({"workload": f"nacl_{sysm.n_atoms}_atoms_{args.frames}_frames_rc{args.cutoff}",
                          "batches": args.batches, "seconds_per_call": out,
                          "ms_per_configuration": 1e3 * min(out) / args.frames,
                          "keys": comp.keys()}))


if __name__ == "__main__":
    main()
