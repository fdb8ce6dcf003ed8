"""Calculator-level throughput from the trajectory store (Project -> Experiment -> FrameBatcher ->
C ABI), i.e. what a user's `experiment.run.RadialDistributionFunction(...)` call costs end to end.
Usage: python scripts/calculator_bench.py [cfg2|cfg3] [n_frames]"""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mdsuite_b200 as mds  # noqa: E402
from mdsuite_b200 import synthetic  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 256
if which == "cfg3":
    sysm = synthetic.lj_fluid(47, 8, seed=3, n_atoms=100_000)
    kw = dict(cutoff=3.0, number_of_bins=300)
else:
    sysm = synthetic.nacl_melt(12, 8, seed=2)
    kw = dict()
pos = np.tile(sysm.positions, ((n_frames + 7) // 8, 1, 1))[:n_frames]
with tempfile.TemporaryDirectory() as tmp:
    project = mds.Project("bench", storage_path=tmp)
    exp = project.add_experiment("e", units="real")
    t0 = time.perf_counter()
    exp.add_positions(pos, sysm.species, sysm.box)
    t_ingest = time.perf_counter() - t0
    out = {"workload": which, "atoms": sysm.n_atoms, "frames": n_frames,
           "store_mb": pos.nbytes / 1e6, "ingest_s": t_ingest}
    for rep in range(6):
        t0 = time.perf_counter()
        calc = exp.run.RadialDistributionFunction
        comp = calc(number_of_configurations=n_frames - rep, plot=False, save=False, **kw)
        dt = time.perf_counter() - t0
        out[f"run_{rep}_s"] = dt
    best = min(out[f"run_{rep}_s"] for rep in range(1, 6))
    out["best_s"] = best
    out["frames_per_s"] = (n_frames - 3) / best
    out["store_gb_per_s"] = (n_frames - 3) * sysm.n_atoms * 12 / best / 1e9
print(json.dumps(out))
