"""Generate tests/golden/rdf_golden.json — SELF-GENERATED golden vectors (the reference holds none
for this path and cannot be imported here).  Inputs are seeded; outputs come from
the literal NumPy restatement oracle/rdf_oracle.py::rdf_counts (the op-by-op one, with the
reference's index tensors and atom-minibatches).  Re-run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import rdf_oracle  # noqa: E402


def case_inputs(case):
    rng = np.random.Generator(np.random.Philox(case["seed"]))
    n = sum(case["counts"])
    box = np.array(case["box"], dtype=np.float64)
    kind = case["kind"]
    if kind == "uniform":
        pos = (rng.random((case["frames"], n, 3)) * box).astype(np.float32)
    elif kind == "unwrapped":
        pos = ((rng.random((case["frames"], n, 3)) * 4 - 1.5) * box).astype(np.float32)
    elif kind == "lattice_edges":
        # atoms on a line, spacing = bin width: distances sit on bin edges
        step = np.float32(case["cutoff"] / case["nbins"])
        pos = np.zeros((case["frames"], n, 3), dtype=np.float32)
        for f in range(case["frames"]):
            x = (np.arange(n, dtype=np.float32) * step) % np.float32(box[0])
            for _ in range(f):
                x = np.nextafter(x, np.float32(np.inf))
            pos[f, :, 0] = np.minimum(x, np.nextafter(np.float32(box[0]), np.float32(0)))
            pos[f, :, 1] = np.float32(0.125 * f)
    else:
        raise ValueError(kind)
    return pos, box


CASES = [
    dict(name="two_species_uniform", kind="uniform", seed=101, counts=[40, 24], frames=8,
         box=[11.0, 11.0, 11.0], cutoff=5.4, nbins=50, minibatch=16),
    dict(name="three_species_noncubic", kind="uniform", seed=102, counts=[17, 1, 30], frames=5,
         box=[7.3, 8.0, 9.1], cutoff=3.6, nbins=77, minibatch=-1),
    dict(name="cutoff_beyond_half_box", kind="uniform", seed=103, counts=[50], frames=4,
         box=[6.0, 6.0, 6.0], cutoff=5.0, nbins=200, minibatch=7),
    dict(name="unwrapped", kind="unwrapped", seed=104, counts=[20, 20], frames=4,
         box=[9.0, 9.0, 9.0], cutoff=4.4, nbins=44, minibatch=9),
    dict(name="distances_on_bin_edges", kind="lattice_edges", seed=105, counts=[150], frames=4,
         box=[30.0, 30.0, 30.0], cutoff=6.0, nbins=300, minibatch=50),
]


def main():
    out = []
    for case in CASES:
        pos, box = case_inputs(case)
        counts = rdf_oracle.rdf_counts(pos.transpose(1, 0, 2), case["counts"], box, case["cutoff"],
                                       case["nbins"], minibatch=case["minibatch"])
        corrected = rdf_oracle.rdf_counts_blocked(pos.transpose(1, 0, 2), case["counts"], box,
                                                  case["cutoff"], case["nbins"], skip_first=False)
        rec = dict(case)
        rec["positions_sha"] = __import__("hashlib").sha256(pos.tobytes()).hexdigest()
        rec["counts_parity"] = counts.astype(int).tolist()
        rec["counts_corrected"] = corrected.astype(int).tolist()
        out.append(rec)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rdf_golden.json")
    with open(path, "w") as fh:
        json.dump({"generator": "tests/golden/make_golden.py", "self_generated": True,
                   "cases": out}, fh)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
