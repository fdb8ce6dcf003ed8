"""ADF path on one GPU: triples/s and neighbour pair tests/s on a cfg2-shaped NaCl melt, with the
NumPy oracle timed on a small sample beside it.  python scripts/adf_bench.py [--cells 12]
[--frames 64] [--cutoff 6.0] [--bins 500]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=12)
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--cutoff", type=float, default=6.0)
    ap.add_argument("--bins", type=int, default=500)
    ap.add_argument("--power", type=int, default=4)
    ap.add_argument("--repeats", type=int, default=3)
    ap.add_argument("--cpu-atoms-cells", type=int, default=4)
    ap.add_argument("--lj", action="store_true",
                    help="cfg3-shaped 100 000-atom LJ fluid (one species) instead of the NaCl melt")
    args = ap.parse_args()
    import torch
    from mdsuite_b200 import synthetic
    from mdsuite_b200._cuda import adf as cuda_adf
    from oracle import adf_oracle

    if args.lj:
        sysm = synthetic.lj_fluid(47, args.frames, seed=3, n_atoms=100_000)
    else:
        sysm = synthetic.nacl_melt(args.cells, args.frames, seed=2)
    pos_d = torch.from_numpy(sysm.positions).cuda()
    with cuda_adf.ADFHandle(sysm.counts, args.bins, args.cutoff, sysm.box, norm_power=args.power,
                            max_frames_in_flight=args.frames) as h:
        out = h.compute_device(pos_d)  # warm-up
        t0 = h.stats().triples
        times = []
        for _ in range(args.repeats):
            h.compute_device(pos_d, out=out)
            times.append(h.stats().kernel_ms)
        st = h.stats()
        host_ms = None
        for _ in range(3):  # the first call also allocates the staging buffers
            host_t = time.perf_counter()
            h.compute(sysm.positions)
            dt = (time.perf_counter() - host_t) * 1e3
            host_ms = dt if host_ms is None else min(host_ms, dt)
    triples_per_call = t0
    ms = float(np.median(times))
    small = synthetic.nacl_melt(args.cpu_atoms_cells, 1, seed=3)
    c0 = time.perf_counter()
    raw = adf_oracle.raw_histograms(small.positions[0], small.counts, small.box, args.cutoff,
                                    args.bins, args.power)
    cpu_s = time.perf_counter() - c0
    cpu_triples = adf_oracle.raw_histograms(small.positions[0], small.counts, small.box, args.cutoff,
                                            args.bins, 0).sum()
    print(json.dumps({
        "workload": f"{sysm.name}_{sysm.n_atoms}_atoms_{args.frames}_frames_rc{args.cutoff}_B{args.bins}",
        "triples_per_call": triples_per_call, "kernel_ms": ms,
        "triples_per_s": triples_per_call / (ms * 1e-3),
        "allatoms_equivalent_pair_tests_per_s": args.frames * sysm.n_atoms ** 2 / (ms * 1e-3),
        "host_call_ms": host_ms, "max_neighbours": st.max_neighbours, "cells": list(st.cells),
        "pair_tests_per_call": st.pair_tests / (args.repeats + 4),
        "cpu_oracle": {"atoms": small.n_atoms, "seconds": cpu_s,
                       "triples_per_s": float(cpu_triples) / cpu_s, "kind": "port (NumPy)"},
        "checksum": float(raw.sum())}))


if __name__ == "__main__":
    main()
