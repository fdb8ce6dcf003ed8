"""The five BASELINE.json configurations through the C ABI on one GPU: kernel-only
(frames resident in HBM) and end-to-end (pinned host frames) rates, median and min of >= 5
repetitions, roofline fraction of the pair kernel, and the correctness gates that go with every
throughput number.  Frame counts are reduced where the full trajectory would not fit the time
budget of a run; the per-frame cost does not depend on the count (frames are independent).

    python scripts/run_configs.py [cfg1 cfg2 cfg3 cfg4 cfg5] > profiles/<name>.json
"""
import json
import os
import statistics
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdsuite_b200 import _cuda, synthetic  # noqa: E402
from oracle import c_oracle  # noqa: E402  (the checker, on a bounded sample)

PEAK = 148 * 128 * 1.965e9 / 22
REPS = 5


def measure(name, sysm, cutoff, nbins, frames, oracle_frames, path=None, parity=True):
    base = sysm.positions
    reps = (frames + base.shape[0] - 1) // base.shape[0]
    host = torch.from_numpy(np.tile(base, (reps, 1, 1))[:frames]).pin_memory()
    dev = host.cuda()
    out = {"config": name, "atoms": sysm.n_atoms, "frames": frames, "cutoff": cutoff, "nbins": nbins,
           "species_pairs": len(sysm.counts) * (len(sysm.counts) + 1) // 2}
    with _cuda.RDFHandle(sysm.counts, nbins, cutoff, sysm.box, parity=parity, path=path) as h:
        kern, step, e2e = [], [], []
        for _ in range(REPS + 1):
            h.reset()
            h.submit_device(dev)
            st = h.stats()
            kern.append(st.pair_kernel_ms)
            step.append(st.submit_ms)
        for _ in range(REPS + 1):
            h.reset()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            h.submit(host, pinned=True)
            got = h.counts()
            e2e.append((time.perf_counter() - t0) * 1e3)
        kern, step, e2e = kern[1:], step[1:], e2e[1:]  # first repetition is the warm-up
        equiv = st.pairs_allpairs_equiv
        out.update({
            "path": st.path_name, "cells": list(st.cells), "cell_stencil": list(st.cell_stencil),
            "pairs_distance_tested_per_frame": st.pairs_tested / frames,
            "pairs_tested_per_frame": st.pairs_evaluated / frames,
            "allpairs_pairs_per_frame": equiv / frames,
            "binned_fraction_of_tested": st.pairs_binned / max(st.pairs_evaluated, 1),
            "pair_kernel_ms": {"median": statistics.median(kern), "min": min(kern)},
            "resident_step_ms": {"median": statistics.median(step), "min": min(step)},
            "e2e_ms": {"median": statistics.median(e2e), "min": min(e2e)},
            "pairs_tested_per_s_kernel": st.pairs_evaluated / (statistics.median(kern) * 1e-3),
            "roofline_frac_kernel": st.pairs_evaluated / (statistics.median(kern) * 1e-3) / PEAK,
            "allpairs_equivalent_per_s_resident": equiv / (statistics.median(step) * 1e-3),
            "allpairs_equivalent_per_s_e2e": equiv / (statistics.median(e2e) * 1e-3),
            "us_per_frame_resident": statistics.median(step) * 1e3 / frames,
            "bins_over_int32": int((got >= np.uint64(2 ** 31)).sum()),
        })
        if name.startswith("cfg5") and not parity:
            # analytic check: uniform random points, every pair counted -> g(r) = 1; compare the
            # counts with the exact shell volumes (minimum-image sphere inside the box: r <= L/2)
            n, L = sysm.n_atoms, float(sysm.box[0])
            k = np.arange(nbins)
            dr = cutoff / nbins
            shell = 4 * np.pi / 3 * (((k + 1) * dr) ** 3 - (k * dr) ** 3)
            expect = frames * n * (n - 1) / 2 * shell / L ** 3
            ok = (k + 1) * dr <= L / 2
            # the timed trajectory repeats the distinct frames `reps` times: the fluctuation of a
            # count is that of the distinct frames, scaled by reps
            z = (got[0][ok].astype(np.float64) - expect[ok]) / np.sqrt(expect[ok] * reps)
            big = expect[ok] > 50  # Gaussian regime
            out["ideal_gas_gate"] = {"bins_checked": int(big.sum()), "distinct_frames": int(base.shape[0]),
                                     "max_abs_z": float(np.abs(z[big]).max()) if big.any() else None,
                                     "rms_z": float(np.sqrt(np.mean(z[big] ** 2))) if big.any() else None,
                                     "g_mean": float(np.mean(got[0][ok][big] / expect[ok][big]))
                                     if big.any() else None}
        # gates: bit-exact vs the oracle on a bounded sample of the same frames
        nf = min(oracle_frames, base.shape[0])
        if nf > 0:
            h.reset()
            h.submit(np.ascontiguousarray(base[:nf]))
            sample = h.counts().astype(np.int64)
            t0 = time.perf_counter()
            want, ev = c_oracle.rdf_counts_c(base[:nf], sysm.counts, sysm.box, cutoff, nbins,
                                             skip_first=parity, layout=c_oracle.FRAME_MAJOR)
            dt = time.perf_counter() - t0
            out["oracle_gate"] = {"frames": nf, "bins_differing": int((sample != want).sum()),
                                  "pairs_identity_ok": bool(ev == sysm.pairs_per_frame(parity) * nf),
                                  "cpu_c_port_pairs_per_s": ev / dt, "cpu_threads": c_oracle.max_threads()}
    return out


def main():
    which = sys.argv[1:] or ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]
    results = []
    if "cfg1" in which:
        s = synthetic.nacl_melt(5, 100, seed=1)
        results.append(measure("cfg1 NaCl 1,000 atoms, 100 cfgs, B=500, rc=10", s, 10.0, 500, 100, 100))
    if "cfg2" in which:
        s = synthetic.nacl_melt(12, 16, seed=2)
        cutoff = float(s.box[0] / 2 - 0.1)
        results.append(measure("cfg2 NaCl 13,824 atoms, 1,000 cfgs, default cutoff/bins", s, cutoff,
                               int(cutoff / 0.01), 1000, 16))
        results.append(measure("cfg2 NaCl 13,824 atoms, 1,000 cfgs, B=500, rc=10", s, 10.0, 500, 1000, 16))
    if "cfg3" in which:
        s = synthetic.lj_fluid(47, 8, seed=3, n_atoms=100_000)
        results.append(measure("cfg3 LJ 100,000 atoms, 1,000 of 10,000 cfgs, B=300, rc=3",
                               s, 3.0, 300, 1000, 1))
    if "cfg4" in which:
        s = synthetic.spc_water(70, 2, seed=4, n_molecules=333_333)
        # oracle gate at FULL size: one 999,999-atom frame through the C oracle's O(N^2) loop
        # (5.0e11 pair evaluations, minutes on the host cores)
        results.append(measure("cfg4 SPC water 999,999 atoms, 32 of 1,000 cfgs, B=800, rc=8",
                               s, 8.0, 800, 32, 1))
    if "cfg5" in which:
        s = synthetic.ideal_gas(50_000, 8, seed=5)
        for rc in (10.0, 25.0, 49.9):
            for nb in (100, 200, 500, 1000, 2000, 5000):
                results.append(measure(f"cfg5 ideal gas 50,000 atoms, 64 cfgs, B={nb}, rc={rc}", s, rc,
                                       nb, 64, 1 if (nb == 1000) else 0, parity=False))
    print(json.dumps({"gpu": torch.cuda.get_device_name(0), "reps": REPS,
                      "peak_pairs_per_s": PEAK, "results": results}, indent=1))


if __name__ == "__main__":
    main()
