"""Exploration helper (not the contract bench): time the cell-list path on cfg3/cfg4/cfg5-shaped
frames resident in HBM.  Usage: python scripts/cells_bench.py [cfg3] [cfg4] [cfg5]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdsuite_b200 import _cuda, synthetic  # noqa: E402

PEAK = 148 * 128 * 1.965e9 / 22


def bench(name, sysm, cutoff, nbins, frames, path="celllist", reps=4):
    base = torch.from_numpy(sysm.positions).cuda()
    pos = base.repeat((frames + base.shape[0] - 1) // base.shape[0], 1, 1)[:frames].contiguous()
    with _cuda.RDFHandle(sysm.counts, nbins, cutoff, sysm.box, path=path,
                         max_frames_in_flight=int(os.environ.get("CB_MAXF", "0"))) as h:
        best_k, best_s = 1e30, 1e30
        for _ in range(reps):
            h.reset()
            h.submit_device(pos)
            st = h.stats()
            best_k = min(best_k, st.pair_kernel_ms)
            best_s = min(best_s, st.submit_ms)
        cand = st.pairs_evaluated / (best_k * 1e-3)
        print(f"{name:28s} path={st.path_name} cells={st.cells} K={st.cell_stencil} frames={frames} "
              f"pair_kernel_ms={best_k:9.3f} submit_ms={best_s:9.3f} cand/s={cand:.3e} "
              f"frac22={cand / PEAK:.3f} binned/cand={st.pairs_binned / st.pairs_evaluated:.3f} "
              f"allpairs_equiv/s={st.pairs_allpairs_equiv / (best_s * 1e-3):.3e} tested/cand={st.pairs_tested / max(st.pairs_evaluated, 1):.2f} "
              f"us/frame={best_s * 1e3 / frames:.1f}", flush=True)


which = sys.argv[1:] or ["cfg3", "cfg5", "cfg4"]
if "cfg3" in which:
    s = synthetic.lj_fluid(46, 4, seed=3)
    bench("cfg3 lj 97k rc3 B300", s, 3.0, 300, int(os.environ.get("CB_FRAMES3", "160")))
if "cfg5" in which:
    s = synthetic.ideal_gas(50_000, 4, seed=5)
    for nb in (100, 1000, 5000):
        bench(f"cfg5 gas 50k rc10 B{nb}", s, 10.0, nb, 64)
if "cfg4" in which:
    s = synthetic.spc_water(69, 1, seed=4)
    bench("cfg4 water 986k rc8 B800", s, 8.0, 800, int(os.environ.get("CB_FRAMES4", "8")))
if "nacl" in which:
    s = synthetic.nacl_melt(12, 4, seed=2)
    bench("nacl 13.8k rc10 B500 cells", s, 10.0, 500, 128, "celllist")
    bench("nacl 13.8k rc10 B500 allp", s, 10.0, 500, 128, "allpairs")
