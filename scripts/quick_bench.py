"""Exploration helper (not the contract bench): sweep kernel variants on cfg2-shaped frames."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdsuite_b200 import _cuda, synthetic  # noqa: E402

frames = int(os.environ.get("QB_FRAMES", "128"))
cells = int(os.environ.get("QB_CELLS", "12"))
sysm = synthetic.nacl_melt(cells, 8, seed=2)
pos = torch.from_numpy(sysm.positions).cuda()
pos = pos.repeat((frames + 7) // 8, 1, 1)[:frames].contiguous()
print("atomic bench (spread, random lanes/clk/SM, MHz):", _cuda.bench_shared_atomics(0, 4096, 4096))
print("atomic bench 12k bins:", _cuda.bench_shared_atomics(0, 11313, 4096))
cases = [("default", sysm.box[0] / 2 - 0.1, int((sysm.box[0] / 2 - 0.1) / 0.01)), ("rc10_b500", 10.0, 500)]
variants = [int(v) for v in os.environ.get("QB_VARIANTS", "1,2,3,4,5").split(",")]
for name, cutoff, nbins in cases:
    for dense in (1, 0):
        for variant in variants:
            os.environ["MDS_RDF_VARIANT"] = str(variant)
            os.environ["MDS_RDF_DENSE"] = str(dense)
            try:
                with _cuda.RDFHandle(sysm.counts, nbins, cutoff, sysm.box) as h:
                    best = 1e30
                    for rep in range(4):
                        h.reset()
                        h.submit_device(pos)
                        st = h.stats()
                        best = min(best, st.kernel_ms)
                    rate = st.pairs_evaluated / (best * 1e-3)
                    print(f"{name:10s} dense={dense} variant={variant} kernel_ms={best:8.3f} "
                          f"pairs/s={rate:.3e} frac22={rate * 22 / (148 * 128 * 1.965e9):.3f} "
                          f"binned_frac={st.pairs_binned / st.pairs_evaluated:.3f}", flush=True)
            except Exception as e:  # noqa: BLE001
                print(name, dense, variant, "FAILED", e, flush=True)
