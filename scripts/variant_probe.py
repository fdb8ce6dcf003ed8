"""Exploration helper: all-pairs variants x pair groups on cfg2 (resident frames)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdsuite_b200 import _cuda, synthetic  # noqa: E402

sysm = synthetic.nacl_melt(12, 8, seed=2)
pos = torch.from_numpy(sysm.positions).cuda().repeat(32, 1, 1).contiguous()  # 256 frames
cutoff = sysm.box[0] / 2 - 0.1
nbins = int(cutoff / 0.01)
PEAK = 148 * 128 * 1.965e9 / 22
for variant in (2, 6, 3):
    for group in (3, 1):
        os.environ["MDS_RDF_VARIANT"] = str(variant)
        os.environ["MDS_RDF_PAIR_GROUP"] = str(group)
        with _cuda.RDFHandle(sysm.counts, nbins, cutoff, sysm.box, path="allpairs") as h:
            best = 1e9
            for _ in range(4):
                h.reset()
                h.submit_device(pos)
                st = h.stats()
                best = min(best, st.submit_ms)
            print(f"variant {variant} pair_group {group}: {best:.3f} ms  "
                  f"frac {st.pairs_evaluated / (best * 1e-3) / PEAK:.3f}", flush=True)
