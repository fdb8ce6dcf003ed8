"""Short single-GPU workloads for ncu captures:
python scripts/profile_target.py allpairs|cells|cellbuild|cells4"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdsuite_b200 import _cuda, synthetic  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "allpairs"
if which == "allpairs":
    s = synthetic.nacl_melt(12, 4, seed=2)
    cutoff = s.box[0] / 2 - 0.1
    nbins, frames, path = int(cutoff / 0.01), 48, "allpairs"
elif which == "cells":
    s = synthetic.lj_fluid(47, 4, seed=3, n_atoms=100_000)
    cutoff, nbins, frames, path = 3.0, 300, 64, "celllist"
elif which == "cellbuild":  # two waves of the one-kernel build (-k regex:cell_build)
    s = synthetic.lj_fluid(47, 4, seed=3, n_atoms=100_000)
    cutoff, nbins, frames, path = 3.0, 300, 296, "celllist"
else:
    s = synthetic.spc_water(69, 1, seed=4)
    cutoff, nbins, frames, path = 8.0, 800, 2, "celllist"
base = torch.from_numpy(s.positions).cuda()
pos = base.repeat((frames + base.shape[0] - 1) // base.shape[0], 1, 1)[:frames].contiguous()
with _cuda.RDFHandle(s.counts, nbins, cutoff, s.box, path=path, max_frames_in_flight=frames) as h:
    for _ in range(3):
        h.reset()
        h.submit_device(pos)
        st = h.stats()
    print(which, st.path_name, "pair_kernel_ms", st.pair_kernel_ms, "cand/s",
          st.pairs_evaluated / (st.pair_kernel_ms * 1e-3))
