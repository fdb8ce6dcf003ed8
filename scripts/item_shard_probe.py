"""What one rank of an 8-way item-sharded run costs (strong scaling inside a frame), measured on a
single GPU: shard 0 of 8 against the unsharded handle, for one 1M-atom water frame (cell path) and
one 100k-atom frame forced onto the all-pairs path."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdsuite_b200 import _cuda, synthetic  # noqa: E402


def probe(name, sysm, cutoff, nbins, path):
    dev = torch.from_numpy(sysm.positions).cuda()
    with _cuda.RDFHandle(sysm.counts, nbins, cutoff, sysm.box, path=path) as h:
        out = {}
        for count in (1, 8):
            h.set_item_shard(0, count)
            best = 1e9
            for _ in range(4):
                h.reset()
                h.submit_device(dev)
                st = h.stats()
                best = min(best, st.submit_ms)
            out[count] = best
        print(f"{name}: unsharded {out[1]:.3f} ms, shard 0 of 8 {out[8]:.3f} ms "
              f"-> {out[1] / out[8]:.2f}x faster (ideal: 8x; build, launch and histogram merge are per rank)")


probe("cfg4 one frame (cell list)", synthetic.spc_water(70, 1, seed=4, n_molecules=333_333), 8.0, 800, None)
probe("100k LJ one frame (all pairs)", synthetic.lj_fluid(47, 1, seed=3, n_atoms=100_000), 3.0, 300, "allpairs")
