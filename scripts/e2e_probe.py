"""Exploration helper (not the contract bench): wall time of the cell path fed from pinned host
memory against the same frames resident in HBM, for several submit-block and batch sizes.
Usage: python scripts/e2e_probe.py [frames]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdsuite_b200 import _cuda, synthetic  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 4800
sysm = synthetic.lj_fluid(46, 4, seed=3)
base = torch.from_numpy(sysm.positions)
for maxf in (0, 164):
    for block in (600, 1200, frames):
        host = base.repeat((block + 3) // 4, 1, 1)[:block].contiguous().pin_memory()
        dev = host.cuda()
        with _cuda.RDFHandle(sysm.counts, 300, 3.0, sysm.box, path="celllist",
                             max_frames_in_flight=maxf) as h:
            def resident():
                h.reset()
                for f0 in range(0, frames, block):
                    h.submit_device(dev[:min(block, frames - f0)])
                return h.counts()

            def e2e():
                h.reset()
                for f0 in range(0, frames, block):
                    h.submit(host[:min(block, frames - f0)].numpy(), pinned=True)
                return h.counts()
            out = {}
            for name, fn in (("resident", resident), ("e2e", e2e)):
                fn()
                best = 1e9
                for _ in range(3):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    fn()
                    best = min(best, time.perf_counter() - t0)
                out[name] = best * 1e3
            st = h.stats()
            print(f"maxf={maxf or 'auto'} block={block} frames={frames} resident={out['resident']:.1f} ms "
                  f"e2e={out['e2e']:.1f} ms  gap={out['e2e'] - out['resident']:.1f} ms "
                  f"h2d_if_alone={frames * sysm.n_atoms * 12 / 52e9 * 1e3:.1f} ms", flush=True)
