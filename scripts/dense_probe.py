import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdsuite_b200 import _cuda, synthetic
s = synthetic.ideal_gas(50_000, 8, seed=5)
pos = torch.from_numpy(np.tile(s.positions, (4,1,1))).cuda()
for rc in (15.0, 20.0, 25.0, 30.0):
    for dense in ("0", "1"):
        os.environ["MDS_RDF_DENSE"] = dense
        with _cuda.RDFHandle(s.counts, 1000, rc, s.box, parity=False, path="allpairs") as h:
            best = 1e9
            for _ in range(3):
                h.reset(); h.submit_device(pos); st = h.stats(); best = min(best, st.pair_kernel_ms)
            print(rc, "dense", dense, "phi=%.3f" % (st.pairs_binned/st.pairs_evaluated), "ms %.2f" % best, "frac %.3f" % (st.pairs_evaluated/(best*1e-3)/(148*128*1.965e9/22)))
