"""The north star names warp-aggregated shared atomics with __match_any_sync for the pair kernel's
histogram; the library uses red.shared.add (ATOMS.POPC.INC) instead.  This script backs the choice
with a number: both builds of the all-pairs kernel on the cfg5 ideal gas (50,000 atoms, cutoff
49.9 — every bin count from 100, where 32 lanes collide all the time, to 5,000).

    make -C mdsuite_b200/_cuda/csrc matchany && python scripts/binning_variants.py > out.json
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, os, sys
import torch
sys.path.insert(0, os.environ["MDS_ROOT"])
from mdsuite_b200 import _cuda, synthetic
sysm = synthetic.ideal_gas(50_000, 2, seed=5)
pos = torch.from_numpy(sysm.positions).cuda().repeat(4, 1, 1).contiguous()
out = {}
for nbins in (100, 200, 500, 1000, 2000, 5000):
    with _cuda.RDFHandle(sysm.counts, nbins, 49.9, sysm.box, path="allpairs") as h:
        best = 1e30
        for _ in range(4):
            h.reset(); h.submit_device(pos); st = h.stats()
            best = min(best, st.pair_kernel_ms)
        counts = h.counts()
    out[str(nbins)] = {"pair_kernel_ms": best, "pair_evals_per_s": st.pairs_evaluated / (best * 1e-3),
                       "checksum": int(counts.sum()), "first_bins": counts[0, :4].tolist()}
print(json.dumps(out))
'''
res = {}
for name, lib in (("atoms_popc_inc", "libmds_rdf.so"), ("match_any_sync", "libmds_rdf_matchany.so")):
    env = dict(os.environ, MDS_ROOT=ROOT, MDS_RDF_LIB=os.path.join(ROOT, "mdsuite_b200", "_cuda", lib))
    p = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    if p.returncode:
        raise SystemExit(p.stderr[-2000:])
    res[name] = json.loads(p.stdout.strip().splitlines()[-1])
same = all(res["atoms_popc_inc"][b]["checksum"] == res["match_any_sync"][b]["checksum"] and
           res["atoms_popc_inc"][b]["first_bins"] == res["match_any_sync"][b]["first_bins"]
           for b in res["atoms_popc_inc"])
print(json.dumps({"workload": "cfg5: 50,000-atom ideal gas, 8 frames resident, cutoff 49.9, all-pairs kernel, "
                              "best of 4", "identical_histograms": same,
                  "ratio_matchany_over_default": {b: res["match_any_sync"][b]["pair_kernel_ms"] /
                                                  res["atoms_popc_inc"][b]["pair_kernel_ms"]
                                                  for b in res["atoms_popc_inc"]}, **res}))
