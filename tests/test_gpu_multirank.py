"""Two ranks, two GPUs, NCCL: the calculator shards the sampled configurations over the ranks,
every rank's handle accumulates its block, ONE allreduce of the device accumulator combines them,
and every rank ends with the full g(r) — equal to the oracle's.  Skipped on boxes with one GPU (the
same logic runs on CPU under gloo in tests/test_distributed_cpu.py)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["MDS_ROOT"])
import mdsuite_b200 as mds
from mdsuite_b200 import synthetic

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
out = os.environ["MDS_OUT"]
for name, sysm, cutoff, nbins in (
        ("nacl", synthetic.nacl_melt(4, 11, seed=21), 9.0, 90),          # all-pairs kernel, 11 frames
        ("lj", synthetic.lj_fluid(16, 5, seed=22), 3.0, 300),           # cell-list kernel, 5 frames
        ("one", synthetic.nacl_melt(5, 1, seed=23), 9.0, 90)):          # 1 frame < 2 ranks: item shards
    project = mds.Project(f"{name}_rank{rank}", storage_path=out)
    exp = project.add_experiment("e")
    exp.add_positions(sysm.positions, sysm.species, sysm.box)
    calc = exp.run.RadialDistributionFunction
    comp = calc(number_of_configurations=sysm.n_frames, number_of_bins=nbins, cutoff=cutoff, plot=False)
    np.save(os.path.join(out, f"{name}_y{rank}.npy"),
            np.array([comp[k]["y"] for k in comp.keys()], dtype=float))
# the ADF: every batch's configurations are split over the ranks, one allreduce of the batch sums
melt = synthetic.nacl_melt(3, 9, seed=24)
project = mds.Project(f"adf_rank{rank}", storage_path=out)
exp = project.add_experiment("e")
exp.add_positions(melt.positions, melt.species, melt.box)
adf = exp.run.AngularDistributionFunction(number_of_configurations=8, cutoff=4.0, start=0,
                                          number_of_bins=60, norm_power=4, batches=3, plot=False)
np.save(os.path.join(out, f"adf{rank}.npy"), np.array([adf[k]["adf"] for k in adf.keys()], dtype=float))
# the collective of the C ABI (mds_rdf_allreduce over a raw ncclComm_t) against torch's process group
from mdsuite_b200 import _cuda
from mdsuite_b200.distributed import allreduce_counts, close_raw_communicators
sysm = synthetic.nacl_melt(3, 6, seed=25)
mine = np.array_split(np.arange(6), dist.get_world_size())[rank]
got = {}
for raw in (True, False):
    with _cuda.RDFHandle(sysm.counts, 80, 8.0, sysm.box, device=local) as h:
        h.submit(sysm.positions[mine])
        allreduce_counts(h, raw=raw)
        got[raw] = h.counts()
        h.reset()                      # work submitted after the collective is ordered behind it
        h.submit(sysm.positions[mine])
        allreduce_counts(h, raw=raw)
        assert np.array_equal(h.counts(), got[raw])
np.save(os.path.join(out, f"raw{rank}.npy"), np.stack([got[True], got[False]]))
close_raw_communicators()
dist.barrier()
dist.destroy_process_group()
'''


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_ranks_nccl_calculator(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MDS_ROOT=ROOT, MDS_OUT=str(tmp_path))
    proc = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
         "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)],
        env=env, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, (proc.stdout + proc.stderr)[-3000:]

    from mdsuite_b200 import synthetic
    from oracle import c_oracle, rdf_oracle
    for name, sysm, cutoff, nbins in (("nacl", synthetic.nacl_melt(4, 11, seed=21), 9.0, 90),
                                      ("lj", synthetic.lj_fluid(16, 5, seed=22), 3.0, 300),
                                      ("one", synthetic.nacl_melt(5, 1, seed=23), 9.0, 90)):
        y0, y1 = np.load(tmp_path / f"{name}_y0.npy"), np.load(tmp_path / f"{name}_y1.npy")
        assert np.array_equal(y0[:, 1:], y1[:, 1:])  # every rank holds the allreduced result
        counts, _ = c_oracle.rdf_counts_c(sysm.positions, sysm.counts, sysm.box, cutoff, nbins,
                                          layout=c_oracle.FRAME_MAJOR)
        names = list(sysm.species)
        want = rdf_oracle.normalise(counts, names, sysm.counts, sysm.box, cutoff, sysm.n_frames)
        for row, key in zip(y0, want):
            assert np.allclose(row[1:], np.array(want[key]["y"])[1:], rtol=1e-12, atol=0)

    # mds_rdf_allreduce(ncclComm_t) == torch's all_reduce == the oracle on all six frames
    sysm = synthetic.nacl_melt(3, 6, seed=25)
    want, _ = c_oracle.rdf_counts_c(sysm.positions, sysm.counts, sysm.box, 8.0, 80,
                                    layout=c_oracle.FRAME_MAJOR)
    for r in range(2):
        both = np.load(tmp_path / f"raw{r}.npy")
        assert np.array_equal(both[0], want.astype(np.uint64))
        assert np.array_equal(both[1], want.astype(np.uint64))

    # ADF: both ranks hold the same curves, equal to the single-process oracle's
    from oracle import adf_oracle
    melt = synthetic.nacl_melt(3, 9, seed=24)
    a0, a1 = np.load(tmp_path / "adf0.npy"), np.load(tmp_path / "adf1.npy")
    assert np.array_equal(a0, a1)
    frames = np.linspace(0, 8, 8, dtype=int)
    want = adf_oracle.adf(melt.positions, melt.counts, ["Na", "Cl"], melt.box, 4.0, 60, 4,
                          [list(b) for b in np.array_split(frames, 3)])
    for row, key in zip(a0, ["Na-Na-Na", "Na-Na-Cl", "Na-Cl-Cl", "Cl-Cl-Cl"]):
        assert np.max(np.abs(row - want[key]["adf"])) <= 2e-6 * want[key]["adf"].max()
