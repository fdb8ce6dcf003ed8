"""world_size-2 gloo tests (CPU) of the multi-rank host logic: configuration sharding, the single
allreduce, and the calculator's rank handling."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mdsuite_b200 import synthetic
from mdsuite_b200.distributed import allreduce_host_counts, shard_configurations
from oracle import c_oracle


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker_histograms(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sysm = synthetic.nacl_melt(2, 7, seed=11)  # 7 frames: uneven split 4 + 3
    frames = shard_configurations(np.arange(7), world, rank)
    part, _ = c_oracle.rdf_counts_c(sysm.positions[frames], sysm.counts, sysm.box, 6.0, 60,
                                    layout=c_oracle.FRAME_MAJOR, n_threads=1)
    total = allreduce_host_counts(part.astype(np.uint64))
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), total)
    dist.destroy_process_group()


def test_sharded_histograms_allreduce_to_the_full_result(tmp_path):
    port = _free_port()
    mp.spawn(_worker_histograms, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sysm = synthetic.nacl_melt(2, 7, seed=11)
    want, _ = c_oracle.rdf_counts_c(sysm.positions, sysm.counts, sysm.box, 6.0, 60,
                                    layout=c_oracle.FRAME_MAJOR)
    for rank in range(2):
        got = np.load(tmp_path / f"rank{rank}.npy")
        assert np.array_equal(got, want.astype(np.uint64))


def _worker_calculator(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mdsuite_b200 as mds
    from mdsuite_b200.calculators.radial_distribution_function import RadialDistributionFunction
    from tests.test_calculator_cpu import OracleHandle
    RadialDistributionFunction.handle_factory = OracleHandle
    sysm = synthetic.nacl_melt(2, 9, seed=12)
    p = mds.Project(f"proj_rank{rank}", storage_path=out_dir)
    e = p.add_experiment("NaCl")
    e.add_positions(sysm.positions, sysm.species, sysm.box)
    comp = e.run.RadialDistributionFunction(number_of_configurations=9, cutoff=6.0,
                                            number_of_bins=30, plot=False)
    np.save(os.path.join(out_dir, f"y{rank}.npy"), np.array(comp["Na_Cl"]["y"]))
    np.save(os.path.join(out_dir, f"frames{rank}.npy"), np.array(OracleHandle.instances[-1].frames))
    dist.destroy_process_group()


def test_calculator_shards_frames_over_ranks(tmp_path):
    port = _free_port()
    mp.spawn(_worker_calculator, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    y0, y1 = np.load(tmp_path / "y0.npy"), np.load(tmp_path / "y1.npy")
    assert np.array_equal(y0[1:], y1[1:])
    assert sorted([int(np.load(tmp_path / f"frames{r}.npy")) for r in range(2)]) == [4, 5]
    from oracle import rdf_oracle
    sysm = synthetic.nacl_melt(2, 9, seed=12)
    counts, _ = c_oracle.rdf_counts_c(sysm.positions, sysm.counts, sysm.box, 6.0, 30,
                                      layout=c_oracle.FRAME_MAJOR)
    want = np.array(rdf_oracle.normalise(counts, ["Na", "Cl"], sysm.counts, sysm.box, 6.0, 9)["Na_Cl"]["y"])
    assert np.allclose(y0[1:], want[1:], rtol=1e-12)


def _worker_adf(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mdsuite_b200 as mds
    from mdsuite_b200.calculators.angular_distribution_function import AngularDistributionFunction
    from tests.test_adf_calculator_cpu import OracleADFHandle
    AngularDistributionFunction.handle_factory = OracleADFHandle
    OracleADFHandle.calls = []
    sysm = synthetic.nacl_melt(2, 8, seed=13)
    p = mds.Project(f"adf_rank{rank}", storage_path=out_dir)
    e = p.add_experiment("NaCl")
    e.add_positions(sysm.positions, sysm.species, sysm.box)
    comp = e.run.AngularDistributionFunction(number_of_configurations=7, cutoff=4.0, start=0,
                                             number_of_bins=40, norm_power=2, batches=2, plot=False)
    np.save(os.path.join(out_dir, f"adf{rank}.npy"),
            np.stack([np.array(comp[k]["adf"]) for k in comp.keys()]))
    np.save(os.path.join(out_dir, f"adf_calls{rank}.npy"), np.array(OracleADFHandle.calls))
    dist.destroy_process_group()


def test_adf_calculator_shards_every_batch_over_ranks(tmp_path):
    """7 configurations in 2 batches (4 + 3), each batch split over 2 ranks (2 + 2, 2 + 1); one
    allreduce of the per-batch weight sums; both ranks get the single-process curves."""
    from oracle import adf_oracle
    port = _free_port()
    mp.spawn(_worker_adf, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a0, a1 = np.load(tmp_path / "adf0.npy"), np.load(tmp_path / "adf1.npy")
    assert np.array_equal(a0, a1)
    assert np.load(tmp_path / "adf_calls0.npy").tolist() == [2, 2]
    assert np.load(tmp_path / "adf_calls1.npy").tolist() == [2, 1]
    sysm = synthetic.nacl_melt(2, 8, seed=13)
    frames = np.linspace(0, 7, 7, dtype=int)
    want = adf_oracle.adf(sysm.positions, sysm.counts, ["Na", "Cl"], sysm.box, 4.0, 40, 2,
                          [list(b) for b in np.array_split(frames, 2)])
    for row, key in zip(a0, ["Na-Na-Na", "Na-Na-Cl", "Na-Cl-Cl", "Cl-Cl-Cl"]):
        assert np.max(np.abs(row - want[key]["adf"])) <= 2e-6 * want[key]["adf"].max()


def _worker_shared_project(rank, world, port, out_dir):
    """Both ranks open the SAME project (the normal torchrun case).  Rank 0 creates it first."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mdsuite_b200 as mds
    from mdsuite_b200.calculators.radial_distribution_function import RadialDistributionFunction
    from tests.test_calculator_cpu import OracleHandle
    RadialDistributionFunction.handle_factory = OracleHandle
    sysm = synthetic.nacl_melt(2, 6, seed=14)
    if rank == 0:
        p = mds.Project("shared", storage_path=out_dir)
        e = p.add_experiment("NaCl")
        e.add_positions(sysm.positions, sysm.species, sysm.box)
    dist.barrier()
    if rank != 0:
        p = mds.Project("shared", storage_path=out_dir)
        e = p.experiments["NaCl"]
    runs = []
    for _ in range(2):  # second call: a cache hit, decided by rank 0 for everybody
        before = len(OracleHandle.instances)
        comp = e.run.RadialDistributionFunction(number_of_configurations=6, cutoff=6.0,
                                                number_of_bins=30, plot=False)
        runs.append((len(OracleHandle.instances) - before, np.array(comp["Na_Cl"]["y"])))
    np.save(os.path.join(out_dir, f"shared_y{rank}.npy"), np.stack([r[1] for r in runs]))
    np.save(os.path.join(out_dir, f"shared_ran{rank}.npy"), np.array([r[0] for r in runs]))
    dist.barrier()
    if rank == 0:
        rows = p.database.count_computations("NaCl", "Radial_Distribution_Function")
        np.save(os.path.join(out_dir, "shared_rows.npy"), np.array([rows]))
    dist.destroy_process_group()


def test_ranks_sharing_one_project_write_one_row_and_agree_on_the_cache(tmp_path):
    port = _free_port()
    mp.spawn(_worker_shared_project, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    y0, y1 = np.load(tmp_path / "shared_y0.npy"), np.load(tmp_path / "shared_y1.npy")
    assert np.array_equal(y0[0][1:], y1[0][1:]) and np.array_equal(y0[1][1:], y1[1][1:])
    assert np.array_equal(y0[0][1:], y0[1][1:])
    # both ranks ran the analysis once (first call), neither on the cache hit
    assert np.load(tmp_path / "shared_ran0.npy").tolist() == [1, 0]
    assert np.load(tmp_path / "shared_ran1.npy").tolist() == [1, 0]
    assert np.load(tmp_path / "shared_rows.npy").tolist() == [1]


def test_shard_configurations_edge_cases():
    assert shard_configurations(np.arange(3), 8, 5).size == 0  # more ranks than frames
    parts = [shard_configurations(np.arange(10), 4, r) for r in range(4)]
    assert np.array_equal(np.concatenate(parts), np.arange(10))
    with pytest.raises(ValueError):
        shard_configurations(np.arange(3), 2, 2)
