"""ADF calculator under torchrun: every rank opens the same synthetic NaCl melt in its own project,
the configurations of every batch are split over the ranks, one allreduce makes the batches whole.
Reports the wall time of the user-level call (max over ranks) at this world size.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P scripts/adf_multirank_bench.py [--frames 2048] [--batches 8]
"""
import argparse
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=12)
    ap.add_argument("--frames", type=int, default=2048)
    ap.add_argument("--batches", type=int, default=8)
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    import mdsuite_b200 as mds
    from mdsuite_b200 import synthetic

    rank, local = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sysm = synthetic.nacl_melt(args.cells, 8, seed=2)
    pos = np.tile(sysm.positions, (args.frames // 8, 1, 1))
    with tempfile.TemporaryDirectory() as tmp:
        project = mds.Project(f"adf_rank{rank}", storage_path=tmp)
        exp = project.add_experiment("NaCl", timestep=0.002, temperature=1400.0, units="metal")
        exp.add_positions(pos, sysm.species, sysm.box)
        times = []
        for rep in range(4):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            comp = exp.run.AngularDistributionFunction(
                number_of_configurations=args.frames - rep, cutoff=6.0, start=0,
                number_of_bins=500, batches=args.batches, plot=False)
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], device="cuda")
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            times.append(float(dt.item()))
        checksum = float(np.nansum(np.array(comp["Na_Na_Cl"]["adf"])))
    if rank == 0:
        print(json.dumps({"workload": f"nacl_{sysm.n_atoms}_atoms_{args.frames}_configurations_rc6.0_B500",
                          "n_gpus": world, "batches": args.batches, "seconds_per_call": times,
                          "best_ms": 1e3 * min(times[1:]), "checksum": checksum}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
