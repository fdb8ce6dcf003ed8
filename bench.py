#!/usr/bin/env python
"""
bench.py — RDF pair-distance evaluations per second (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (restated)

A "step" is one pass of the hot path over one batch of synthetic frames: the RDF histogram of
the workload's configurations for every species pair.  Default workload (configs[1] of
BASELINE.json): 13,824-atom NaCl melt, 1,000 configurations per GPU, 3 species pairs, with the
reference's default cutoff (box/2 - 0.1 = 37.7) and default bins (int(cutoff / 0.01)), parity mode.

value  : pair evaluations / s with the frames already resident in HBM (raw float32 xyz), timed
         with CUDA events on torch's current stream (the library's stream is joined to it), max
         over ranks.
e2e    : same metric through the public host-buffer call (RDFHandle.submit of pinned host frames
         -> H2D on a side stream -> kernels -> D2H of the histogram), wall clock, max over ranks.
roofline / cpu_baseline / clocks: see DESIGN.md §6.

Under torchrun (N > 1) every rank processes its own 1,000 frames (weak scaling; frames are the
shard unit) and the per-GPU histograms are summed with one NCCL allreduce per step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "rdf_pair_distance_evals_per_sec"
UNIT = "pair-evals/s"
FLOP_PER_PAIR = 22  # 3 sub 3 div 3 rint 3 mul 3 sub 3 mul 2 add 1 sqrt 1 cmp


# ----------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------
def make_workload(name: str, rank: int, frames: int):
    from mdsuite_b200 import synthetic
    if name == "cfg2":
        sysm = synthetic.nacl_melt(12, frames, seed=2 + 1000 * rank)
        cutoff = float(sysm.box[0] / 2 - 0.1)  # check_input default, reference :226-229
        nbins = int(cutoff / 0.01)  #            reference :239-242
        desc = (f"cfg2: {sysm.n_atoms}-atom NaCl melt, {frames} configurations per GPU, 3 species "
                f"pairs, default cutoff {cutoff:.1f} A and default {nbins} bins, parity mode")
    elif name == "cfg2_rc10":
        sysm = synthetic.nacl_melt(12, frames, seed=2 + 1000 * rank)
        cutoff, nbins = 10.0, 500
        desc = (f"cfg2 variant: {sysm.n_atoms}-atom NaCl melt, {frames} configurations per GPU, "
                f"cutoff 10 A, 500 bins, parity mode")
    elif name == "cfg1":
        sysm = synthetic.nacl_melt(5, frames, seed=1 + 1000 * rank)
        cutoff, nbins = 10.0, 500
        desc = f"cfg1: 1000-atom NaCl melt, {frames} configurations, cutoff 10 A, 500 bins"
    elif name == "cfg5":
        sysm = synthetic.ideal_gas(50000, frames, seed=5 + 1000 * rank)
        cutoff, nbins = 49.9, 1000
        desc = f"cfg5: 50k-atom ideal gas, {frames} configurations, cutoff 49.9, 1000 bins"
    elif name == "cfg5_rc10":
        sysm = synthetic.ideal_gas(50000, frames, seed=5 + 1000 * rank)
        cutoff, nbins = 10.0, 1000
        desc = (f"cfg5: 50k-atom ideal gas, {frames} configurations, cutoff 10, 1000 bins "
                f"(cell-list path)")
    elif name == "cfg3":
        sysm = synthetic.lj_fluid(47, frames, seed=3 + 1000 * rank, n_atoms=100_000)
        cutoff, nbins = 3.0, 300
        desc = (f"cfg3: {sysm.n_atoms}-atom LJ fluid (rho*=0.8442, L=49.1 sigma), {frames} "
                f"configurations per GPU, cutoff 3.0 sigma, 300 bins, cell-list path, parity mode")
    elif name == "cfg4":
        sysm = synthetic.spc_water(70, frames, seed=4 + 1000 * rank, n_molecules=333_333)
        cutoff, nbins = 8.0, 800
        desc = (f"cfg4: {sysm.n_atoms}-atom SPC water (333,333 molecules, L=215.3 A), {frames} "
                f"configurations per GPU, cutoff 8 A, 800 bins, 3 species pairs, cell-list path, "
                f"parity mode")
    else:
        raise SystemExit(f"unknown workload {name}")
    return sysm, cutoff, nbins, desc


# ----------------------------------------------------------------------------------------------
# nvidia-smi clock sampler (B200_PROFILING.md "clocks DURING the timed region")
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                clk, mx = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            smax = mx
            if t0 - 0.05 <= ts <= t1 + 0.05:
                sm.append(clk)
                try:
                    power.append(float(parts[2]))
                except ValueError:
                    pass
                for n, v in zip(names, parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "power_w_max": max(power) if power else None, "samples": len(sm),
                "reasons": sorted(reasons)}


def counts_tensor(handle, n_pairs: int, nbins: int):
    """Wrap the library's device accumulator as an int64 torch tensor (for NCCL)."""
    import torch
    ptr, _stream = handle.device_counts()

    class _Wrap:
        __cuda_array_interface__ = {"shape": (n_pairs, nbins), "typestr": "<i8", "data": (ptr, False),
                                    "version": 3, "strides": None}
    return torch.as_tensor(_Wrap(), device=f"cuda:{handle.device}")


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "MEASURED_PEAKS.json"
    except OSError:
        return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py touches oracle/)
# ----------------------------------------------------------------------------------------------
def cpu_reference_rate(sysm, cutoff, nbins, n_frames: int):
    """Op-by-op torch-CPU restatement of the reference's batch loop (oracle/rdf_torch_cpu.py): one
    configuration per batch and atom-minibatches of `number_of_configurations` rows, which is what
    the reference's MemoryManager picks at this size."""
    import torch
    from oracle import rdf_torch_cpu
    torch.set_num_threads(os.cpu_count() or 1)
    total = None
    evaluated = 0
    t0 = time.perf_counter()
    for f in range(n_frames):
        batch = np.ascontiguousarray(sysm.positions[f:f + 1].transpose(1, 0, 2))  # (N, 1, 3)
        c, ev = rdf_torch_cpu.rdf_counts(batch, sysm.counts, sysm.box, cutoff, nbins, minibatch=1000)
        total = c if total is None else total + c
        evaluated += ev
    dt = time.perf_counter() - t0
    return evaluated / dt, dt, total, torch.get_num_threads()


def cpu_fused_c_rate(sysm, cutoff, nbins, n_frames: int):
    from oracle import c_oracle
    t0 = time.perf_counter()
    c, ev = c_oracle.rdf_counts_c(sysm.positions[:n_frames], sysm.counts, sysm.box, cutoff, nbins,
                                  skip_first=True, layout=c_oracle.FRAME_MAJOR)
    dt = time.perf_counter() - t0
    return ev / dt, dt, c, c_oracle.max_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    per_step = max(1, min(8, int(60.0 / (4.5 * (args.steps + args.warmup)))))
    sysm, cutoff, nbins, desc = make_workload(args.workload, 0, per_step)
    # same workload as the GPU arm; a step of this arm is a bounded sample of it (see "sample")
    desc = desc.replace(f" {per_step} configurations", f" {args.frames} configurations", 1)
    for _ in range(args.warmup):
        cpu_reference_rate(sysm, cutoff, nbins, per_step)
    t0 = time.perf_counter()
    evaluated = 0.0
    for _ in range(args.steps):
        rate, dt, _c, threads = cpu_reference_rate(sysm, cutoff, nbins, per_step)
        evaluated += rate * dt
    total = time.perf_counter() - t0
    value = evaluated / total
    sample = (f"{per_step} configuration(s) of the workload per step, torch-CPU op-by-op restatement "
              f"of the reference loop (TensorFlow not installable here)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": desc, "frames_per_gpu": args.frames, "atoms": sysm.n_atoms,
                   "nbins": nbins, "cutoff": cutoff,
                   "species_pairs": len(sysm.counts) * (len(sysm.counts) + 1) // 2,
                   "sample": f"each step times {per_step} configuration(s) of the workload; the "
                             f"rate is per pair evaluation, so it does not depend on the count"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------------
def user_api_probe(sysm, cutoff, nbins, pairs_per_step):
    """The workload through the call a user makes: frames ingested into a project's trajectory
    store, then ``experiment.run.RadialDistributionFunction(...)`` (FrameBatcher -> C ABI ->
    normalisation); wall time of the best of three calls after a first one."""
    import tempfile
    import mdsuite_b200 as mds
    with tempfile.TemporaryDirectory() as tmp:
        project = mds.Project("bench", storage_path=tmp)
        exp = project.add_experiment("workload", units="real")
        exp.add_positions(sysm.positions, sysm.species, sysm.box)
        times = []
        for _ in range(4):
            t0 = time.perf_counter()
            exp.run.RadialDistributionFunction(number_of_configurations=sysm.n_frames, cutoff=cutoff,
                                               number_of_bins=nbins, plot=False, save=False)
            times.append(time.perf_counter() - t0)
    best = min(times[1:])
    return {"call": "experiment.run.RadialDistributionFunction(number_of_configurations=%d, ...)"
                    % sysm.n_frames, "seconds": best, "first_call_seconds": times[0],
            "value": pairs_per_step / best, "unit": UNIT,
            "note": "frames read from the frame-major store through the pinned ring; includes "
                    "normalisation; results not saved"}


def cell_list_probe(device: int, fp32_peak: float, frames: int = 160):
    """cfg3-shaped (100k-atom LJ fluid, cutoff 3 sigma, 300 bins) frames resident in HBM through the
    cell-list kernel: candidate-pair rate against the same FP32 roofline, the all-pairs-equivalent
    rate, and a parity gate against the all-pairs kernel on the same frames."""
    import torch
    from mdsuite_b200 import _cuda, synthetic
    sysm = synthetic.lj_fluid(47, 4, seed=3, n_atoms=100_000)
    base = torch.from_numpy(sysm.positions).cuda(device)
    pos = base.repeat((frames + 3) // 4, 1, 1)[:frames].contiguous()
    best = None
    with _cuda.RDFHandle(sysm.counts, 300, 3.0, sysm.box, device=device) as h:
        for _ in range(4):
            h.reset()
            h.submit_device(pos)
            st = h.stats()
            if best is None or st.pair_kernel_ms < best[0]:
                best = (st.pair_kernel_ms, st.submit_ms)
        h.reset()
        h.submit_device(base)
        cells = h.counts()
    with _cuda.RDFHandle(sysm.counts, 300, 3.0, sysm.box, device=device, path="allpairs") as h:
        h.submit_device(base)
        allp = h.counts()
    cand = st.pairs_evaluated / (best[0] * 1e-3)
    return {
        "workload": f"cfg3-shaped: {sysm.n_atoms}-atom LJ fluid, cutoff 3.0, 300 bins, {frames} frames "
                    f"resident in HBM, best of 4",
        "kernel": "rdf_cells_kernel", "cells": list(st.cells),
        "candidate_pair_evals_per_s": cand, "frac_of_fp32_roofline": cand * FLOP_PER_PAIR / fp32_peak,
        "binned_over_candidates": st.pairs_binned / st.pairs_evaluated,
        "allpairs_equivalent_per_s": st.pairs_allpairs_equiv / (best[1] * 1e-3),
        "us_per_frame": best[1] * 1e3 / frames,
        "bins_differing_vs_allpairs_kernel": int((cells != allp).sum()),
    }


def adf_probe(device: int, frames: int = 64, check: bool = True):
    """The sibling path (include/mds_adf.h): angular distribution function of a
    cfg2-shaped NaCl melt (13 824 atoms, cutoff 6 A, 500 bins, norm_power 4) resident in HBM; triples
    per second, and (with the CPU-baseline leg) a parity gate of the triple counts against the NumPy
    oracle on a 512-atom melt."""
    import torch
    from mdsuite_b200 import synthetic
    from mdsuite_b200._cuda import adf as cuda_adf
    sysm = synthetic.nacl_melt(12, 4, seed=2)
    pos = torch.from_numpy(sysm.positions).cuda(device).repeat(frames // 4, 1, 1).contiguous()
    with cuda_adf.ADFHandle(sysm.counts, 500, 6.0, sysm.box, norm_power=4, device=device,
                            max_frames_in_flight=frames) as h:
        out = h.compute_device(pos)
        t_before = h.stats().triples
        best = None
        for _ in range(4):
            h.compute_device(pos, out=out)
            ms = h.stats().kernel_ms
            best = ms if best is None or ms < best else best
        st = h.stats()
    differing = None
    if check:
        from oracle import adf_oracle  # the checker, as in the cpu_baseline leg
        small = synthetic.nacl_melt(4, 1, seed=3)
        with cuda_adf.ADFHandle(small.counts, 90, 4.5, small.box, norm_power=0, device=device) as h:
            got = h.compute(small.positions)
        want = adf_oracle.raw_histograms(small.positions[0], small.counts, small.box, 4.5, 90, 0)
        differing = int((got[0] != want).sum())
    return {
        "workload": f"cfg2-shaped NaCl melt: {sysm.n_atoms} atoms, cutoff 6.0, 500 bins, norm_power 4, "
                    f"{frames} configurations resident in HBM, best of 4",
        "kernel": "adf_kernel<cells>", "cells": list(st.cells),
        "triples_per_call": t_before, "kernel_ms": best,
        "triples_per_s": t_before / (best * 1e-3),
        "max_neighbours": st.max_neighbours,
        "bins_differing_vs_oracle_512_atoms": differing,
    }


def run_ours(args):
    import torch
    from mdsuite_b200 import _cuda

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU path "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the image exports NCCL_DEBUG=VERSION, which makes NCCL print a banner on stdout; the
        # contract is ONE JSON line there
        if "BENCH_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["BENCH_NCCL_DEBUG"]
        else:
            os.environ.pop("NCCL_DEBUG", None)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    sysm, cutoff, nbins, desc = make_workload(args.workload, rank, args.frames)
    n_frames, n_atoms = sysm.n_frames, sysm.n_atoms
    host = torch.from_numpy(sysm.positions).pin_memory()
    dev = host.cuda(non_blocking=False)
    handle = _cuda.RDFHandle(sysm.counts, nbins, cutoff, sysm.box, parity=True, device=local_rank)
    n_pairs = handle.n_pairs
    counts_dev = counts_tensor(handle, n_pairs, nbins) if world > 1 else None
    pairs_per_step = sysm.pairs_per_frame(parity=True) * n_frames

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        if dist is not None:
            handle.wait()  # the previous step's allreduce reads the accumulator
        handle.reset()
        handle.submit_device(dev)
        handle.join()
        if dist is not None:
            dist.all_reduce(counts_dev)

    def step_e2e():
        if dist is not None:
            handle.wait()
        handle.reset()
        handle.submit(host, pinned=True)
        if dist is not None:
            handle.join()
            dist.all_reduce(counts_dev)
            handle.wait()
        return handle.counts()

    # ---- HBM-resident timing (value) -------------------------------------------------------
    for _ in range(max(args.warmup, 0)):
        step_resident()
    handle.stats()  # clears the per-launch event log
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ
                           else int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
    if rank == 0:
        sampler.start()
        t_wait = time.perf_counter()
        while sampler.proc is not None and not sampler.rows and time.perf_counter() - t_wait < 5.0:
            time.sleep(0.05)  # nvidia-smi takes a moment to print its first sample
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        step_resident()
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    st = handle.stats()
    result = handle.counts()
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * pairs_per_step * args.steps / (ms * 1e-3)

    # ---- end to end through host buffers (e2e) ---------------------------------------------
    for _ in range(max(1, min(args.warmup, 3))):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        result_e2e = step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * pairs_per_step * args.steps / e2e_s
    if not np.array_equal(result, result_e2e):
        raise SystemExit("bench.py: HBM-resident and end-to-end histograms differ")

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (the all-pairs pair kernel) -----------------------
    peaks, peaks_src = measured_peaks()
    sm_count = st.sm_count
    f_max = float(peaks.get("sm_max_mhz", 1965.0)) * 1e6
    fp32_peak = sm_count * 128 * f_max  # non-FMA FP32 issue, op/s
    launches = max(1, st.pair_kernel_launches)
    avg_launch_s = st.pair_kernel_ms * 1e-3 / launches
    # pairs the pair kernel distance-tested per launch: every reference pair on the all-pairs
    # path, the candidates of adjacent cells on the cell-list path
    pairs_per_launch = st.pairs_evaluated * args.steps / launches
    kernel_rate = pairs_per_launch / avg_launch_s
    kernel_name = "rdf_cells_kernel" if st.path == _cuda.PATH_CELLLIST else "rdf_allpairs_kernel"
    achieved = kernel_rate * FLOP_PER_PAIR
    kernel_share = st.pair_kernel_ms / ms if world == 1 else None
    atom = {}
    try:
        spread, rnd, mhz = _cuda.bench_shared_atomics(local_rank, max(64, n_pairs * (nbins + 1)), 4096)
        binned_frac = st.pairs_binned / max(st.pairs_evaluated, 1.0)
        # dense binning issues one shared atomic per evaluated pair (out-of-cutoff lanes hit the
        # sink slot); the sparse path only for binned pairs
        atomics_per_pair = 1.0 if st.dense else binned_frac
        atomic_peak_pairs = sm_count * rnd * f_max / max(atomics_per_pair, 1e-12)
        atom = {"lanes_per_clk_per_sm_conflict_free": spread, "lanes_per_clk_per_sm_random_bins": rnd,
                "sm_mhz_during_microbench": mhz, "atomics_per_pair": atomics_per_pair,
                "binned_fraction": binned_frac, "pair_evals_per_s_bound": atomic_peak_pairs,
                "frac": kernel_rate / atomic_peak_pairs}
    except Exception as exc:  # noqa: BLE001
        atom = {"error": str(exc)}
    # DRAM bytes per launch from the committed ncu --set full capture of this kernel
    # (profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per frame of the capture)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as fh:
                per_frame = json.load(fh).get(args.workload, {}).get("dram_bytes_per_frame")
            if per_frame is not None:
                traffic = per_frame * n_frames * args.steps / launches
        except (OSError, ValueError, AttributeError):
            traffic = None
    roofline = {
        "bound": "fp32",
        "bound_note": "FP32 issue rate and shared-memory atomics govern this path (SURVEY.md 8d): a "
                      "distance histogram is not a contraction (no tensor-core work) and the frame "
                      "stream uses <0.1% of HBM bandwidth (see hbm)",
        "kernel": kernel_name, "achieved": achieved / 1e12,
        "peak": fp32_peak / 1e12, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
        "traffic": traffic, "pair_evals_per_s": kernel_rate,
        "pair_evals_per_s_peak": fp32_peak / FLOP_PER_PAIR, "flop_per_pair": FLOP_PER_PAIR,
        "allpairs_equivalent_per_s": pairs_per_step * args.steps / launches / avg_launch_s,
        "launches_timed": st.pair_kernel_launches, "avg_launch_ms": avg_launch_s * 1e3,
        "kernel_share_of_step": kernel_share,
        "peak_source": f"{sm_count} SMs x 128 lanes x sm_max_mhz from {peaks_src} (non-FMA issue rate)",
        "frac_at_sampled_clock": (achieved / (sm_count * 128 * clocks["sm_mhz"] * 1e6)
                                  if clocks and clocks.get("sm_mhz") else None),
        "shared_atomic": atom,
        "hbm": {"algorithmic_bytes_per_launch": 12.0 * n_atoms * n_frames * args.steps / launches,
                "achieved_gbs": 12.0 * n_atoms * n_frames * args.steps / launches / avg_launch_s / 1e9,
                "peak_gbs": peaks.get("hbm_gbs")},
    }

    # ---- the other kernel of the path: a short cell-list measurement (rank 0, N = 1 only) ----
    cell_list = None
    if world == 1 and st.path != _cuda.PATH_CELLLIST and not args.no_cell_probe:
        try:
            cell_list = cell_list_probe(local_rank, fp32_peak)
        except Exception as exc:  # noqa: BLE001
            cell_list = {"error": str(exc)}

    # ---- the sibling structural path, the ADF (rank 0, N = 1 only) ---------------------------
    adf = None
    if world == 1 and not args.no_adf_probe:
        try:
            adf = adf_probe(local_rank, check=not args.no_cpu_baseline)
        except Exception as exc:  # noqa: BLE001
            adf = {"error": str(exc)}

    # ---- the same step through the user-level call (rank 0, N = 1 only) ---------------------
    user_api = None
    if world == 1 and not args.no_user_api:
        try:
            user_api = user_api_probe(sysm, cutoff, nbins, pairs_per_step)
        except Exception as exc:  # noqa: BLE001
            user_api = {"error": str(exc)}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ---------------------------------------
    cpu_baseline = None
    parity = None
    if world == 1 and not args.no_cpu_baseline:
        n_ref = max(1, min(n_frames, args.cpu_frames))
        rate, dt, c_ref, threads = cpu_reference_rate(sysm, cutoff, nbins, n_ref)
        n_c = max(1, min(n_frames, 8 * args.cpu_frames))
        c_rate, c_dt, c_fused, c_threads = cpu_fused_c_rate(sysm, cutoff, nbins, n_c)
        # parity gate on the same frames: CUDA path vs both CPU restatements
        handle.reset()
        handle.submit_device(dev[:n_ref].contiguous())
        gpu_ref = handle.counts().astype(np.int64)
        handle.reset()
        handle.submit_device(dev[:n_c].contiguous())
        gpu_c = handle.counts().astype(np.int64)
        parity = {"frames_checked_op_by_op": n_ref, "bins_differing_op_by_op": int((gpu_ref != c_ref).sum()),
                  "frames_checked_fused_c": n_c, "bins_differing_fused_c": int((gpu_c != c_fused).sum())}
        cpu_baseline = {
            "value": rate, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": (f"first {n_ref} configurations of the workload ({dt:.1f} s), torch-CPU op-by-op "
                       f"restatement of the reference's batch loop (TensorFlow not installable)"),
            "fused_c_port": {"value": c_rate, "unit": UNIT, "cores": c_threads,
                             "sample": f"first {n_c} configurations ({c_dt:.1f} s), oracle/rdf_oracle.c + OpenMP"},
        }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "frames_per_gpu": n_frames, "atoms": n_atoms, "nbins": nbins,
                   "cutoff": cutoff, "species_pairs": n_pairs,
                   "l2": f"inputs larger than L2: {12 * n_atoms * n_frames / 1e6:.0f} MB raw xyz + "
                         f"{12 * n_atoms * n_frames / 1e6:.0f} MB packed per step vs 126 MB L2",
                   "parallelism": f"frames sharded over {world} GPU(s), 1 NCCL allreduce/step"
                   if world > 1 else "1 GPU"},
        "roofline": roofline, "cell_list_path": cell_list, "adf_path": adf, "user_api": user_api,
        "cpu_baseline": cpu_baseline,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": 1e3 * e2e_s / args.steps,
                "h2d_bytes_per_step": int(12 * n_atoms * n_frames),
                "d2h_bytes_per_step": int(8 * n_pairs * nbins)},
        "gpu_launches": int(st.kernel_launches * args.steps),
        "clocks": clocks, "parity": parity,
        "kernel_variant": {"path": st.path_name, "variant": st.variant, "dense": st.dense,
                           "cells": list(st.cells)},
    }
    emit(line)
    handle.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner with the image's
    NCCL_DEBUG=VERSION, torchrun notices) write there too, so file descriptor 1 is pointed at
    stderr for the whole run and the JSON line goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--frames", type=int, default=1000, help="configurations per GPU per step")
    ap.add_argument("--cpu-frames", type=int, default=12,
                    help="configurations timed by the op-by-op CPU baseline (x8 for the C port)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-user-api", action="store_true",
                    help="skip the Project -> Experiment -> run.RadialDistributionFunction timing")
    ap.add_argument("--no-adf-probe", action="store_true",
                    help="skip the short angular-distribution measurement added at N=1")
    ap.add_argument("--no-cell-probe", action="store_true",
                    help="skip the short cfg3-shaped cell-list measurement added at N=1")
    args = ap.parse_args()
    if args.steps < 1:
        raise SystemExit("--steps must be >= 1")
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
