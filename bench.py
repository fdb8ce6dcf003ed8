#!/usr/bin/env python
"""
bench.py — RDF pair-distance evaluations per second (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (restated)

A "step" is one pass of the hot path over one batch of synthetic frames: the RDF histogram of
the workload's configurations for every species pair.  Default workload (configs[1] of
BASELINE.json): 13,824-atom NaCl melt, 1,000 configurations IN TOTAL, 3 species pairs, with the
reference's default cutoff (box/2 - 0.1 = 37.7) and default bins (int(cutoff / 0.01)), parity mode.

value  : pair evaluations / s with the frames already resident in HBM (raw float32 xyz), timed
         with CUDA events on torch's current stream (the library's stream is joined to it), max
         over ranks.
e2e    : same metric through the public host-buffer call (RDFHandle.submit of pinned host frames
         -> H2D on a side stream -> kernels -> D2H of the histogram), wall clock, max over ranks.
roofline / cpu_baseline / clocks: see DESIGN.md §6.

Under torchrun (N > 1) the SAME fixed workload is split over the ranks (strong scaling): rank g
takes block g of np.array_split(configurations, N) — the primitive the reference splits batches
with (calculators/radial_distribution_function.py:594) — and the per-GPU histograms are summed
with one NCCL allreduce per step.  Every line also carries `target_config`: BASELINE.json's
north-star target (configs[2]: 100,000-atom LJ fluid, 10,000 configurations, cell-list path) at
full size, sharded the same way, with its own parity block.
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
def make_workload(name: str, frames: int):
    """The whole workload (every rank builds the same trajectory and takes its block of it)."""
    from mdsuite_b200 import synthetic
    if name == "cfg2":
        sysm = synthetic.nacl_melt(12, frames, seed=2)
        cutoff = float(sysm.box[0] / 2 - 0.1)  # check_input default, reference :226-229
        nbins = int(cutoff / 0.01)  #            reference :239-242
        desc = (f"cfg2: {sysm.n_atoms}-atom NaCl melt, {frames} configurations, 3 species "
                f"pairs, default cutoff {cutoff:.1f} A and default {nbins} bins, parity mode")
    elif name == "cfg2_rc10":
        sysm = synthetic.nacl_melt(12, frames, seed=2)
        cutoff, nbins = 10.0, 500
        desc = (f"cfg2 variant: {sysm.n_atoms}-atom NaCl melt, {frames} configurations, "
                f"cutoff 10 A, 500 bins, parity mode")
    elif name == "cfg1":
        sysm = synthetic.nacl_melt(5, frames, seed=1)
        cutoff, nbins = 10.0, 500
        desc = f"cfg1: 1000-atom NaCl melt, {frames} configurations, cutoff 10 A, 500 bins"
    elif name == "cfg5":
        sysm = synthetic.ideal_gas(50000, frames, seed=5)
        cutoff, nbins = 49.9, 1000
        desc = f"cfg5: 50k-atom ideal gas, {frames} configurations, cutoff 49.9, 1000 bins"
    elif name == "cfg5_rc10":
        sysm = synthetic.ideal_gas(50000, frames, seed=5)
        cutoff, nbins = 10.0, 1000
        desc = (f"cfg5: 50k-atom ideal gas, {frames} configurations, cutoff 10, 1000 bins "
                f"(cell-list path)")
    elif name == "cfg3":
        sysm = synthetic.lj_fluid(47, frames, seed=3, n_atoms=100_000)
        cutoff, nbins = 3.0, 300
        desc = (f"cfg3: {sysm.n_atoms}-atom LJ fluid (rho*=0.8442, L=49.1 sigma), {frames} "
                f"configurations, cutoff 3.0 sigma, 300 bins, cell-list path, parity mode")
    elif name == "cfg4":
        sysm = synthetic.spc_water(70, frames, seed=4, n_molecules=333_333)
        cutoff, nbins = 8.0, 800
        desc = (f"cfg4: {sysm.n_atoms}-atom SPC water (333,333 molecules, L=215.3 A), {frames} "
                f"configurations, cutoff 8 A, 800 bins, 3 species pairs, cell-list path, "
                f"parity mode")
    else:
        raise SystemExit(f"unknown workload {name}")
    return sysm, cutoff, nbins, desc


def workload_config(desc, n_total, n_atoms, nbins, cutoff, n_pairs):
    """The `config` object of a bench line: what the workload IS, identical for both arms and for
    every N (how a run lays it out over GPUs is in `run_layout`)."""
    return {"workload": desc, "frames_total": int(n_total), "atoms": int(n_atoms), "nbins": int(nbins),
            "cutoff": float(cutoff), "species_pairs": int(n_pairs),
            "l2": f"a step reads {12 * n_atoms * n_total / 1e6:.0f} MB raw xyz + as much packed, "
                  f"split over the ranks; a rank whose share (raw + packed) is above 1.5 x 126 MB L2 runs its steps "
                  f"back to back, a rank below it flushes L2 (256 MB write, outside the timed "
                  f"intervals) before every timed step - which one this run did is in run_layout.l2"}


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


L2_BYTES = 126e6  # B200


class TimedSteps:
    """K steps timed on the device.  If a step's inputs are larger than L2 the K steps are one
    CUDA-event interval; otherwise L2 is flushed (a 256 MB write) before every step and the steps
    are timed one by one, the flushes outside the intervals (timing rules of the bench contract)."""

    def __init__(self, device, bytes_per_step: float):
        import torch
        self.torch = torch
        self.flush = bytes_per_step < 1.5 * L2_BYTES
        self.buf = torch.empty(256 << 20, dtype=torch.uint8, device=device) if self.flush else None
        self.note = ("L2 flushed (256 MB write) before every timed step, steps timed one by one"
                     if self.flush else "inputs of a step larger than L2, no flush")

    def run(self, step, k: int) -> float:
        """Milliseconds of k calls of step() on torch's current stream."""
        torch = self.torch
        if not self.flush:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(k):
                step()
            ev1.record()
            torch.cuda.synchronize()
            return ev0.elapsed_time(ev1)
        events = []
        for _ in range(k):
            self.buf.zero_()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            step()
            ev1.record()
            events.append((ev0, ev1))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in events)


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
    sysm, cutoff, nbins, desc = make_workload(args.workload, per_step)
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
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(desc, args.frames, sysm.n_atoms, nbins, cutoff,
                                  len(sysm.counts) * (len(sysm.counts) + 1) // 2),
        "run_layout": {"sample": f"each step times {per_step} configuration(s) of the workload on "
                                 f"the host cores; the rate is per pair evaluation, so it does not "
                                 f"depend on the count"},
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


def load_integration_stub():
    """The reference-side binding exactly as INTEGRATION.md §2 prints it (the first python block of
    that section), executed with the library name resolved to the in-tree build."""
    from mdsuite_b200 import _cuda
    with open(os.path.join(ROOT, "INTEGRATION.md")) as fh:
        text = fh.read()
    section = text[text.index("## 2."):]
    code = section[section.index("```python") + len("```python"):]
    code = code[:code.index("```")]
    code = code.replace('ctypes.CDLL("libmds_rdf.so")', "ctypes.CDLL(%r)" % _cuda.LIB_PATH)
    namespace = {}
    exec(compile(code, "INTEGRATION.md#2", "exec"), namespace)  # noqa: S102 - our own document
    return namespace["RDFHandle"]


def reference_binding_probe(sysm, cutoff, nbins, device, e2e_value):
    """The call a reference maintainer would make: the literal ctypes stub,
    PAGEABLE NumPy batches in the reference's (N, C, 3) layout, one submit per batch as in the loop
    it replaces (reference calculators/radial_distribution_function.py:846-885) — with C = 1 (what
    the reference's memory manager picks at this size) and C = 64."""
    RDFHandle = load_integration_stub()
    n_frames = sysm.n_frames
    pairs = sysm.pairs_per_frame(parity=True) * n_frames
    out = {"stub": "INTEGRATION.md section 2, executed as printed", "host_memory": "pageable numpy",
           "layout": "(N, C, 3) atom-major", "e2e_value_pinned_frame_major": e2e_value}
    want = None
    for C in (1, 64):
        batches = [np.ascontiguousarray(sysm.positions[f0:f0 + C].transpose(1, 0, 2))
                   for f0 in range(0, n_frames, C)]
        # one handle per calculator run, as in the patched loop; the first pass over the batches
        # pays for the handle's lazily allocated staging buffers (reported separately), the timed
        # passes add to the same accumulator (the stub has no reset): k passes = k x the histogram
        handle = RDFHandle(sysm.counts, nbins, cutoff, sysm.box, device=device)
        t0 = time.perf_counter()
        for b in batches:
            handle.submit(b)
        first = handle.counts()
        first_s = time.perf_counter() - t0
        best = None
        for k in range(2, 5):
            t0 = time.perf_counter()
            for b in batches:
                handle.submit(b)
            counts = handle.counts()
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
            if not np.array_equal(counts, first * np.uint64(k)):
                raise RuntimeError("reference binding: accumulated histogram is not k x the first pass")
        handle.close()
        counts = first
        if want is None:
            want = counts
        out[f"C{C}"] = {"batches": len(batches), "seconds": best, "first_pass_seconds": first_s,
                        "value": pairs / best, "unit": UNIT,
                        "ratio_to_e2e": pairs / best / e2e_value,
                        "same_histogram": bool(np.array_equal(counts, want))}
    return out


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


FULL_CONFIGS = {
    # BASELINE.json configs[2], the north-star target
    "cfg3": dict(n_total=10_000, n_base=40, block=1200, cutoff=3.0, nbins=300, oracle="fused_c",
                 system=lambda n: __import__("mdsuite_b200.synthetic", fromlist=["x"]).lj_fluid(
                     47, n, seed=3, n_atoms=100_000),
                 workload="cfg3 (BASELINE.json configs[2], the north-star target) at full size: "
                          "100,000-atom LJ fluid (rho*=0.8442, L=49.1 sigma), 10,000 configurations in "
                          "total, cutoff 3.0 sigma, 300 bins, cell-list path, parity mode"),
    # BASELINE.json configs[3]
    "cfg4": dict(n_total=1_000, n_base=2, block=100, cutoff=8.0, nbins=800, oracle="allpairs_kernel",
                 system=lambda n: __import__("mdsuite_b200.synthetic", fromlist=["x"]).spc_water(
                     70, n, seed=4, n_molecules=333_333),
                 workload="cfg4 (BASELINE.json configs[3]) at full size: 999,999-atom SPC water (333,333 "
                          "molecules, L=215.3 A), 1,000 configurations in total, cutoff 8 A, 800 bins, 3 "
                          "species pairs (O-O, O-H, H-H), cell-list path, parity mode"),
}


def target_config_probe(args, world, rank, local_rank, dist, comm, fp32_peak, name="cfg3"):
    """A BASELINE.json multi-GPU configuration at FULL size (cfg3: the north-star target, the
    100,000-atom LJ fluid over 10,000 configurations; cfg4: the 999,999-atom water over 1,000),
    the configurations split over the ranks with np.array_split, one NCCL allreduce per step.

    The trajectory is n_base distinct seeded frames used n_total / n_base times each (frame i is
    base frame i % n_base): generating 12 GB of displacements would take longer than every
    measurement here, and the kernels do the same work on a repeated frame.  That also gives the
    parity gate: the allreduced histogram must equal (n_total / n_base) x the single-handle
    histogram of the base frames; one base frame is checked against the fused-C oracle (cfg3) or,
    where the oracle's O(N^2) loop takes minutes (cfg4: scripts/run_configs.py does that once per
    round), against the independently written all-pairs kernel.

    resident: this rank's frames in HBM as raw xyz.  e2e: pinned host frames through
    RDFHandle.submit (a block of frames submitted repeatedly: every byte crosses PCIe every step)."""
    import torch
    from mdsuite_b200 import _cuda
    spec = FULL_CONFIGS[name]
    n_total, n_base = spec["n_total"], spec["n_base"]
    steps = max(1, min(args.steps, args.target_steps))
    sysm = spec["system"](n_base)
    cutoff, nbins = spec["cutoff"], spec["nbins"]
    mine = np.array_split(np.arange(n_total), world)[rank]
    base_dev = torch.from_numpy(sysm.positions).cuda(local_rank)
    idx = torch.from_numpy(mine % n_base).cuda(local_rank)
    dev = base_dev[idx].contiguous() if len(mine) else base_dev[:0]
    # a multiple of the base frames: every block is the same sequence of frames
    block = min(len(mine), spec["block"])
    host = torch.from_numpy(sysm.positions[mine[:block] % n_base]).pin_memory() if block else None
    handle = _cuda.RDFHandle(sysm.counts, nbins, cutoff, sysm.box, parity=True, device=local_rank)
    counts_dev = counts_tensor(handle, handle.n_pairs, nbins) if dist is not None else None
    pairs_total = sysm.pairs_per_frame(parity=True) * n_total

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def collective():
        """ONE allreduce of the [n_pairs][nbins] uint64 accumulator, ordered after the kernels."""
        if dist is None:
            return
        if comm is not None:
            handle.allreduce(comm)  # on the library's stream
        else:
            handle.join()
            dist.all_reduce(counts_dev)
            handle.wait()

    def step_resident():
        if dist is not None and comm is None:
            handle.wait()
        handle.reset()
        if len(mine):
            handle.submit_device(dev)
        collective()
        handle.join()

    def step_e2e():
        if dist is not None and comm is None:
            handle.wait()
        handle.reset()
        done = 0
        while done < len(mine):
            n = min(block, len(mine) - done)
            # frames done .. done + n of this rank are base frames (mine[done + k] % 40): the pinned
            # block holds exactly that sequence when `done` is a multiple of the block
            handle.submit(host[:n], pinned=True)
            done += n
        collective()
        return handle.counts()

    timer = TimedSteps(torch.device("cuda", local_rank), 24.0 * sysm.n_atoms * len(mine))

    def timed(fn):
        fn()
        handle.stats()
        barrier()
        box = {}

        def one():
            box["out"] = fn()
        t0 = time.perf_counter()
        ms = timer.run(one, steps)
        barrier()
        wall = time.perf_counter() - t0
        out = box.get("out")
        if dist is not None:
            t = torch.tensor([ms, wall], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0].item()), float(t[1].item())
        return ms / steps, wall / steps, out

    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ
                           else int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
    if rank == 0:
        sampler.start()
    t_clk0 = time.perf_counter()
    res_ms, _w, _ = timed(step_resident)
    clocks = sampler.stop(t_clk0, time.perf_counter()) if rank == 0 else None
    st = handle.stats()
    result = handle.counts()
    _ms, e2e_s, result_e2e = timed(step_e2e)

    # local statistics of one un-reduced pass (the accumulator holds the global sum after a step)
    handle.reset()
    if len(mine):
        handle.submit_device(dev)
    st_local = handle.stats()
    parity = None
    if rank == 0:
        with _cuda.RDFHandle(sysm.counts, nbins, cutoff, sysm.box, parity=True, device=local_rank) as h:
            h.submit_device(base_dev)
            h40 = h.counts()
            h.reset()
            h.submit_device(base_dev[:1].contiguous())
            h1 = h.counts()
        parity = {
            "frames_total": n_total,
            "expected": f"{n_total // n_base} x the single-handle histogram of the {n_base} base "
                        f"frames (rank 0)",
            "bins_differing_resident": int((result != h40 * np.uint64(n_total // n_base)).sum()),
            "bins_differing_e2e": int((result_e2e != h40 * np.uint64(n_total // n_base)).sum()),
        }
        if spec["oracle"] == "allpairs_kernel":
            with _cuda.RDFHandle(sysm.counts, nbins, cutoff, sysm.box, parity=True, device=local_rank,
                                 path="allpairs") as h:
                h.submit_device(base_dev[:1].contiguous())
                parity["frames_checked_allpairs_kernel"] = 1
                parity["bins_differing_allpairs_kernel"] = int((h.counts() != h1).sum())
        elif not args.no_cpu_baseline:
            from oracle import c_oracle  # the checker
            want, _ev = c_oracle.rdf_counts_c(sysm.positions[:1], sysm.counts, sysm.box, cutoff, nbins,
                                              skip_first=True, layout=c_oracle.FRAME_MAJOR)
            parity["frames_checked_fused_c"] = 1
            parity["bins_differing_fused_c"] = int((h1.astype(np.int64) != want).sum())
    handle_pairs = handle.n_pairs
    handle.close()
    del dev, host, base_dev
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    launches = max(1, st.pair_kernel_launches)
    kernel_s = st.pair_kernel_ms * 1e-3
    cand_rate = st_local.pairs_evaluated * steps / max(kernel_s, 1e-12)
    tested_rate = st_local.pairs_tested * steps / max(kernel_s, 1e-12)
    return {
        "workload": spec["workload"],
        "frames_total": n_total, "frames_rank0": int(len(mine)), "distinct_frames": n_base,
        "steps": steps, "n_gpus": world,
        "ms_per_step_resident": res_ms, "ms_per_step_e2e": 1e3 * e2e_s,
        "allpairs_equivalent_per_s_resident": pairs_total / (res_ms * 1e-3),
        "allpairs_equivalent_per_s_e2e": pairs_total / e2e_s,
        "h2d_bytes_per_step": int(12 * sysm.n_atoms * n_total), "d2h_bytes_per_step": int(8 * nbins * handle_pairs),
        "kernel": "rdf_cells_kernel", "cells": list(st.cells), "cell_stencil": list(st.cell_stencil),
        "rank0_pair_kernel_ms_per_step": st.pair_kernel_ms / steps,
        "rank0_pair_kernel_launches_per_step": launches / steps,
        "rank0_kernel_share_of_resident_step": st.pair_kernel_ms / steps / res_ms,
        "candidate_pair_evals_per_s_per_gpu": cand_rate,
        "frac_of_fp32_roofline_candidates": cand_rate * FLOP_PER_PAIR / fp32_peak,
        "tested_pair_evals_per_s_per_gpu": tested_rate,
        "frac_of_fp32_roofline_tested": tested_rate * FLOP_PER_PAIR / fp32_peak,
        "binned_over_candidates": st_local.pairs_binned / max(st_local.pairs_evaluated, 1.0),
        "clocks_resident": clocks,
        "frac_of_fp32_roofline_candidates_at_sampled_clock":
            (cand_rate * FLOP_PER_PAIR / (fp32_peak * clocks["sm_mhz"] / clocks["sm_max_mhz"])
             if clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz") else None),
        "parity": parity,
    }


def cfg5_sweep_probe(args, world, rank, local_rank, dist, comm, fp32_peak):
    """BASELINE.json configs[4]: the uniform-random ideal gas (50,000 atoms, 64 configurations:
    8 distinct x 8) over the bins / cutoff sweep, configurations split over the ranks, in CORRECTED
    mode (every pair counted) so that the analytic g(r) = 1 can be checked: z = (count - expected) /
    sqrt(expected x repeats) per bin with the exact shell volumes."""
    import torch
    from mdsuite_b200 import _cuda, synthetic
    n_base, repeats = 8, 8
    n_total = n_base * repeats
    sysm = synthetic.ideal_gas(50_000, n_base, seed=5)
    mine = np.array_split(np.arange(n_total), world)[rank]
    base_dev = torch.from_numpy(sysm.positions).cuda(local_rank)
    dev = base_dev[torch.from_numpy(mine % n_base).cuda(local_rank)].contiguous() if len(mine) \
        else base_dev[:0]
    n, L = sysm.n_atoms, float(sysm.box[0])
    pairs_total = sysm.pairs_per_frame(parity=False) * n_total
    timer = TimedSteps(torch.device("cuda", local_rank), 24.0 * n * len(mine))  # 38 MB at N = 1
    rows = []
    for cutoff in (10.0, 49.9):
        for nbins in (100, 1000, 5000):
            with _cuda.RDFHandle(sysm.counts, nbins, cutoff, sysm.box, parity=False,
                                 device=local_rank) as handle:
                counts_dev = counts_tensor(handle, handle.n_pairs, nbins) if dist is not None else None

                def step():
                    if dist is not None and comm is None:
                        handle.wait()
                    handle.reset()
                    if len(mine):
                        handle.submit_device(dev)
                    if dist is not None:
                        if comm is not None:
                            handle.allreduce(comm)
                        else:
                            handle.join()
                            dist.all_reduce(counts_dev)
                            handle.wait()
                    handle.join()

                step()
                if dist is not None:
                    dist.barrier()
                torch.cuda.synchronize()
                ms = timer.run(step, 2) / 2
                if dist is not None:
                    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms = float(t.item())
                got = handle.counts()
                st = handle.stats()
            if rank != 0:
                continue
            k = np.arange(nbins)
            dr = cutoff / nbins
            shell = 4 * np.pi / 3 * (((k + 1) * dr) ** 3 - (k * dr) ** 3)
            expect = n_total * n * (n - 1) / 2 * shell / L ** 3
            ok = ((k + 1) * dr <= L / 2) & (expect > 50)
            z = (got[0][ok].astype(np.float64) - expect[ok]) / np.sqrt(expect[ok] * repeats)
            rows.append({"cutoff": cutoff, "nbins": nbins, "path": st.path_name,
                         "cell_stencil": list(st.cell_stencil), "ms_per_step": ms,
                         "allpairs_equivalent_per_s": pairs_total / (ms * 1e-3),
                         "ideal_gas_gate": {"bins_checked": int(ok.sum()),
                                            "max_abs_z": float(np.abs(z).max()) if ok.any() else None,
                                            "g_mean": float(np.mean(got[0][ok] / expect[ok])) if ok.any() else None}})
    del dev, base_dev
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    return {"workload": "cfg5 (BASELINE.json configs[4]): 50,000-atom uniform-random ideal gas, L=100, 64 "
                        "configurations in total (8 distinct x 8), corrected mode, bins x cutoff sweep",
            "n_gpus": world, "frames_total": n_total, "l2": timer.note, "rows": rows}


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
    # the feeding process next to its GPU: pinned buffers are first-touched on that NUMA node
    from mdsuite_b200.utils import affinity
    binding = affinity.bind_to_gpu(local_rank) if not args.no_numa_binding else {"bound": False,
                                                                                 "reason": "--no-numa-binding"}
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL's log (NCCL_DEBUG) stays as the environment sets it: fd 1 already points at stderr
        # (claim_stdout), so nothing but the JSON line reaches stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # the collective goes through the C ABI (mds_rdf_allreduce over a raw ncclComm_t, on the
    # library's stream) unless --torch-collective asks for torch's process group
    comm = None
    if dist is not None and not args.torch_collective:
        from mdsuite_b200.distributed import raw_communicator
        comm = raw_communicator(local_rank)

    # the same fixed workload on every rank; rank g takes block g of np.array_split (strong scaling)
    sysm, cutoff, nbins, desc = make_workload(args.workload, args.frames)
    n_total, n_atoms = sysm.n_frames, sysm.n_atoms
    mine = np.array_split(np.arange(n_total), world)[rank]
    n_frames = len(mine)
    host = torch.from_numpy(sysm.positions[mine[0]:mine[-1] + 1] if n_frames else
                            sysm.positions[:0]).pin_memory()
    dev = host.cuda(non_blocking=False)
    handle = _cuda.RDFHandle(sysm.counts, nbins, cutoff, sysm.box, parity=True, device=local_rank)
    n_pairs = handle.n_pairs
    counts_dev = counts_tensor(handle, n_pairs, nbins) if world > 1 else None
    pairs_per_step = sysm.pairs_per_frame(parity=True) * n_total  # the whole job, all ranks

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def collective():
        """ONE allreduce of the [n_pairs][nbins] uint64 accumulator, ordered after the kernels."""
        if dist is None:
            return
        if comm is not None:
            handle.allreduce(comm)  # on the library's stream
        else:
            handle.join()
            dist.all_reduce(counts_dev)
            handle.wait()

    def step_resident():
        if dist is not None and comm is None:
            handle.wait()  # the previous step's allreduce reads the accumulator
        handle.reset()
        if n_frames:
            handle.submit_device(dev)
        collective()
        handle.join()

    def step_e2e():
        if dist is not None and comm is None:
            handle.wait()
        handle.reset()
        if n_frames:
            handle.submit(host, pinned=True)
        collective()
        return handle.counts()

    # ---- this rank's own statistics: one un-reduced pass (after a step the accumulator holds
    # the sum over all ranks) --------------------------------------------------------------------
    handle.reset()
    if n_frames:
        handle.submit_device(dev)
    st_local = handle.stats()

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
    # raw + packed frames of this rank per step against L2 (strong scaling: from N = 4 on a rank's
    # share of cfg2 would fit)
    timer = TimedSteps(torch.device("cuda", local_rank), 24.0 * n_atoms * n_frames)
    t_wall0 = time.perf_counter()
    ms = timer.run(step_resident, args.steps)
    barrier()
    t_wall1 = time.perf_counter()
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    st = handle.stats()
    result = handle.counts()
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = pairs_per_step * args.steps / (ms * 1e-3)

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
    e2e_value = pairs_per_step * args.steps / e2e_s
    if not np.array_equal(result, result_e2e):
        raise SystemExit("bench.py: HBM-resident and end-to-end histograms differ")

    peaks, peaks_src = measured_peaks()
    f_max = float(peaks.get("sm_max_mhz", 1965.0)) * 1e6
    fp32_peak = st.sm_count * 128 * f_max  # non-FMA FP32 issue, op/s

    # ---- the north-star target configuration at full size (every rank takes part) -----------
    target = None
    if not args.no_target_config:
        try:
            target = target_config_probe(args, world, rank, local_rank, dist, comm, fp32_peak)
        except Exception as exc:  # noqa: BLE001
            if dist is not None:
                raise  # a rank that drops out would leave the others in the collective
            target = {"error": str(exc)}

    # ---- the other multi-GPU configurations BASELINE.json names: cfg4 at full size, the cfg5 sweep
    other = None
    if not args.no_other_configs:
        other = {}
        for key, fn in (("cfg4_full", lambda: target_config_probe(args, world, rank, local_rank, dist,
                                                                  comm, fp32_peak, name="cfg4")),
                        ("cfg5_sweep", lambda: cfg5_sweep_probe(args, world, rank, local_rank, dist,
                                                                comm, fp32_peak))):
            try:
                other[key] = fn()
            except Exception as exc:  # noqa: BLE001
                if dist is not None:
                    raise
                other[key] = {"error": str(exc)}

    # ---- multi-rank parity: the sharded + allreduced histogram against ONE handle on rank 0 over
    # all the frames, and the first frames against the fused-C oracle ---------------------------
    parity_multi = None
    if world > 1 and rank == 0:
        with _cuda.RDFHandle(sysm.counts, nbins, cutoff, sysm.box, parity=True, device=local_rank) as h:
            h.submit(sysm.positions)
            single = h.counts()
        parity_multi = {"frames_total": n_total, "ranks": world,
                        "bins_differing_vs_single_handle": int((single != result).sum()),
                        "bins_differing_e2e_vs_single_handle": int((single != result_e2e).sum())}
        if not args.no_cpu_baseline:
            n_c = max(1, min(n_total, args.cpu_frames))
            _r, _dt, c_fused, _th = cpu_fused_c_rate(sysm, cutoff, nbins, n_c)
            with _cuda.RDFHandle(sysm.counts, nbins, cutoff, sysm.box, parity=True,
                                 device=local_rank) as h:
                h.submit(sysm.positions[:n_c])
                got = h.counts().astype(np.int64)
            parity_multi["frames_checked_fused_c"] = n_c
            parity_multi["bins_differing_fused_c"] = int((got != c_fused).sum())

    if rank != 0:
        handle.close()
        if dist is not None:
            dist.barrier()
            from mdsuite_b200.distributed import close_raw_communicators
            close_raw_communicators()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (the all-pairs pair kernel), rank 0's launches -------
    sm_count = st.sm_count
    launches = max(1, st.pair_kernel_launches)
    avg_launch_s = st.pair_kernel_ms * 1e-3 / launches
    # pairs the pair kernel distance-tested per launch: every reference pair on the all-pairs
    # path, the candidates of nearby cells on the cell-list path — this rank's own
    pairs_per_launch = st_local.pairs_evaluated * args.steps / launches
    kernel_rate = pairs_per_launch / avg_launch_s
    kernel_name = "rdf_cells_kernel" if st.path == _cuda.PATH_CELLLIST else "rdf_allpairs_kernel"
    achieved = kernel_rate * FLOP_PER_PAIR
    kernel_share = st.pair_kernel_ms / ms
    atom = {}
    try:
        spread, rnd, mhz = _cuda.bench_shared_atomics(local_rank, max(64, n_pairs * (nbins + 1)), 4096)
        binned_frac = st_local.pairs_binned / max(st_local.pairs_evaluated, 1.0)
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
        "measured_on": "rank 0's launches (per GPU)",
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

    # ---- the literal reference-side binding of INTEGRATION.md §2 (rank 0, N = 1 only) --------
    binding = None
    if world == 1 and not args.no_binding_probe:
        try:
            binding = reference_binding_probe(sysm, cutoff, nbins, local_rank, e2e_value)
        except Exception as exc:  # noqa: BLE001
            binding = {"error": str(exc)}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ---------------------------------------
    cpu_baseline = None
    parity = parity_multi
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
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(desc, n_total, n_atoms, nbins, cutoff, n_pairs),
        "run_layout": {"frames_rank0": n_frames, "raw_mb_rank0": 12 * n_atoms * n_frames / 1e6,
                       "l2": timer.note,
                       "parallelism": f"the {n_total} configurations split over {world} GPU(s) with "
                                      f"np.array_split, 1 NCCL allreduce/step "
                                      f"({'torch.distributed' if comm is None else 'mds_rdf_allreduce(ncclComm_t)'})"
                       if world > 1 else "1 GPU"},
        "roofline": roofline, "target_config": target, "other_configs": other, "adf_path": adf,
        "user_api": user_api,
        "e2e_reference_binding": binding,
        "cpu_baseline": cpu_baseline,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": 1e3 * e2e_s / args.steps,
                "h2d_bytes_per_step": int(12 * n_atoms * n_total),
                "d2h_bytes_per_step": int(8 * n_pairs * nbins) * world},
        "gpu_launches": int(st.kernel_launches * args.steps),
        "gpu_launches_note": "rank 0's kernels in the timed region" if world > 1 else None,
        "clocks": clocks, "parity": parity, "host_binding_rank0": binding,
        "kernel_variant": {"path": st.path_name, "variant": st.variant, "dense": st.dense,
                           "cells": list(st.cells)},
    }
    emit(line)
    handle.close()
    if dist is not None:
        dist.barrier()
        from mdsuite_b200.distributed import close_raw_communicators
        close_raw_communicators()
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
    ap.add_argument("--frames", type=int, default=1000, help="configurations per step")
    ap.add_argument("--cpu-frames", type=int, default=12,
                    help="configurations timed by the op-by-op CPU baseline (x8 for the C port)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-user-api", action="store_true",
                    help="skip the Project -> Experiment -> run.RadialDistributionFunction timing")
    ap.add_argument("--no-adf-probe", action="store_true",
                    help="skip the short angular-distribution measurement added at N=1")
    ap.add_argument("--no-numa-binding", action="store_true",
                    help="leave the process on whatever CPUs it was started on")
    ap.add_argument("--torch-collective", action="store_true",
                    help="allreduce with torch.distributed instead of mds_rdf_allreduce(ncclComm_t)")
    ap.add_argument("--no-target-config", action="store_true",
                    help="skip the full-size north-star block (100k atoms x 10k configurations)")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the cfg4 full-size block and the cfg5 sweep")
    ap.add_argument("--target-steps", type=int, default=3,
                    help="timed steps of the full-size north-star block (at most --steps)")
    ap.add_argument("--no-binding-probe", action="store_true",
                    help="skip the INTEGRATION.md reference-side binding measurement added at N=1")
    args = ap.parse_args()
    if args.steps < 1:
        raise SystemExit("--steps must be >= 1")
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
