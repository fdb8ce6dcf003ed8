"""
AngularDistributionFunction — drop-in for the reference calculator
(mdsuite/calculators/angular_distribution_function.py:71-609) with its batch loop body
(``_build_histograms`` :374-405: neighbour list, float16 cutoff, triples, angles, weights, bin
sums) routed to the CUDA library behind include/mds_adf.h.  Everything around it follows the
reference: argument handling (:143-260), sampled configurations and species index ranges
(:262-301), the split of the configurations into batches (:536-575), per-batch density
normalisation and float32 accumulation (:392-404), the degree axis and ``max_peak`` (:407-441).

Reference behaviour kept on purpose:
  * a species combination "A-B-C" (combinations_with_replacement of the species in storage order)
    holds the angles at atoms of species A between neighbours of species B and C — angles around
    the LATER species with an EARLIER neighbour (e.g. Na-Cl-Na around Cl) are never produced;
  * the cutoff is compared in float16 (utils/neighbour_list.py:147-149);
  * every batch is density-normalised on its own and the batches are summed, so the curve's scale
    is the number of batches, and which configurations share a batch changes the weighting.  The
    reference derives the batch size from free host memory (quadratic scale function, factor 10);
    so does this class, with ``batches=`` as the same override.

Not supported (raises): ``molecules=True`` without molecule groups in the experiment.
"""
from __future__ import annotations

import itertools
import logging
import time
from dataclasses import dataclass
from typing import List, Union

import numpy as np

from mdsuite_b200.calculators.calculator import Calculator, call, resolve_species_data

log = logging.getLogger(__name__)

BIN_RANGE = (0.0, 3.15)  # "from 0 to a chemists pi" (reference :215)
# of the free host memory: MemoryManager takes it from the global config
# (memory_management/memory_manager.py:108, utils/config.py:47)
MEMORY_FRACTION = 0.5


@dataclass
class Args:
    """The saved properties (reference :52-68) — every field takes part in the cache lookup."""
    number_of_bins: int
    number_of_configurations: int
    correlation_time: int
    atom_selection: object
    data_range: int
    cutoff: float
    start: int
    stop: int
    norm_power: int
    species: list
    molecules: bool


def normalise_batch(raw: np.ndarray, nbins: int) -> np.ndarray:
    """np.histogram(..., weights=float32, density=True) from the summed weights of one batch
    (numpy/lib/_histograms_impl.py): float32 sums, divided by the float32 bin widths and the
    total; returned as the float32 the reference casts to (:398)."""
    edges = np.linspace(BIN_RANGE[0], BIN_RANGE[1], nbins + 1, dtype=np.float32)
    db = np.array(np.diff(edges), float)
    n = raw.astype(np.float32)
    with np.errstate(invalid="ignore", divide="ignore"):
        return (n / db / n.sum()).astype(np.float32)


class AngularDistributionFunction(Calculator):
    #: factory of the GPU engine (the C ABI binding); a seam for host-logic tests
    handle_factory = None

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.scale_function = {"quadratic": {"outer_scale_factor": 10}}
        self.loaded_property = "Positions"
        self.analysis_name = "Angular_Distribution_Function"
        self.x_label = r"$$\text{Angle} / \theta$$"
        self.y_label = r"$$\text{ADF} / a.u.$$"
        self.result_keys = ["max_peak"]
        self.result_series_keys = ["angle", "adf"]
        self.sample_configurations = None
        self.number_of_atoms = None
        self.adf_minibatch = None
        self.override_n_batches = None
        self.device = None
        self.run_stats = None
        self.raw = None  # (n_sampled, n_combinations, B) float64 weight sums of the last run

    @call
    def __call__(self, batch_size: int = 1, minibatch: int = -1, number_of_configurations: int = 5,
                 cutoff: float = 6.0, start: int = 1, stop: int = None, number_of_bins: int = 500,
                 species: list = None, use_tf_function: bool = False, molecules: bool = False,
                 atom_selection: Union[slice, dict] = np.s_[:], plot: bool = True,
                 norm_power: int = 4, **kwargs):
        """Reference :143-216 — only stores arguments; the ``@call`` decorator runs the rest."""
        self.args = Args(number_of_bins=number_of_bins, cutoff=cutoff, start=start, stop=stop,
                         atom_selection=atom_selection, data_range=1, correlation_time=1,
                         molecules=molecules, species=species,
                         number_of_configurations=number_of_configurations, norm_power=norm_power)
        self.plot = plot
        self.save_to_db = kwargs.pop("save", True)
        self.adf_minibatch = minibatch  # accepted for compatibility; the kernel tiles on its own
        self.override_n_batches = kwargs.get("batches")
        self.device = kwargs.pop("device", None)

    # ---- argument defaults (reference :218-301) ---------------------------------------------
    def check_input(self):
        if self.args.stop is None:
            self.args.stop = self.experiment.number_of_configurations - 1
        if self.args.species is None:
            self.args.species = list(self.experiment.molecules if self.args.molecules
                                     else self.experiment.species)

    def _species_counts(self) -> List[int]:
        counts = []
        for item in self.args.species:
            if isinstance(self.args.atom_selection, dict):
                counts.append(len(self.args.atom_selection[item]))
            elif item in self.experiment.species:
                counts.append(self.experiment.species[item].n_particles)
            else:
                counts.append(self.experiment.molecules[item].n_particles)
        return counts

    def _prepare_data_structure(self):
        """Sampled configurations and (name, first, last+1) index ranges (reference :262-301)."""
        sample_configs = np.linspace(self.args.start, self.args.stop,
                                     self.args.number_of_configurations, dtype=int)
        species_indices, start = [], 0
        for name, n in zip(self.args.species, self._species_counts()):
            species_indices.append((name, start, start + n))
            start += n
        return sample_configs, species_indices

    def _number_of_batches(self, paths) -> int:
        """Reference :536-563 on top of MemoryManager.get_batch_size with the quadratic scale
        function: per-configuration memory = 10 * (bytes per configuration)^2."""
        if self.override_n_batches is not None:
            return int(self.override_n_batches)
        import psutil
        per_configuration = 0.0
        n_configs = 1
        for p in paths:
            _n, n_configs, n_bytes = self.experiment.database.get_data_size(p)
            per_configuration += n_bytes / n_configs
        per_configuration = 10.0 * per_configuration ** 2
        batch_size = int(np.clip(MEMORY_FRACTION * psutil.virtual_memory().available
                                 / per_configuration, 1, n_configs))
        if batch_size > self.args.number_of_configurations:
            return 1
        return int(self.args.number_of_configurations / batch_size)

    def prepare_computation(self):
        paths, selections = resolve_species_data(self.experiment, self.args, self.loaded_property)
        n_batches = self._number_of_batches(paths)
        # the batch count follows from the host's free memory: under torch.distributed the ranks
        # could round it differently, and every batch is normalised on its own — rank 0 decides
        from mdsuite_b200.distributed import distributed_context
        _rank, _world, dist = distributed_context()
        if dist is not None:
            box = [int(n_batches)]
            dist.broadcast_object_list(box, src=0)
            n_batches = int(box[0])
        split_arr = np.array_split(self.sample_configurations, n_batches)
        return paths, selections, split_arr

    def _load(self, paths, selections, frames, counts, staging=None) -> np.ndarray:
        """(len(frames), N, 3) float32, species concatenated in args.species order (:459-487);
        written into ``staging`` (page-locked, reused by every batch) when given."""
        if staging is not None:
            out = staging[:len(frames)]
        else:
            out = np.empty((len(frames), sum(counts), 3), dtype=np.float32)
        off = 0
        for p, sel, n in zip(paths, selections, counts):
            out[:, off:off + n] = self.experiment.database.load_frames(p, frames, atom_selection=sel)
            off += n
        return out

    # ---- run (reference :577-609) ------------------------------------------------------------
    def run_calculator(self):
        self.check_input()
        self.sample_configurations, species_indices = self._prepare_data_structure()
        paths, selections, split_arr = self.prepare_computation()
        counts = [hi - lo for _, lo, hi in species_indices]
        self.number_of_atoms = sum(counts)
        injected = type(self).handle_factory
        if injected is None:
            import torch
            if not torch.cuda.is_available():
                raise RuntimeError("AngularDistributionFunction: no CUDA device; this build has no "
                                   "CPU path")
            from mdsuite_b200._cuda import adf as cuda_adf
            device = self.device if self.device is not None else torch.cuda.current_device()
            factory = lambda *a, **k: cuda_adf.acquire_handle(*a, device=device, **k)  # noqa: E731
        else:
            factory = injected
        from mdsuite_b200.distributed import (allreduce_host_sums, distributed_context,
                                              shard_configurations)
        nbins = self.args.number_of_bins
        combos = list(itertools.combinations_with_replacement(range(len(counts)), 3))
        rank, world, dist = distributed_context()
        # Configurations are independent up to the per-batch normalisation: every rank takes its
        # block of every batch, and ONE allreduce of the per-batch weight sums
        # [batch][combination][bin] (float64) makes the batches whole on all ranks.
        batch_sums = np.zeros((len(split_arr), len(combos), nbins), dtype=np.float64)
        raw_all = []
        t0 = time.perf_counter()
        handle = factory(counts, nbins, self.args.cutoff, self.experiment.box_array,
                         norm_power=self.args.norm_power)
        # a handle taken back from the previous call keeps counting: statistics are differences
        before = handle.stats() if hasattr(handle, "stats") else None
        # Two buffers for the largest batch (page-locked for long runs): the next batch is read from
        # the store by a worker thread while the GPU works on the current one, and no batch
        # allocates.
        my_batches = [(b, shard_configurations(frames, world, rank) if dist is not None else frames)
                      for b, frames in enumerate(split_arr)]
        my_batches = [(b, f) for b, f in my_batches if len(f)]
        staging, ring = [None, None], None
        if injected is None and my_batches:
            from mdsuite_b200.memory_management.memory_manager import _acquire_ring
            largest = max(len(f) for _, f in my_batches)
            # two page-locked buffers from the per-process ring cache (page-locking costs ~0.5 ms
            # per MB the first time; the next call of the process gets the same buffers back)
            pin = largest * sum(counts) * 12 <= (1 << 30)
            ring = _acquire_ring((largest, sum(counts), 3), pin)
            staging = [t.numpy() for t in ring[1]]
        failed = True
        try:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=1) as pool:
                def fetch(k):
                    return pool.submit(self._load, paths, selections, my_batches[k][1], counts,
                                       staging[k % 2])
                pending = fetch(0) if my_batches else None
                for k, (b, mine) in enumerate(my_batches):
                    positions = pending.result()
                    pending = fetch(k + 1) if k + 1 < len(my_batches) else None
                    raw = handle.compute(positions)  # (frames, combinations, bins) float64
                    raw_all.append(raw)
                    batch_sums[b] = raw.sum(axis=0)
            stats = handle.stats() if hasattr(handle, "stats") else None
            failed = False
        finally:
            if injected is None and not failed:
                cuda_adf.park_handle(handle)  # kept for the next call with this configuration
            else:
                handle.close()
            if ring is not None:
                from mdsuite_b200.memory_management.memory_manager import _release_ring
                _release_ring(*ring)
        if dist is not None:
            batch_sums = allreduce_host_sums(batch_sums, device=getattr(handle, "device", None))
        angles = [None] * len(combos)
        for b, frames in enumerate(split_arr):
            if len(frames) == 0:
                continue
            for c in range(len(combos)):
                hist = normalise_batch(batch_sums[b, c], nbins)
                angles[c] = hist if angles[c] is None else angles[c] + hist
        self.raw = np.concatenate(raw_all, axis=0) if raw_all else None  # this rank's share
        elapsed = time.perf_counter() - t0
        if stats is not None:
            triples = stats.triples - (before.triples if before is not None else 0.0)
            pair_tests = stats.pair_tests - (before.pair_tests if before is not None else 0.0)
            self.run_stats = {"triples_this_rank": triples,
                              "pair_tests_this_rank": pair_tests, "seconds": elapsed,
                              "max_neighbours": stats.max_neighbours,
                              "frames": int(len(self.sample_configurations)),
                              "batches": len(split_arr), "world_size": world}
            log.info(f"ADF: {triples:.3e} triples on this rank in {elapsed:.3f} s")
        self._compute_adfs(angles, combos)

    def _compute_adfs(self, angles, combos):
        """Degree axis, max_peak, one record per species combination (reference :407-441)."""
        nbins = self.args.number_of_bins
        axis = np.linspace(BIN_RANGE[0] * (180 / 3.14159), BIN_RANGE[1] * (180 / 3.14159), nbins)
        self.data_range = self.args.number_of_configurations
        for c, combo in enumerate(combos):
            names = [self.args.species[s] for s in combo]
            hist = angles[c] if angles[c] is not None else np.full(nbins, np.nan, np.float32)
            data = {self.result_keys[0]: float(axis[int(np.argmax(hist))]),
                    self.result_series_keys[0]: axis.tolist(),
                    self.result_series_keys[1]: hist.tolist()}
            self.queue_data(data=data, subjects=names)

    def plot_data(self, data: dict):
        for selected_species, val in data.items():
            axis = np.linspace(BIN_RANGE[0] * (180 / 3.14159), BIN_RANGE[1] * (180 / 3.14159),
                               len(val[self.result_series_keys[0]]))
            title_value = axis[int(np.argmax(val[self.result_series_keys[1]]))]
            self.run_visualization(x_data=np.array(val[self.result_series_keys[0]]),
                                   y_data=np.array(val[self.result_series_keys[1]]),
                                   title=f"{selected_species} - Max: {title_value:.3f} degrees ")
