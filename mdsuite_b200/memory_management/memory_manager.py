"""
Batch sizing and the frame batcher that feeds the GPU.

The reference's MemoryManager picks configurations-per-batch from free host RAM and a per-calculator
scale function (mdsuite/memory_management/memory_manager.py:179-219); for the RDF the quadratic
function (calculators/radial_distribution_function.py:119-121) collapses that to ONE configuration
per batch for anything above ~1 000 atoms, because the TensorFlow path materialises
O(m*N) pair tensors.  The CUDA path materialises nothing per pair: a configuration costs 12 B/atom
of staging plus 12 B/atom packed (up to 44 B/atom on the cell-list path), so the scale function is linear and batches hold hundreds of
frames.  ``get_batch_size`` keeps the reference's (batch_size, n_batches, remainder) contract.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from mdsuite_b200.utils.config import config

# bytes held per configuration byte on disk: pinned host ring (2 x 1) + device staging (2 x 1) +
# packed quad-block frames (2 x 1); the cell-list path adds sorted atoms and ranks (2 x 20/12)
GPU_LINEAR_SCALE = 2 + 2 + 2 + 2 * 20 / 12


def linear_scale_function(memory_usage: float, scale_factor: float = 1.0) -> float:
    """mdsuite/utils/scale_functions.py:30-47."""
    return memory_usage * scale_factor


class MemoryManager:
    def __init__(self, data_path: Sequence[str] = None, database=None, memory_fraction: float = None,
                 scale_function: dict = None, device: Optional[int] = None, offset: int = 0,
                 free_memory: Optional[int] = None):
        self.data_path = list(data_path) if data_path is not None else None
        self.database = database
        self.memory_fraction = config.memory_fraction if memory_fraction is None else memory_fraction
        self.scale_function_parameters = (scale_function or {"linear": {"scale_factor": GPU_LINEAR_SCALE}})
        self.offset = offset
        self.device = device
        self._free_memory = free_memory
        self.batch_size = None
        self.n_batches = None
        self.remainder = None

    @property
    def machine_properties(self) -> dict:
        """{'memory': bytes} — free HBM of the target GPU (free host RAM in the reference,
        utils/meta_functions.py:132-158)."""
        if self._free_memory is not None:
            return {"memory": self._free_memory}
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("MemoryManager: no CUDA device (the RDF path has no CPU fallback)")
        free, _total = torch.cuda.mem_get_info(self.device if self.device is not None else 0)
        return {"memory": free}

    def get_batch_size(self) -> tuple:
        if self.data_path is None:
            raise ValueError("No tensor_values have been requested.")
        per_configuration_memory = 0.0
        n_configs = 0
        for item in self.data_path:
            _n_particles, n_configs, n_bytes = self.database.get_data_size(item)
            per_configuration_memory += n_bytes / max(n_configs, 1)
        name, params = next(iter(self.scale_function_parameters.items()))
        if name != "linear":
            raise KeyError("Invalid choice")
        per_configuration_memory = linear_scale_function(per_configuration_memory, **params)
        max_loaded = int(np.clip(self.memory_fraction * self.machine_properties["memory"]
                                 / max(per_configuration_memory, 1.0), 1,
                                 max(n_configs - self.offset, 1)))
        batch_size = max_loaded
        n_batches, remainder = divmod(n_configs - self.offset, batch_size)
        self.batch_size, self.n_batches, self.remainder = batch_size, n_batches, remainder
        return batch_size, n_batches, remainder


_RING_CACHE = {}  # (shape, pinned) -> list of idle rings


def _acquire_ring(shape, pin: bool):
    """Two staging buffers of ``shape``.  Page-locking costs ~0.2 ms/MB, so rings are returned to
    a small per-process cache when a batcher is closed and reused by the next calculator call."""
    import torch
    key = (tuple(shape), bool(pin))
    idle = _RING_CACHE.get(key)
    if idle:
        return key, idle.pop()
    return key, [torch.empty(shape, dtype=torch.float32, pin_memory=pin) for _ in range(2)]


def _release_ring(key, ring):
    for other in [k for k in _RING_CACHE if k != key]:  # keep one shape: bounded pinned memory
        del _RING_CACHE[other]
    idle = _RING_CACHE.setdefault(key, [])
    if len(idle) < 8:
        idle.append(ring)


class FrameBatcher:
    """Streams configurations from the trajectory store into a ring of two pinned host buffers and
    submits them to a ``RDFHandle``; the library copies them H2D on its side stream while the
    previous batch's kernels run, so disk reads, copies and compute overlap
    (replaces DataManager.batch_generator + tf.data prefetch, database/data_manager.py:118-286,
    calculators/trajectory_calculator.py:309-363)."""

    def __init__(self, database, paths: Sequence[str], n_particles: Sequence[int], handle,
                 frames_per_batch: int, atom_selections: Optional[List] = None, pin: bool = True):
        import torch
        self.database = database
        self.paths = list(paths)
        self.n_particles = [int(n) for n in n_particles]
        self.handle = handle
        self.frames_per_batch = max(1, int(frames_per_batch))
        self.atom_selections = atom_selections or [None] * len(self.paths)
        n_atoms = sum(self.n_particles)
        self._ring_key, self._ring = _acquire_ring((self.frames_per_batch, n_atoms, 3), pin)
        self.pinned = pin
        self.frames_submitted = 0
        self.bytes_staged = 0

    def close(self):
        """Hand the staging buffers back (call after the handle has been synchronised)."""
        if self._ring is not None:
            _release_ring(self._ring_key, self._ring)
            self._ring = None

    def batches(self, sample_configurations: Sequence[int]):
        """Generator: every step loads and submits one batch of the given configuration indices
        (in order) and yields the number of frames submitted so far.  Several batchers — one per
        device of an in-process multi-GPU run — are fed in turn by stepping their generators
        round-robin, so that no device waits for another device's whole shard."""
        frames = np.asarray(sample_configurations, dtype=np.int64)
        # Ramp: nothing can overlap the first load, so the first two batches are small (1/8 and
        # 1/3 of a full one) and the GPU starts within a fraction of a full batch's load time.
        # Batches that must stay whole (``align``, e.g. HDF5 chunks decoded in place) are not cut.
        fpb, align = self.frames_per_batch, max(1, int(getattr(self, "align", 1)))
        starts, lo, k = [], 0, 0
        while lo < len(frames):
            size = fpb
            if getattr(self, "ramp", False) and len(frames) > 2 * fpb and k < 2:
                size = max(align, (fpb // (8 if k == 0 else 3)) // align * align)
            starts.append((lo, min(len(frames), lo + size)))
            lo += size
            k += 1
        tickets = [None, None]  # the submit that last read each ring buffer
        for k, (lo, hi) in enumerate(starts):
            idx = frames[lo:hi]
            buf = self._ring[k & 1]
            if k >= 2:
                # the buffer's previous H2D copy has finished — the copy of the OTHER buffer,
                # submitted a moment ago, keeps running while this one is refilled
                if tickets[k & 1] is not None and hasattr(self.handle, "wait_host_ticket"):
                    self.handle.wait_host_ticket(tickets[k & 1])
                else:
                    self.handle.wait_host_copies()
            view = buf.numpy()
            off = 0
            for path, n, sel in zip(self.paths, self.n_particles, self.atom_selections):
                self.database.load_frames(path, idx, out=view[:len(idx), off:off + n],
                                          atom_selection=sel)
                off += n
            self.handle.submit(buf[:len(idx)], pinned=self.pinned)
            if self.pinned and hasattr(self.handle, "host_ticket"):
                tickets[k & 1] = self.handle.host_ticket()
            self.frames_submitted += len(idx)
            self.bytes_staged += len(idx) * off * 12
            yield self.frames_submitted

    def run(self, sample_configurations: Sequence[int]) -> int:
        """Submit the given configuration indices (in order); returns the number of frames."""
        for _ in self.batches(sample_configurations):
            pass
        return self.frames_submitted
