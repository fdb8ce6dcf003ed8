"""
Shared part of the native-backed text readers — the counterpart of the reference's
``TabularTextFileProcessor`` (mdsuite/file_io/tabular_text_files.py:32-220): a subclass analyses
its format's header into a layout (atoms, configurations, species -> line indices, property ->
columns, box, sample rate) and names the native handle; metadata and the chunked stream into the
trajectory store are the same for every format.
"""
from __future__ import annotations

from typing import Iterator

import numpy as np

from mdsuite_b200.database.simulation_database import (PropertyInfo, SpeciesInfo,
                                                       TrajectoryChunkData, TrajectoryMetadata)
from mdsuite_b200.file_io.file_read import FileProcessor


class NativeTabularTextFile(FileProcessor):
    """Subclasses set ``file_path`` / ``chunk_bytes`` and implement ``_open`` (the native handle),
    ``_analyse`` (the layout dict: n_particles, n_configs, species, props, box_l, sample_rate) and
    ``_sort_by_id`` (whether the rows of a configuration are ordered by the id column)."""

    chunk_bytes = 256 << 20
    _mdata = None

    def __str__(self):
        return self.file_path

    def _open(self):
        raise NotImplementedError

    def _analyse(self) -> dict:
        raise NotImplementedError

    def _sort_by_id(self) -> bool:
        return False

    def _get_metadata(self) -> TrajectoryMetadata:
        lay = self._analyse()
        species_list = [SpeciesInfo(name=n, n_particles=len(idx),
                                    properties=[PropertyInfo(p, len(c)) for p, c in lay["props"].items()])
                        for n, idx in lay["species"].items()]
        return TrajectoryMetadata(n_configurations=lay["n_configs"], species_list=species_list,
                                  box_l=lay["box_l"], sample_rate=lay["sample_rate"])

    def get_configurations_generator(self) -> Iterator[TrajectoryChunkData]:
        """Reference tabular_text_files.py:122-220: chunks of configurations, every property of
        every species as (configurations, particles, dimensions)."""
        lay = self._analyse()
        dump = self._open()
        n, species_list = lay["n_particles"], self.metadata.species_list
        prop_items = list(lay["props"].items())
        all_cols = [c for _, cols in prop_items for c in cols]
        if not all_cols:
            return
        per_frame = max(1, n * len(all_cols) * 4)
        chunk_frames = max(1, min(lay["n_configs"], self.chunk_bytes // per_frame))
        # the native reader scatters the atom of rank k (k-th id, or k-th line) to row dest[k]:
        # species-contiguous, so the per-species slices below are plain views
        dest = np.empty(n, dtype=np.int32)
        bounds, off = {}, 0
        for name, idx in lay["species"].items():
            dest[np.asarray(idx, dtype=np.int64)] = np.arange(off, off + len(idx), dtype=np.int32)
            bounds[name] = (off, off + len(idx))
            off += len(idx)
        done = 0
        while done < lay["n_configs"]:
            k = min(chunk_frames, lay["n_configs"] - done)
            table = dump.read_columns(done, k, all_cols, dest=dest, sort_by_id=self._sort_by_id())
            chunk = TrajectoryChunkData(species_list, k)
            off = 0
            for prop, cols in prop_items:
                values = table[:, :, off:off + len(cols)]
                off += len(cols)
                for name, (lo, hi) in bounds.items():
                    chunk.add_data(values[:, lo:hi], 0, name, prop)
            done += k
            yield chunk
