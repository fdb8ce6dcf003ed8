"""
TEST INFRASTRUCTURE ONLY — pure-Python restatement of the reference's extended XYZ reader
(reference mdsuite/file_io/extxyz_files.py:55-296 and tabular_text_files.py:122-220), the checker
for the native-backed mdsuite_b200.file_io.extxyz_files.EXTXYZFile on files too large to commit as
golden vectors.  Never imported by the product.

Pinned: tests/test_reader_golden.py holds this restatement to the output of the reference's own
EXTXYZFile (tests/golden/reader_golden.json, made by tests/golden/make_reader_golden.py).

Line-by-line reading with str.split() and NumPy's str -> float64 conversion, as the reference does;
the header conventions (Lattice entries 0/4/8, Properties column counting over known names only,
time= difference) are restated independently of the product's helpers.
"""
from __future__ import annotations

import pathlib
from typing import Dict, Iterator, List

import numpy as np

from mdsuite_b200.database.simulation_database import (PropertyInfo, SpeciesInfo,
                                                       TrajectoryChunkData, TrajectoryMetadata)
from mdsuite_b200.file_io.file_read import FileProcessor

# reference mdsuite/file_io/extxyz_files.py:44-52
KNOWN = {"pos": "Positions", "vel": "Velocities", "force": "Forces", "stress": "Stress",
         "energies": "Energy", "time": "Time", "momenta": "Momenta"}


def _lattice(header: str) -> List[float]:
    tokens = header.split()
    first = [i for i, t in enumerate(tokens) if "Lattice" in t][0]
    last = [i for i in range(first, len(tokens)) if tokens[i].endswith('"')][0]
    numbers = " ".join(tokens[first:last + 1]).split("=", 1)[1].replace('"', "").split()
    return [float(numbers[0]), float(numbers[4]), float(numbers[8])]


def _time(header: str):
    found = None
    for t in header.split():
        if "time" in t:
            tail = t.split("=")[-1]
            try:
                found = float(tail)
            except ValueError:
                found = float(tail.split(",")[0])
    return found


def _columns(header: str, known: Dict[str, str]):
    token = [t for t in header.split() if "Properties" in t][-1]
    fields = token.split("=")[1].replace(":S", "").replace(":R", "").split(":")
    species_col, props, col = None, {}, 0
    for k, f in enumerate(fields):
        if f == "species":
            species_col, col = col, col + 1
        if f in known:
            width = int(fields[k + 1])
            props[known[f]] = list(range(col, col + width))
            col += width
    if species_col is None:
        raise RuntimeError("could not find species in header")
    return species_col, props


class EXTXYZFileOracle(FileProcessor):
    def __init__(self, file_path, custom_data_map: dict = None, chunk_configurations: int = 64):
        self.file_path = str(pathlib.Path(file_path))
        self.known = dict(KNOWN)
        for prop, name in (custom_data_map or {}).items():
            self.known[name] = prop
        self.chunk_configurations = chunk_configurations
        self._mdata = None
        self._layout = None

    def __str__(self):
        return self.file_path

    def _analyse(self):
        if self._layout is not None:
            return self._layout
        with open(self.file_path, "r") as fh:
            n_particles = int(fh.readline())
            header = fh.readline()
            species_col, props = _columns(header, self.known)
            first = [fh.readline().split() for _ in range(n_particles)]
            fh.readline()
            header_1 = fh.readline()
            fh.seek(0)
            n_lines = sum(1 for _ in fh)
        n_configs_f = n_lines / (n_particles + 2)
        n_configs = int(round(n_configs_f))
        if abs(n_configs_f - n_configs) > 1e-10:
            raise ValueError("file length is not a whole number of configurations")
        species: Dict[str, List[int]] = {}
        for i, row in enumerate(first):
            species.setdefault(row[species_col], []).append(i)
        t0, t1 = _time(header), _time(header_1)
        sample_rate = int(round(t1 - t0)) if t0 is not None and t1 is not None else None
        self._layout = dict(n_particles=n_particles, n_configs=n_configs, species=species,
                            props=props, box_l=_lattice(header), sample_rate=sample_rate)
        return self._layout

    def _get_metadata(self) -> TrajectoryMetadata:
        lay = self._analyse()
        species_list = [SpeciesInfo(name=n, n_particles=len(idx),
                                    properties=[PropertyInfo(p, len(c)) for p, c in lay["props"].items()])
                        for n, idx in lay["species"].items()]
        return TrajectoryMetadata(n_configurations=lay["n_configs"], species_list=species_list,
                                  box_l=lay["box_l"], sample_rate=lay["sample_rate"])

    def get_configurations_generator(self) -> Iterator[TrajectoryChunkData]:
        lay = self._analyse()
        n, species_list = lay["n_particles"], self.metadata.species_list
        with open(self.file_path, "r") as fh:
            done = 0
            while done < lay["n_configs"]:
                chunk_n = min(self.chunk_configurations, lay["n_configs"] - done)
                chunk = TrajectoryChunkData(species_list, chunk_n)
                for c in range(chunk_n):
                    fh.readline()
                    fh.readline()
                    table = np.array([fh.readline().split() for _ in range(n)])
                    for prop, cols in lay["props"].items():
                        values = table[:, cols].astype(np.float64)
                        for name, idx in lay["species"].items():
                            chunk.add_data(values[idx][None], c, name, prop)
                done += chunk_n
                yield chunk
