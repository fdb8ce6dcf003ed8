"""
TEST INFRASTRUCTURE ONLY — pure-Python restatement of the reference's LAMMPS text dump reader
(reference mdsuite/file_io/lammps_trajectory_files.py:69-298 and tabular_text_files.py:122-220),
the checker for the native reader behind include/mds_io.h (mdsuite_b200/_native).  Never imported
by the product.

Pinned: tests/test_reader_golden.py holds this restatement to the output of the reference's own
LAMMPSTrajectoryFile (tests/golden/reader_golden.json, made by tests/golden/make_reader_golden.py,
which executes the reference's reader modules unmodified); tests/golden/lammps_golden.json is an
additional hand-computed vector.

Same conventions as the reference: 9 header lines per configuration, box lengths from header
lines 5-7, the ``id`` column sorts the atoms of every configuration, species come from the
``element`` column (else ``type``) of the FIRST configuration in order of first appearance, and
``x y z`` become the ``Positions`` property.  Only the properties the RDF loads are extracted.
"""
from __future__ import annotations

import pathlib
from typing import Dict, Iterator, List

import numpy as np

from mdsuite_b200.database.simulation_database import (PropertyInfo, SpeciesInfo,
                                                       TrajectoryChunkData, TrajectoryMetadata)
from mdsuite_b200.file_io.file_read import FileProcessor

N_HEADER_LINES = 9
# reference mdsuite/file_io/lammps_trajectory_files.py:39-64
COLUMN_NAMES = {
    "Positions": ["x", "y", "z"], "Scaled_Positions": ["xs", "ys", "zs"],
    "Unwrapped_Positions": ["xu", "yu", "zu"], "Scaled_Unwrapped_Positions": ["xsu", "ysu", "zsu"],
    "Velocities": ["vx", "vy", "vz"], "Forces": ["fx", "fy", "fz"], "Box_Images": ["ix", "iy", "iz"],
    "Dipole_Orientation_Magnitude": ["mux", "muy", "muz"],
    "Angular_Velocity_Spherical": ["omegax", "omegay", "omegaz"],
    "Angular_Velocity_Non_Spherical": ["angmomx", "angmomy", "angmomz"],
    "Torque": ["tqx", "tqy", "tqz"], "Charge": ["q"], "Kinetic_Energy": ["c_KE"],
    "Potential_Energy": ["c_PE"],
    "Stress": ["c_Stress[%d]" % k for k in range(1, 7)],
}


class LAMMPSTrajectoryFileOracle(FileProcessor):
    def __init__(self, file_path, trajectory_is_sorted_by_ids: bool = False,
                 chunk_configurations: int = 64, custom_data_map: dict = None):
        self.file_path = str(pathlib.Path(file_path))
        self.trajectory_is_sorted_by_ids = trajectory_is_sorted_by_ids
        self.chunk_configurations = chunk_configurations
        self.column_names = dict(COLUMN_NAMES)
        self.column_names.update({k: list(v) for k, v in (custom_data_map or {}).items()})
        self._mdata = None
        self._layout = None

    def __str__(self):
        return self.file_path

    # ---- header analysis ------------------------------------------------------------------
    def _analyse(self):
        if self._layout is not None:
            return self._layout
        with open(self.file_path, "r") as fh:
            header = [fh.readline() for _ in range(N_HEADER_LINES)]
            n_particles = int(header[3].split()[0])
            columns = header[8].split()[2:]
            if "id" not in columns:
                raise ValueError("LAMMPS dump needs an 'id' column")
            if "element" in columns:
                species_col = columns.index("element")
            elif "type" in columns:
                species_col = columns.index("type")
            else:
                raise ValueError("Insufficient species or type identification available.")
            box_l = [float(line.split()[1]) - float(line.split()[0]) for line in header[5:8]]
            first = np.array([fh.readline().split() for _ in range(n_particles)])
            step0 = int(header[1])
            second = [fh.readline() for _ in range(N_HEADER_LINES)]
            sample_rate = int(second[1]) - step0 if second[1].strip() else None
            fh.seek(0)
            n_lines = sum(1 for _ in fh)
        n_configs_f = n_lines / (n_particles + N_HEADER_LINES)
        n_configs = int(round(n_configs_f))
        if abs(n_configs_f - n_configs) > 1e-10:
            raise ValueError("file length is not a whole number of configurations")
        id_col = columns.index("id")
        if not self.trajectory_is_sorted_by_ids:
            first = first[np.argsort(first[:, id_col].astype(np.int64), kind="stable")]
        species: Dict[str, List[int]] = {}
        for i, name in enumerate(first[:, species_col]):
            species.setdefault(str(name), []).append(i)
        props = {}
        for prop, names in self.column_names.items():
            if all(n in columns for n in names):
                props[prop] = [columns.index(n) for n in names]
        self._layout = dict(n_particles=n_particles, n_configs=n_configs, columns=columns,
                            id_col=id_col, species=species, props=props, box_l=box_l,
                            sample_rate=sample_rate)
        return self._layout

    def _get_metadata(self) -> TrajectoryMetadata:
        lay = self._analyse()
        species_list = [SpeciesInfo(name=n, n_particles=len(idx),
                                    properties=[PropertyInfo(p, len(c)) for p, c in lay["props"].items()])
                        for n, idx in lay["species"].items()]
        return TrajectoryMetadata(n_configurations=lay["n_configs"], species_list=species_list,
                                  box_l=lay["box_l"], sample_rate=lay["sample_rate"])

    # ---- streaming --------------------------------------------------------------------------
    def get_configurations_generator(self) -> Iterator[TrajectoryChunkData]:
        lay = self._analyse()
        n, species_list = lay["n_particles"], self.metadata.species_list
        n_cols = len(lay["columns"])
        with open(self.file_path, "r") as fh:
            done = 0
            while done < lay["n_configs"]:
                chunk_n = min(self.chunk_configurations, lay["n_configs"] - done)
                chunk = TrajectoryChunkData(species_list, chunk_n)
                for c in range(chunk_n):
                    for _ in range(N_HEADER_LINES):
                        fh.readline()
                    tokens = np.array("".join(fh.readline() for _ in range(n)).split())
                    table = tokens.reshape(n, n_cols)
                    if not self.trajectory_is_sorted_by_ids:
                        order = np.argsort(table[:, lay["id_col"]].astype(np.int64), kind="stable")
                        table = table[order]
                    for prop, cols in lay["props"].items():
                        values = table[:, cols].astype(np.float64)
                        for name, idx in lay["species"].items():
                            chunk.add_data(values[idx][None], c, name, prop)
                done += chunk_n
                yield chunk
