"""
LAMMPS text dump reader — drop-in for the reference's ``LAMMPSTrajectoryFile``
(mdsuite/file_io/lammps_trajectory_files.py:69-298, tabular_text_files.py:122-220) on top of the
native reader (include/mds_io.h, mdsuite_b200/_native): the file is mmap'ed, its configurations are
indexed and parsed by all host threads, and positions arrive as float32(float64(token)) exactly as
the reference's NumPy conversion + float32 store would produce them.

Same conventions as the reference: 9 header lines per configuration, box lengths from header
lines 6-8, the ``id`` column orders the atoms of every configuration, species come from the
``element`` column (else ``type``) of the FIRST configuration in order of first appearance, and
``x y z`` become the ``Positions`` property.
"""
from __future__ import annotations

import pathlib
from typing import Dict, List

from mdsuite_b200.file_io.tabular_text_files import NativeTabularTextFile

# dump column names of every per-atom property the reference recognises
# (mdsuite/file_io/lammps_trajectory_files.py:39-64; names from database/mdsuite_properties.py)
COLUMN_NAMES = {
    "Positions": ["x", "y", "z"],
    "Scaled_Positions": ["xs", "ys", "zs"],
    "Unwrapped_Positions": ["xu", "yu", "zu"],
    "Scaled_Unwrapped_Positions": ["xsu", "ysu", "zsu"],
    "Velocities": ["vx", "vy", "vz"],
    "Forces": ["fx", "fy", "fz"],
    "Box_Images": ["ix", "iy", "iz"],
    "Dipole_Orientation_Magnitude": ["mux", "muy", "muz"],
    "Angular_Velocity_Spherical": ["omegax", "omegay", "omegaz"],
    "Angular_Velocity_Non_Spherical": ["angmomx", "angmomy", "angmomz"],
    "Torque": ["tqx", "tqy", "tqz"],
    "Charge": ["q"],
    "Kinetic_Energy": ["c_KE"],
    "Potential_Energy": ["c_PE"],
    "Stress": ["c_Stress[1]", "c_Stress[2]", "c_Stress[3]", "c_Stress[4]", "c_Stress[5]",
               "c_Stress[6]"],
}


class LAMMPSTrajectoryFile(NativeTabularTextFile):
    def __init__(self, file_path, trajectory_is_sorted_by_ids: bool = False,
                 custom_data_map: dict = None, chunk_bytes: int = 256 << 20, n_threads: int = 0):
        """``custom_data_map`` maps further property names to dump columns, e.g.
        {"Reduced_Momentum": ["rp_x", "rp_y", "rp_z"]} (reference :80-90)."""
        self.file_path = str(pathlib.Path(file_path))
        self.trajectory_is_sorted_by_ids = trajectory_is_sorted_by_ids
        self.column_names = dict(COLUMN_NAMES)
        self.column_names.update(custom_data_map or {})
        self.chunk_bytes = chunk_bytes
        self.n_threads = n_threads
        self._mdata = None
        self._layout = None
        self._dump = None

    def _open(self):
        if self._dump is None:
            from mdsuite_b200 import _native
            self._dump = _native.LammpsDump(self.file_path, n_threads=self.n_threads)
        return self._dump

    # ---- header analysis (reference lammps_trajectory_files.py:103-179) ---------------------
    def _analyse(self):
        if self._layout is not None:
            return self._layout
        dump = self._open()
        info = dump.info
        columns = info.columns
        if info.id_column < 0:
            raise ValueError("LAMMPS dump needs an 'id' column")
        if info.species_column < 0:
            raise ValueError("Insufficient species or type identification available.")
        labels = dump.first_frame_labels(info.species_column,
                                         sort_by_id=not self.trajectory_is_sorted_by_ids)
        species: Dict[str, List[int]] = {}
        for i, name in enumerate(labels):
            species.setdefault(name, []).append(i)
        props = {}
        for prop, names in self.column_names.items():
            if all(n in columns for n in names):
                props[prop] = [columns.index(n) for n in names]
        box_l = [hi - lo for lo, hi in zip(info.box_lo, info.box_hi)]
        sample_rate = info.timestep1 - info.timestep0 if info.n_configurations > 1 else None
        self._layout = dict(n_particles=info.n_atoms, n_configs=info.n_configurations,
                            columns=columns, id_col=info.id_column, species=species, props=props,
                            box_l=box_l, sample_rate=sample_rate)
        return self._layout

    def _sort_by_id(self) -> bool:
        return not self.trajectory_is_sorted_by_ids
