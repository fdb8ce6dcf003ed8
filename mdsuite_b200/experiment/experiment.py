"""
Experiment: one trajectory and its metadata — the slice of the reference's Experiment the RDF call
needs (mdsuite/experiment/experiment.py:89-222, :459-552, :599-639 and
mdsuite/database/experiment_database.py:46-77, :195-260).

Metadata live in ``project.db`` (experiment_attributes / experiment_species rows); positions live in
the frame-major trajectory store under ``<project>/<experiment>/database/``.
"""
from __future__ import annotations

import logging
import os
import pathlib
from types import SimpleNamespace
from typing import Dict, List, Union

import numpy as np

from mdsuite_b200.database.simulation_database import (Database, HDF5Database, PropertyInfo,
                                                       SpeciesInfo, TrajectoryChunkData,
                                                       TrajectoryMetadata)
from mdsuite_b200.experiment.run import RunComputation
from mdsuite_b200.file_io.file_read import FileProcessor
from mdsuite_b200.file_io.lammps_trajectory_files import LAMMPSTrajectoryFile
from mdsuite_b200.file_io.extxyz_files import EXTXYZFile
from mdsuite_b200.file_io.script_input import ScriptInput
from mdsuite_b200.utils.units import Units, resolve_units

log = logging.getLogger(__name__)


def _get_processor(simulation_data) -> FileProcessor:
    """experiment.py:62-86 — pick a reader from the suffix, or take a FileProcessor as is."""
    if isinstance(simulation_data, FileProcessor):
        return simulation_data
    if isinstance(simulation_data, (str, pathlib.Path)):
        suffix = pathlib.Path(simulation_data).suffix
        if suffix in (".lammpstraj", ".lammpstrj", ".dump"):
            return LAMMPSTrajectoryFile(simulation_data)
        if suffix == ".extxyz":
            return EXTXYZFile(simulation_data)
        raise ValueError(f"datafile ending '{suffix}' not recognized; pass a FileProcessor "
                         f"(LAMMPS dumps use .lammpstraj, extended XYZ files .extxyz)")
    raise ValueError(f"simulation_data of type {type(simulation_data)} not supported")


class Experiment:
    def __init__(self, project, name: str, time_step: float = None, temperature: float = None,
                 units: Union[str, Units] = None, cluster_mode: bool = False):
        self.project = project
        self.name = name
        self.cluster_mode = cluster_mode
        self.storage_path = os.path.abspath(project.storage_path)
        self.experiment_path = os.path.join(self.storage_path, project.name, name)
        self.database_path = os.path.join(self.experiment_path, "database")
        self.figures_path = os.path.join(self.experiment_path, "figures")
        os.makedirs(self.database_path, exist_ok=True)
        os.makedirs(self.figures_path, exist_ok=True)
        project.database.add_experiment(name)
        hdf5 = os.path.join(self.experiment_path, "database.hdf5")
        if os.path.exists(hdf5) and not os.path.exists(os.path.join(self.database_path, "index.json")):
            # an existing MDSuite project: read its atom-major gzip store in place (own HDF5
            # reader); datasets are over-allocated on resize, the SQL record has the true length
            self.database = HDF5Database(hdf5, n_configurations=self._get("number_of_configurations"))
        else:
            self.database = Database(self.database_path)
        if time_step is not None:
            self._set("time_step", time_step)
        if temperature is not None:
            self._set("temperature", temperature)
        if units is not None or self._get("units") is None:
            self.units = resolve_units(units)
        if self._get("version") is None:
            self._set("version", 0)

    # ---- SQL-backed attributes (experiment_database.py) -------------------------------------
    def _get(self, name, default=None):
        return self.project.database.get_attribute(self.name, name, default)

    def _set(self, name, value):
        self.project.database.set_attribute(self.name, name, value)

    @property
    def units(self) -> Units:
        return Units(**self._get("units"))

    @units.setter
    def units(self, value: Units):
        u = resolve_units(value)
        self._set("units", {k: getattr(u, k) for k in Units.__dataclass_fields__})

    @property
    def version(self) -> int:
        return int(self._get("version", 0))

    @property
    def box_array(self) -> list:
        return self._get("box_array")

    @property
    def volume(self) -> float:
        """experiment_database.py:318-331 — product of the box lengths."""
        return float(np.prod(self.box_array))

    @property
    def number_of_configurations(self) -> int:
        return int(self._get("number_of_configurations", 0))

    @property
    def number_of_atoms(self) -> int:
        return int(self._get("number_of_atoms", 0))

    @property
    def sample_rate(self):
        return self._get("sample_rate")

    @property
    def temperature(self):
        return self._get("temperature")

    @property
    def time_step(self):
        return self._get("time_step")

    @property
    def read_files(self) -> list:
        return list(self._get("read_files", []))

    @property
    def species(self) -> Dict[str, SimpleNamespace]:
        """{name: obj(n_particles, mass, charge, properties)} in storage order."""
        raw = self.project.database.get_species(self.name, molecule=False)
        return {k: SimpleNamespace(name=k, **v) for k, v in raw.items()}

    @property
    def molecules(self) -> Dict[str, SimpleNamespace]:
        raw = self.project.database.get_species(self.name, molecule=True)
        return {k: SimpleNamespace(name=k, **v) for k, v in raw.items()}

    @property
    def run(self) -> RunComputation:
        """experiment.py:213-222."""
        return RunComputation(experiment=self)

    # ---- ingest (experiment.py:459-552, :599-639) -------------------------------------------
    def add_data(self, simulation_data, force: bool = False, update_with_pubchempy: bool = False):
        if isinstance(simulation_data, list):
            for item in simulation_data:
                self.add_data(item, force=force)
            return
        proc = _get_processor(simulation_data)
        already_read = str(proc) in self.read_files
        if already_read and not force:
            log.info("This file has already been read, skipping this now.")
            return
        if not isinstance(self.database, Database):
            raise RuntimeError("cannot append to a reference HDF5 database; create a new experiment")
        metadata = proc.metadata
        first_ingest = self.number_of_configurations == 0
        self._store_metadata(metadata, first_ingest)
        if first_ingest:
            self.database.initialize_database(metadata.species_list)
        for chunk in proc.get_configurations_generator():
            self.database.add_data(chunk)
        self._set("number_of_configurations",
                  self.number_of_configurations + metadata.n_configurations)
        self._set("version", self.version + 1)  # invalidates cached computations (:547)
        self._set("read_files", self.read_files + [str(proc)])

    def _store_metadata(self, metadata: TrajectoryMetadata, first_ingest: bool):
        if first_ingest:
            self._set("box_array", [float(v) for v in metadata.box_l])
            self._set("sample_rate", metadata.sample_rate)
            species = {sp.name: {"n_particles": int(sp.n_particles), "mass": sp.mass,
                                 "charge": sp.charge,
                                 "properties": [p.name for p in sp.properties]}
                       for sp in metadata.species_list}
            self.project.database.set_species(self.name, species)
            self._set("number_of_atoms", int(sum(sp.n_particles for sp in metadata.species_list)))
            if metadata.temperature is not None and self.temperature is None:
                self._set("temperature", metadata.temperature)
        else:
            have = {k: v.n_particles for k, v in self.species.items()}
            new = {sp.name: sp.n_particles for sp in metadata.species_list}
            if have != new:
                raise ValueError("appended data has a different species layout")

    def add_positions(self, positions: np.ndarray, species: Dict[str, int], box, name: str = None,
                      sample_rate: int = 1):
        """Convenience ingest of an in-memory (F, N, 3) array in species-contiguous order — a
        ``ScriptInput`` built for you (file_io/script_input.py)."""
        positions = np.asarray(positions, dtype=np.float32)
        species_list = [SpeciesInfo(n, c, [PropertyInfo("Positions", 3)]) for n, c in species.items()]
        chunk = TrajectoryChunkData(species_list, positions.shape[0])
        off = 0
        for n, c in species.items():
            chunk.add_data(positions[:, off:off + c], 0, n, "Positions")
            off += c
        md = TrajectoryMetadata(n_configurations=positions.shape[0], species_list=species_list,
                                box_l=list(np.asarray(box, dtype=float)), sample_rate=sample_rate)
        self.add_data(ScriptInput(chunk, md, name or f"{self.name}_in_memory_{self.version}"))

    def load_matrix(self, species: List[str] = None, property_name: str = "Positions",
                    select_slice=np.s_[:]):
        """Reference-shaped read: {path: (n_atoms, n_configurations, 3)} (experiment.py:554-597)."""
        species = species or list(self.species)
        paths = [f"{s}/{property_name}" for s in species]
        return self.database.load_data(paths, select_slice)
