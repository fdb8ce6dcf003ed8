"""
Project: a directory, its ``project.db`` and the experiment registry — the slice of the
reference's Project the RDF call needs (mdsuite/project/project.py:45-338).
"""
from __future__ import annotations

import logging
import os
from typing import Dict, Union

from mdsuite_b200.database.project_database import ProjectDatabase
from mdsuite_b200.experiment.experiment import Experiment
from mdsuite_b200.experiment.run import RunComputation
from mdsuite_b200.utils.units import Units

log = logging.getLogger(__name__)


class _Experiments(dict):
    """``project.experiments.NaCl`` and ``project.experiments["NaCl"]``."""

    def __getattr__(self, item):
        try:
            return self[item]
        except KeyError:
            raise AttributeError(item)


class Project:
    def __init__(self, name: str = None, storage_path: str = "./", description: str = None):
        self.name = "MDSuite_Project" if name is None else name
        self.storage_path = str(storage_path)
        self.project_path = os.path.join(self.storage_path, self.name)
        os.makedirs(self.project_path, exist_ok=True)
        self.database = ProjectDatabase(os.path.join(self.project_path, "project.db"), description)
        self._experiments: Dict[str, Experiment] = {}
        self._attach_file_logger()
        for exp_name in self.database.list_experiments():
            self._experiments[exp_name] = Experiment(self, exp_name)

    def _attach_file_logger(self):
        """project.py:132-145 — DEBUG log to <project>/mdsuite.log."""
        root = logging.getLogger("mdsuite_b200")
        path = os.path.join(self.project_path, "mdsuite.log")
        if not any(getattr(h, "baseFilename", None) == os.path.abspath(path) for h in root.handlers):
            handler = logging.FileHandler(path)
            handler.setLevel(logging.DEBUG)
            handler.setFormatter(logging.Formatter("%(asctime)s %(levelname)s (%(module)s): %(message)s"))
            root.addHandler(handler)
            root.setLevel(logging.DEBUG)

    @property
    def description(self):
        return self.database.description

    @description.setter
    def description(self, value: str):
        self.database.description = value

    @property
    def experiments(self) -> _Experiments:
        return _Experiments(self._experiments)

    @property
    def active_experiments(self) -> _Experiments:
        active = self.database.list_experiments()
        return _Experiments({k: v for k, v in self._experiments.items() if active.get(k)})

    def activate_experiments(self, names):
        for n in ([names] if isinstance(names, str) else names):
            self.database.set_active(n, True)

    def disable_experiments(self, names):
        for n in ([names] if isinstance(names, str) else names):
            self.database.set_active(n, False)

    def add_experiment(self, name: str = None, timestep: float = None, temperature: float = None,
                       units: Union[str, Units] = None, cluster_mode: bool = None,
                       active: bool = True, simulation_data=None,
                       update_with_pubchempy: bool = False) -> Experiment:
        """project.py:157-245."""
        if name is None:
            name = f"Experiment_{len(self._experiments):02d}"
        if name in self._experiments:
            log.info("This experiment already exists")
            exp = self._experiments[name]
        else:
            exp = Experiment(self, name, time_step=timestep, temperature=temperature, units=units,
                             cluster_mode=bool(cluster_mode))
            self.database.set_active(name, active)
            self._experiments[name] = exp
        if simulation_data is not None:
            exp.add_data(simulation_data)
        return exp

    @property
    def run(self) -> RunComputation:
        """project.py:308-316 — run over all active experiments; returns {name: Computation}."""
        return RunComputation(experiments=list(self.active_experiments.values()))
