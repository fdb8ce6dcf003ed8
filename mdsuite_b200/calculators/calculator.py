"""
Calculator runtime: the ``@call`` orchestration and the results-cache plumbing
(reference mdsuite/calculators/calculator.py:52-148, :151-317 and
mdsuite/database/calculator_database.py:91-248).
"""
from __future__ import annotations

import functools
import logging
from dataclasses import fields
from typing import Dict, List, Union

from mdsuite_b200.database.project_database import Computation
from mdsuite_b200.visualizer.d2_data_visualization import DataVisualizer2D

log = logging.getLogger(__name__)


def call(func):
    """Decorator of ``Calculator.__call__`` (calculator.py:52-148): for every experiment make a
    fresh calculator, store the user args, return the cached computation if one matches them and the
    experiment version; otherwise run, save, re-set the user args and reload; then plot."""

    @functools.wraps(func)
    def inner(self, *args, **kwargs) -> Union[Computation, Dict[str, Computation]]:
        from mdsuite_b200.distributed import distributed_context
        return_dict = self.experiment is None
        out = {}
        # Under torch.distributed every rank runs the analysis (its share of the configurations)
        # but ONE rank owns the project database: rank 0 looks the computation up and tells the
        # others whether it is a hit, rank 0 alone writes the result, and the others work with the
        # in-memory copy — no double rows, no sqlite lock contention, and no rank that skips
        # run_analysis (a late cache hit) while the others wait in the allreduce.
        rank, world, dist = distributed_context()
        for experiment in self.experiments:
            cls = self.__class__(experiment=experiment)
            func(cls, *args, **kwargs)
            data = cls.get_computation_data() if rank == 0 else None
            hit = data is not None
            if dist is not None:
                box = [hit]
                dist.broadcast_object_list(box, src=0)
                hit = bool(box[0])
            if not hit:
                cls.prepare_db_entry()
                cls.save_computation_args()
                cls.run_analysis()
                if rank == 0:
                    cls.save_db_data()
                else:
                    cls.keep_computation_in_memory()
                if dist is not None:
                    dist.barrier()  # the row is written before any rank goes on
                func(cls, *args, **kwargs)  # defaults filled in by check_input must not leak
                data = (cls.get_computation_data() if rank == 0 else None) or cls.last_computation
            elif dist is not None:
                box = [data]
                dist.broadcast_object_list(box, src=0)  # the cached result, for every rank
                data = box[0]
            if cls.plot and data is not None and rank == 0:
                cls.plotter = DataVisualizer2D(title=cls.analysis_name, path=experiment.figures_path)
                cls.plot_data(data.data_dict)
                cls.plotter.grid_show(cls.plot_array)
            out[cls.experiment.name] = data
        if return_dict:
            return out
        return out[self.experiment.name]

    return inner


class Calculator:
    """Base class: experiment binding, args dataclass, cache lookup, queue/save of results."""

    def __init__(self, experiment=None, experiments=None, **kwargs):
        self.experiment = experiment
        self.experiments = experiments if experiments is not None else (
            [experiment] if experiment is not None else [])
        self.args = None
        self.plot = False
        self.plot_array: List = []
        self.plotter = None
        self.analysis_name = None
        self.x_label = None
        self.y_label = None
        self.result_series_keys = None
        self._queued_data = []
        self._db_attributes = None
        self.last_computation = None
        self.save_to_db = True

    # ---- cache / persistence (calculator_database.py) ---------------------------------------
    def _args_dict(self) -> dict:
        out = {f.name: getattr(self.args, f.name) for f in fields(self.args)}
        out["version"] = self.experiment.version
        return out

    def get_computation_data(self):
        """calculator_database.py:103-172."""
        return self.experiment.project.database.find_computation(
            self.experiment.name, self.analysis_name, self._args_dict())

    def prepare_db_entry(self):
        """calculator_database.py:91-101."""
        self._queued_data = []

    def save_computation_args(self):
        """Snapshot the USER args (cutoff=None stays None), calculator_database.py:174-194."""
        self._db_attributes = self._args_dict()

    def queue_data(self, data: dict, subjects: list):
        """calculator_database.py:236-248."""
        self._queued_data.append((data, list(subjects)))

    def keep_computation_in_memory(self):
        """The finished run as a Computation that is not written to the project database
        (``save=False``, and every rank but 0 of a distributed run)."""
        attrs = [(k, {"serialized_value": v}) for k, v in self._db_attributes.items()]
        results = [(d, [(s, c) for s, c in _count_subjects(subj)]) for d, subj in self._queued_data]
        self.last_computation = Computation(-1, self.analysis_name, -1, attrs, results)

    def save_db_data(self):
        """calculator_database.py:196-234 — only reached when the run completed."""
        if not self.save_to_db:
            self.keep_computation_in_memory()
            return
        self.experiment.project.database.save_computation(
            self.experiment.name, self.analysis_name, self._db_attributes, self._queued_data)

    # ---- run --------------------------------------------------------------------------------
    def run_analysis(self):
        """calculator.py:310-317."""
        self.run_calculator()

    def run_calculator(self):
        raise NotImplementedError

    def run_visualization(self, x_data, y_data, title: str, layouts=None):
        """calculator.py:267-308."""
        self.plot_array.append(self.plotter.construct_plot(
            x_data=x_data, y_data=y_data, title=title, x_label=self.x_label, y_label=self.y_label))

    def plot_data(self, data: dict):
        for selected_species, val in data.items():
            self.run_visualization(x_data=val[self.result_series_keys[0]],
                                   y_data=val[self.result_series_keys[1]], title=selected_species)


def resolve_species_data(experiment, args, loaded_property: str):
    """Store paths and per-species atom selections of a trajectory calculator (reference
    calculators/radial_distribution_function.py:565-599, angular_distribution_function.py:565-575):
    ``atom_selection`` is ``np.s_[:]`` or a dict {species: [indices]}."""
    from mdsuite_b200.utils.meta_functions import join_path
    if args.molecules and not getattr(experiment, "molecules", None):
        raise ValueError("molecules=True needs molecule groups in the experiment "
                         "(the MolecularMap transformation is outside this package's scope)")
    paths = [join_path(item, loaded_property) for item in args.species]
    for p in paths:
        if not experiment.database.check_existence(p):
            raise KeyError(f"dataset {p} not found in the trajectory store")
    if isinstance(args.atom_selection, dict):
        selections = [list(args.atom_selection[item]) for item in args.species]
    elif isinstance(args.atom_selection, slice) and args.atom_selection == slice(None):
        selections = [None] * len(paths)
    else:
        raise ValueError("atom_selection must be np.s_[:] or a dict {species: [indices]}")
    return paths, selections


def _count_subjects(subjects):
    seen = {}
    for s in subjects:
        seen[s] = seen.get(s, 0) + 1
    return list(seen.items())
