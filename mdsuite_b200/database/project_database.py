"""
Results store: the reference's SQL layout in ``<storage>/<project>/project.db`` through stdlib
sqlite3 (SQLAlchemy is not available in this image).

Tables, columns and JSON conventions follow the reference's ORM so an MDSuite install can open the
file (mdsuite/database/scheme.py):
  projects(id, description)                                     :65-72
  experiments(id, name, active, project_id)                     :75-95
  experiment_attributes(id, name, data, experiment_id)          :131-156
  experiment_species(id, name, data, molecule, experiment_id)   :167-187
  computations(id, name, experiment_id)                         :193-212
  computation_attributes(id, name, data, computation_id)        :345-357
  computation_results(id, data, computation_id)                 :363-377
  species_association(computation_results_id, experiment_species_id, count)  :39-57
Every ``data`` column is a JSON string (database/types.py:33-55); non-dict values are wrapped as
{"serialized_value": v} (database/calculator_database.py:60-67).
"""
from __future__ import annotations

import json
import os
import sqlite3
from collections import Counter
from contextlib import contextmanager
from typing import Any, Dict, List, Optional, Sequence

SCHEMA = """
CREATE TABLE IF NOT EXISTS projects (id INTEGER PRIMARY KEY, description VARCHAR);
CREATE TABLE IF NOT EXISTS experiments (
    id INTEGER PRIMARY KEY, name VARCHAR, active BOOLEAN DEFAULT 0,
    project_id INTEGER REFERENCES projects(id) ON DELETE CASCADE);
CREATE TABLE IF NOT EXISTS experiment_attributes (
    id INTEGER PRIMARY KEY, name VARCHAR, data VARCHAR,
    experiment_id INTEGER REFERENCES experiments(id) ON DELETE CASCADE);
CREATE TABLE IF NOT EXISTS experiment_species (
    id INTEGER PRIMARY KEY, name VARCHAR, data VARCHAR, molecule BOOLEAN DEFAULT 0,
    experiment_id INTEGER REFERENCES experiments(id) ON DELETE CASCADE);
CREATE TABLE IF NOT EXISTS computations (
    id INTEGER PRIMARY KEY, name VARCHAR DEFAULT 'Computation',
    experiment_id INTEGER REFERENCES experiments(id) ON DELETE CASCADE);
CREATE TABLE IF NOT EXISTS computation_attributes (
    id INTEGER PRIMARY KEY, name VARCHAR, data VARCHAR,
    computation_id INTEGER REFERENCES computations(id) ON DELETE CASCADE);
CREATE TABLE IF NOT EXISTS computation_results (
    id INTEGER PRIMARY KEY, data VARCHAR,
    computation_id INTEGER REFERENCES computations(id) ON DELETE CASCADE);
CREATE TABLE IF NOT EXISTS species_association (
    computation_results_id INTEGER REFERENCES computation_results(id) ON DELETE CASCADE,
    experiment_species_id INTEGER REFERENCES experiment_species(id) ON DELETE CASCADE,
    count INTEGER DEFAULT 1,
    PRIMARY KEY (computation_results_id, experiment_species_id));
"""


def is_jsonable(x) -> bool:
    try:
        json.dumps(x)
        return True
    except (TypeError, OverflowError, ValueError):
        return False


def conv_to_db(val) -> dict:
    """database/calculator_database.py:60-67."""
    if is_jsonable(val):
        if not isinstance(val, dict):
            val = {"serialized_value": val}
    else:
        val = {"serialized_value": str(val)}
    return val


def _dumps(val) -> str:
    # the reference's JSONEncodedDict stores json.dumps(value); NaN / Infinity are written as the
    # bare tokens Python's json emits and reads back (the RDF's y[0] is nan or inf, quirk Q2)
    return json.dumps(val)


class Computation:
    """Plain-Python twin of ``db.Computation`` (database/scheme.py:193-333)."""

    def __init__(self, id: int, name: str, experiment_id: int,
                 attributes: Sequence[tuple], results: Sequence[tuple]):
        self.id = id
        self.name = name
        self.experiment_id = experiment_id
        self.computation_attributes = [{"name": n, "data": d} for n, d in attributes]
        # results: (data, [(species_name, count), ...])
        self.computation_results = [{"data": d, "species": s} for d, s in results]

    def __repr__(self):
        return f"Exp{self.experiment_id}_{self.name}_{self.id}"

    @property
    def data_dict(self) -> dict:
        """{"Na_Cl": {"x": [...], "y": [...]}, ...} — keys joined from the species associations
        (a same-species pair is one association with count 2), scheme.py:225-268."""
        out = {}
        for result in self.computation_results:
            keys: List[str] = []
            for species_name, count in result["species"]:
                keys += count * [species_name]
            key = "_".join(keys) or "System"
            out[key] = result["data"]
        return out

    def __getitem__(self, item):
        try:
            return self.data_dict[item]
        except KeyError:
            raise KeyError(f"Could not find {item} - available keys are {self.data_dict.keys()}")

    def keys(self) -> list:
        return list(self.data_dict.keys())

    @property
    def computation_parameter(self) -> dict:
        return {a["name"]: a["data"]["serialized_value"] for a in self.computation_attributes}

    @property
    def data_range(self) -> Optional[int]:
        for a in self.computation_attributes:
            if a["name"] == "data_range":
                return int(a["data"]["serialized_value"])
        return None


class ProjectDatabase:
    """One sqlite file per project (mdsuite/database/project_database.py:41-43,
    database_base.py:59-73)."""

    def __init__(self, path: str, description: Optional[str] = None):
        self.path = path
        new = not os.path.exists(path)
        with self.session() as con:
            con.executescript(SCHEMA)
            if new or con.execute("SELECT COUNT(*) FROM projects").fetchone()[0] == 0:
                con.execute("INSERT INTO projects (description) VALUES (?)", (description,))

    @contextmanager
    def session(self):
        con = sqlite3.connect(self.path)
        con.execute("PRAGMA foreign_keys = ON")
        try:
            yield con
            con.commit()
        finally:
            con.close()

    # ---- project ------------------------------------------------------------------------
    @property
    def description(self) -> Optional[str]:
        with self.session() as con:
            return con.execute("SELECT description FROM projects ORDER BY id LIMIT 1").fetchone()[0]

    @description.setter
    def description(self, value: str):
        with self.session() as con:
            con.execute("UPDATE projects SET description = ? WHERE id = (SELECT MIN(id) FROM projects)",
                        (value,))

    # ---- experiments --------------------------------------------------------------------
    def experiment_id(self, name: str) -> Optional[int]:
        with self.session() as con:
            row = con.execute("SELECT id FROM experiments WHERE name = ?", (name,)).fetchone()
        return row[0] if row else None

    def add_experiment(self, name: str, active: bool = True) -> int:
        existing = self.experiment_id(name)
        if existing is not None:
            return existing
        with self.session() as con:
            pid = con.execute("SELECT MIN(id) FROM projects").fetchone()[0]
            cur = con.execute("INSERT INTO experiments (name, active, project_id) VALUES (?, ?, ?)",
                              (name, int(active), pid))
            return cur.lastrowid

    def list_experiments(self) -> Dict[str, bool]:
        with self.session() as con:
            rows = con.execute("SELECT name, active FROM experiments ORDER BY id").fetchall()
        return {n: bool(a) for n, a in rows}

    def set_active(self, name: str, active: bool):
        with self.session() as con:
            con.execute("UPDATE experiments SET active = ? WHERE name = ?", (int(active), name))

    def set_attribute(self, experiment: str, name: str, value: Any):
        eid = self.experiment_id(experiment)
        with self.session() as con:
            con.execute("DELETE FROM experiment_attributes WHERE experiment_id = ? AND name = ?",
                        (eid, name))
            con.execute("INSERT INTO experiment_attributes (name, data, experiment_id) VALUES (?, ?, ?)",
                        (name, _dumps(conv_to_db(value)), eid))

    def get_attribute(self, experiment: str, name: str, default=None):
        eid = self.experiment_id(experiment)
        with self.session() as con:
            row = con.execute("SELECT data FROM experiment_attributes WHERE experiment_id = ? AND "
                              "name = ?", (eid, name)).fetchone()
        if row is None:
            return default
        data = json.loads(row[0])
        return data["serialized_value"] if set(data) == {"serialized_value"} else data

    def set_species(self, experiment: str, species: Dict[str, dict], molecule: bool = False):
        eid = self.experiment_id(experiment)
        with self.session() as con:
            for name, data in species.items():
                row = con.execute("SELECT id FROM experiment_species WHERE experiment_id = ? AND "
                                  "name = ? AND molecule = ?", (eid, name, int(molecule))).fetchone()
                if row:
                    con.execute("UPDATE experiment_species SET data = ? WHERE id = ?",
                                (_dumps(data), row[0]))
                else:
                    con.execute("INSERT INTO experiment_species (name, data, molecule, experiment_id) "
                                "VALUES (?, ?, ?, ?)", (name, _dumps(data), int(molecule), eid))

    def get_species(self, experiment: str, molecule: bool = False) -> Dict[str, dict]:
        eid = self.experiment_id(experiment)
        with self.session() as con:
            rows = con.execute("SELECT name, data FROM experiment_species WHERE experiment_id = ? AND "
                               "molecule = ? ORDER BY id", (eid, int(molecule))).fetchall()
        return {n: json.loads(d) for n, d in rows}

    # ---- computations -------------------------------------------------------------------
    def _load_computation(self, con, cid: int) -> Computation:
        name, eid = con.execute("SELECT name, experiment_id FROM computations WHERE id = ?",
                                (cid,)).fetchone()
        attrs = [(n, json.loads(d)) for n, d in con.execute(
            "SELECT name, data FROM computation_attributes WHERE computation_id = ? ORDER BY id",
            (cid,))]
        results = []
        for rid, data in con.execute("SELECT id, data FROM computation_results WHERE computation_id = ? "
                                     "ORDER BY id", (cid,)).fetchall():
            species = con.execute(
                "SELECT s.name, a.count FROM species_association a JOIN experiment_species s "
                "ON s.id = a.experiment_species_id WHERE a.computation_results_id = ? "
                "ORDER BY a.rowid", (rid,)).fetchall()
            results.append((json.loads(data), [(n, c) for n, c in species]))
        return Computation(cid, name, eid, attrs, results)

    def find_computation(self, experiment: str, analysis_name: str,
                         attributes: Dict[str, Any]) -> Optional[Computation]:
        """A computation whose stored attributes equal ``attributes`` (user args + experiment
        version) — the cache lookup of calculator_database.py:103-172."""
        eid = self.experiment_id(experiment)
        want = {k: _dumps(conv_to_db(v)) for k, v in attributes.items()}
        with self.session() as con:
            cids = [r[0] for r in con.execute(
                "SELECT id FROM computations WHERE experiment_id = ? AND name = ? ORDER BY id",
                (eid, analysis_name))]
            for cid in cids:
                have = dict(con.execute("SELECT name, data FROM computation_attributes WHERE "
                                        "computation_id = ?", (cid,)).fetchall())
                if all(have.get(k) == v for k, v in want.items()):
                    return self._load_computation(con, cid)
        return None

    def count_computations(self, experiment: str, analysis_name: str) -> int:
        """Stored computations of one calculator for an experiment (rows of ``computations``)."""
        eid = self.experiment_id(experiment)
        with self.session() as con:
            return int(con.execute("SELECT COUNT(*) FROM computations WHERE experiment_id = ? AND "
                                   "name = ?", (eid, analysis_name)).fetchone()[0])

    def save_computation(self, experiment: str, analysis_name: str, attributes: Dict[str, Any],
                         results: Sequence[tuple]) -> int:
        """results: [(data_dict, [subject, ...]), ...] — one computation_results row per entry plus
        its species associations (calculator_database.py:196-234).  Written in one transaction, so
        nothing is stored unless the run completed (calculators/calculator.py:123-126)."""
        eid = self.experiment_id(experiment)
        with self.session() as con:
            cid = con.execute("INSERT INTO computations (name, experiment_id) VALUES (?, ?)",
                              (analysis_name, eid)).lastrowid
            for k, v in attributes.items():
                con.execute("INSERT INTO computation_attributes (name, data, computation_id) "
                            "VALUES (?, ?, ?)", (k, _dumps(conv_to_db(v)), cid))
            for data, subjects in results:
                rid = con.execute("INSERT INTO computation_results (data, computation_id) VALUES (?, ?)",
                                  (_dumps(data), cid)).lastrowid
                ids = []
                for s in subjects:
                    row = con.execute("SELECT id FROM experiment_species WHERE name = ? AND "
                                      "experiment_id = ? ORDER BY id LIMIT 1", (s, eid)).fetchone()
                    if row is not None:
                        ids.append(row[0])
                for sid, count in Counter(ids).items():
                    con.execute("INSERT INTO species_association (computation_results_id, "
                                "experiment_species_id, count) VALUES (?, ?, ?)", (rid, sid, count))
        return cid
