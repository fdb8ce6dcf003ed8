"""
Extended XYZ reader — drop-in for the reference's ``EXTXYZFile``
(mdsuite/file_io/extxyz_files.py:55-296, tabular_text_files.py:122-220) on top of the native
reader (include/mds_io.h ``mds_xyz_open``, mdsuite_b200/_native): the file is mmap'ed, its
configurations are indexed and parsed by all host threads, values arrive as
float32(float64(token)) exactly as the reference's NumPy conversion + float32 store produce them.

Same conventions as the reference: two header lines per configuration (atom count, key=value
line); box lengths are entries 0, 4, 8 of ``Lattice="..."``; ``Properties=species:S:1:pos:R:3:...``
names the columns; species come from the ``species`` column of the FIRST configuration in order of
first appearance; atoms keep their file order; the sample rate is the rounded difference of the
``time=`` values of the first two configurations.

Where the reference's header handling is fragile this reader does the following:
  * a property the reference does not know in front of known ones (``Z:I:1``): the reference's
    column counter does not advance over it (:279-288), so the following properties are read from
    shifted columns.  Reproduced as is (pinned by tests/golden/extxyz_unknown_property.extxyz), with
    a warning; map the property through ``custom_data_map`` to get the true columns.
  * ``Lattice`` not being the first key of the line, or a file with a single configuration: the
    reference raises (IndexError :224, StopIteration :127); here both are read as intended.
"""
from __future__ import annotations

import logging
import pathlib
from typing import Dict, List, Optional, Tuple

import numpy as np

from mdsuite_b200.file_io.tabular_text_files import NativeTabularTextFile

log = logging.getLogger(__name__)

# extended-XYZ property names of the per-atom properties the reference recognises
# (mdsuite/file_io/extxyz_files.py:44-52; names from database/mdsuite_properties.py)
VAR_NAMES = {
    "Positions": "pos",
    "Velocities": "vel",
    "Forces": "force",
    "Stress": "stress",
    "Energy": "energies",
    "Time": "time",
    "Momenta": "momenta",
}


def get_box_l(header: str) -> List[float]:
    """Box lengths from ``Lattice="ax ay az bx by bz cx cy cz"`` (reference ``_get_box_l``
    :183-224): entries 0, 4 and 8 of the quoted list that starts at the token holding "Lattice"."""
    data = header.split()
    start = next((i for i, item in enumerate(data) if "Lattice" in item), None)
    if start is None:
        raise RuntimeError("Could not find lattice size in file header")
    stop = next((i for i in range(start, len(data)) if data[i][-1] == '"'), None)
    if stop is None:
        raise RuntimeError("Could not find the end of the lattice in file header")
    lattice = data[start:stop + 1]
    lattice[0] = lattice[0].split("=")[1].replace('"', "")
    lattice[-1] = lattice[-1].replace('"', "")
    values = np.array(lattice).astype(float)
    return [float(values[0]), float(values[4]), float(values[8])]


def get_time(header: str) -> Optional[float]:
    """Value of the last token holding "time" (reference ``_get_time`` :227-245); a trailing comma
    after the number is tolerated the same way."""
    time = None
    for item in header.split():
        if VAR_NAMES["Time"] in item:
            value = item.split("=")[-1]
            try:
                time = float(value)
            except ValueError:
                time = float(value.split(",")[0])
    return time


def get_property_to_column_idx_dict(header: str, var_names: Dict[str, str]
                                    ) -> Tuple[int, Dict[str, List[int]]]:
    """(species column, {property: columns}) from the ``Properties=`` token (reference
    ``_get_property_to_column_idx_dict`` :248-296).  The column counter advances only over
    ``species`` and the properties in ``var_names`` — the reference's rule, see the module doc."""
    properties_string = None
    for item in header.split():
        if "Properties" in item:
            properties_string = item
    if properties_string is None:
        raise RuntimeError("Could not find properties in header")
    properties_list = (properties_string.split("=")[1].replace(":S", "").replace(":R", "")
                       .split(":"))
    names = list(var_names.keys())
    file_names = list(var_names.values())
    summary: Dict[str, List[int]] = {}
    species_index = None
    index = 0
    for idx, item in enumerate(properties_list):
        if item == "species":
            species_index = index
            index += 1
        if item in file_names:
            length = int(properties_list[idx + 1])
            summary[names[file_names.index(item)]] = [index + i for i in range(length)]
            index += length
    if species_index is None:
        raise RuntimeError("could not find species in header")
    return species_index, summary


def _declared_columns(header: str) -> Optional[int]:
    """Columns per atom line according to ``Properties=name:T:n:...`` (None if malformed)."""
    for item in header.split():
        if item.startswith("Properties="):
            parts = item.split("=", 1)[1].split(":")
            if len(parts) % 3:
                return None
            try:
                return sum(int(n) for n in parts[2::3])
            except ValueError:
                return None
    return None


class EXTXYZFile(NativeTabularTextFile):
    def __init__(self, file_path, custom_data_map: dict = None, chunk_bytes: int = 256 << 20,
                 n_threads: int = 0):
        """``custom_data_map`` maps further property names to the names used in the file's
        ``Properties=`` list, e.g. {"Reduced_Momentum": "redmom"} for "redmom:R:3"
        (reference :59-73)."""
        self.file_path = str(pathlib.Path(file_path))
        self.var_names = dict(VAR_NAMES)
        self.var_names.update(custom_data_map or {})
        self.chunk_bytes = chunk_bytes
        self.n_threads = n_threads
        self.n_header_lines = 2
        self._mdata = None
        self._layout = None
        self._dump = None

    def _open(self):
        if self._dump is None:
            from mdsuite_b200 import _native
            self._dump = _native.XyzDump(self.file_path, n_threads=self.n_threads)
        return self._dump

    # ---- header analysis (reference extxyz_files.py:93-181) ----------------------------------
    def _analyse(self):
        if self._layout is not None:
            return self._layout
        dump = self._open()
        info = dump.info
        header = dump.header_line(0, 1)
        species_idx, props = get_property_to_column_idx_dict(header, self.var_names)
        for prop, cols in props.items():
            if cols and cols[-1] >= len(info.columns):
                raise ValueError(f"property {prop} names column {cols[-1]} but the atom lines have "
                                 f"{len(info.columns)} columns")
        used = 1 + sum(len(c) for c in props.values())
        declared = _declared_columns(header)
        if declared is not None and declared != used:
            log.warning("%s: the Properties list declares %d columns but only %d belong to known "
                        "properties; like the reference, columns are counted over known properties "
                        "only, so values may be read from shifted columns — name the other "
                        "properties in custom_data_map", self.file_path, declared, used)
        labels = dump.first_frame_labels(species_idx, sort_by_id=False)
        species: Dict[str, List[int]] = {}
        for i, name in enumerate(labels):
            species.setdefault(name, []).append(i)
        box_l = get_box_l(header)
        sample_rate = None
        if info.n_configurations > 1:
            time_0, time_1 = get_time(header), get_time(dump.header_line(1, 1))
            if time_0 is not None and time_1 is not None:
                sample_rate = int(round(time_1 - time_0))
        if sample_rate is None:
            log.warning("Could not read sample rate from file. Please adjust the sample rate "
                        "manually if required.")
        self._layout = dict(n_particles=info.n_atoms, n_configs=info.n_configurations,
                            species_idx=species_idx, species=species, props=props, box_l=box_l,
                            sample_rate=sample_rate)
        return self._layout
