"""Small helpers the RDF path touches (reference mdsuite/utils/meta_functions.py)."""
import numpy as np


def join_path(a, b) -> str:
    """HDF5-style dataset path '<species>/<property>' (meta_functions.py:73-93)."""
    return f"{a}/{b}"


def split_array(data: np.ndarray, condition: np.ndarray) -> list:
    """Split by a condition; a single-element list when nothing fails it (meta_functions.py:468-490)."""
    passed, failed = data[condition], data[~condition]
    if len(failed) == 0:
        return [passed]
    return [passed, failed]
