"""Generate tests/golden/adf_golden.json by EXECUTING THE REFERENCE'S OWN angular-distribution code

    mdsuite/calculators/angular_distribution_function.py   (_prepare_data_structure,
        _build_histograms, _compute_rijk_matrices, _compute_angles, _compute_adfs, _format_data)
    mdsuite/utils/neighbour_list.py                        (get_neighbour_list, get_triplets,
                                                            get_triu_indicies)
    mdsuite/utils/linalg.py                                (get_angles, angle_between, unit_vector)

loaded unmodified from /root/reference over a NumPy stand-in for the TensorFlow leaf ops they call
(TensorFlow is not installed here).  What this pins: which (configuration, i, j, k) triples exist
(float16 cutoff, roll-based enumeration), which species combination takes which triple, the
histogram / density / float32 accumulation over batches, the x axis and max_peak.  What it cannot
pin: the float32 results of TensorFlow's own norm / divide / acos / pow kernels — the stand-ins
implement the contracts written in oracle/adf_oracle.py.

Runs only where /root/reference exists.  Re-run:  python tests/golden/make_adf_golden.py
"""
import importlib.util
import json
import os
import sys
import types
from types import SimpleNamespace
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/mdsuite"
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from make_reference_golden import _band_part, _norm  # noqa: E402  (the same leaf stand-ins)


class _TensorArray:
    def __init__(self, dtype, size):
        self.dtype, self.items = dtype, [None] * int(size)

    def write(self, idx, value):
        self.items[int(idx)] = np.asarray(value, dtype=self.dtype)
        return self

    def stack(self):
        return np.stack(self.items)


def _scatter_nd(indices, updates, shape):
    out = np.zeros([int(s) for s in shape], dtype=np.asarray(updates).dtype)
    idx = np.asarray(indices)
    out[tuple(idx[:, d] for d in range(idx.shape[1]))] = updates
    return out


def _gather_nd(params, indices):
    idx = np.asarray(indices)
    return np.asarray(params)[tuple(idx[:, d] for d in range(idx.shape[1]))]


def _where(cond, x=None, y=None):
    return np.argwhere(cond) if x is None else np.where(cond, x, y)


def _concat(values, axis=0):
    if not isinstance(values, (list, tuple)):   # tf.concat wraps a lone tensor: identity
        return np.asarray(values)
    return np.concatenate([np.asarray(v) for v in values], axis=axis)


def _einsum(spec, a, b):
    spec = spec.replace(" ", "")
    assert spec == "ij,ij->i" and a.dtype == np.float32 and b.dtype == np.float32
    p = a * b
    return (p[:, 0] + p[:, 1]) + p[:, 2]


def _acos(x):
    assert x.dtype == np.float32
    return np.arccos(x.astype(np.float64)).astype(np.float32)


def make_tf_shim():
    tf = types.ModuleType("tensorflow")
    tf.float16, tf.float32, tf.float64 = np.float16, np.float32, np.float64
    tf.int32, tf.int64, tf.bool = np.int32, np.int64, bool
    tf.Tensor = np.ndarray
    tf.dtypes = SimpleNamespace(int32=np.int32, float32=np.float32)  # default arguments elsewhere
    tf.function = lambda *a, **k: (a[0] if a and callable(a[0]) else (lambda f: f))
    tf.cast = lambda x, dtype: np.asarray(x).astype(dtype)
    tf.shape = lambda x: np.asarray(x).shape
    tf.ones = lambda shape, dtype=np.float32: np.ones([int(s) for s in shape], dtype=dtype)
    tf.ones_like = np.ones_like
    tf.stack = lambda vals, axis=0: np.stack([np.asarray(v) for v in vals], axis=axis)
    tf.unstack = lambda x, axis=0: [np.take(x, k, axis=axis) for k in range(x.shape[axis])]
    tf.concat = _concat
    tf.transpose = lambda x, perm=None: np.transpose(x, perm)
    tf.where = _where
    tf.gather = lambda params, indices, axis=0: np.take(params, indices, axis=axis)
    tf.gather_nd = _gather_nd
    tf.scatter_nd = _scatter_nd
    tf.split = lambda value, n: np.split(value, n)
    tf.roll = lambda x, shift, axis: np.roll(x, shift, axis=axis)
    tf.expand_dims = np.expand_dims
    tf.norm = _norm
    tf.einsum = _einsum
    tf.clip_by_value = lambda x, lo, hi: np.clip(x, np.float32(lo), np.float32(hi))
    tf.logical_and = np.logical_and
    tf.TensorArray = _TensorArray
    tf.math = SimpleNamespace(rint=np.rint, acos=_acos, argmax=lambda x: int(np.argmax(x)),
                              logical_and=lambda x, y: np.logical_and(x, y))
    tf.linalg = SimpleNamespace(norm=_norm, band_part=_band_part)
    return tf


class _F32(np.ndarray):
    """float32 array with the two tensor methods the reference calls on histograms."""

    def numpy(self):
        return np.asarray(self)


def load_reference_adf():
    tf = make_tf_shim()
    real_cast = tf.cast
    # `tf.cast(histogram, dtype=tf.float32)` results get `.numpy()` called on them later
    tf.cast = lambda x, dtype: (real_cast(x, dtype).view(_F32) if dtype is np.float32
                                else real_cast(x, dtype))
    stubs = {"tensorflow": tf}
    for name in ("GPUtil", "psutil", "bokeh", "bokeh.models", "bokeh.plotting"):
        stubs[name] = mock.MagicMock()
    for pkg in ("mdsuite", "mdsuite.calculators", "mdsuite.utils", "mdsuite.database"):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REF, *pkg.split(".")[1:])]
        stubs[pkg] = m
    calc = types.ModuleType("mdsuite.calculators.calculator")
    calc.call = lambda f: f
    stubs["mdsuite.calculators.calculator"] = calc
    traj = types.ModuleType("mdsuite.calculators.trajectory_calculator")
    traj.TrajectoryCalculator = type("TrajectoryCalculator", (), {"__init__": lambda self, **kw: None})
    stubs["mdsuite.calculators.trajectory_calculator"] = traj
    props = types.ModuleType("mdsuite.database.mdsuite_properties")
    props.mdsuite_properties = SimpleNamespace(positions=SimpleNamespace(name="Positions"))
    stubs["mdsuite.database.mdsuite_properties"] = props
    exc = types.ModuleType("mdsuite.utils.exceptions")
    exc.NoGPUInSystem = type("NoGPUInSystem", (Exception,), {})
    stubs["mdsuite.utils.exceptions"] = exc
    units = types.ModuleType("mdsuite.utils.units")
    units.golden_ratio = (1 + 5 ** 0.5) / 2
    stubs["mdsuite.utils.units"] = units
    sys.modules.update(stubs)

    def load(modname, relpath):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, relpath))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod

    load("mdsuite.utils.meta_functions", "utils/meta_functions.py")
    load("mdsuite.utils.linalg", "utils/linalg.py")
    load("mdsuite.utils.neighbour_list", "utils/neighbour_list.py")
    return load("mdsuite.calculators.angular_distribution_function",
                "calculators/angular_distribution_function.py")


def run_reference(adf_mod, pos_fm, species, box, number_of_bins, cutoff, number_of_configurations,
                  norm_power, n_batches, start=1, stop=None):
    """The reference's run_calculator loop over in-memory positions; ``n_batches`` is what its
    memory manager would have decided (the `batches` keyword overrides it the same way)."""
    cls = adf_mod.AngularDistributionFunction
    calc = object.__new__(cls)
    offs = np.concatenate([[0], np.cumsum(list(species.values()))])
    calc.experiment = SimpleNamespace(
        box_array=[float(b) for b in box], number_of_configurations=pos_fm.shape[0],
        species={n: SimpleNamespace(n_particles=c) for n, c in species.items()}, molecules={})
    calc.args = adf_mod.Args(number_of_bins=number_of_bins, cutoff=cutoff, start=start, stop=stop,
                             atom_selection=np.s_[:], data_range=1, correlation_time=1,
                             molecules=False, species=None,
                             number_of_configurations=number_of_configurations,
                             norm_power=norm_power)
    calc.use_tf_function = False
    calc.cutoff = cutoff
    calc.adf_minibatch = -1
    calc.bin_range = [0.0, 3.15]
    calc.norm_power = norm_power
    calc.override_n_batches = n_batches
    calc.dtype = np.float32
    calc.loaded_property = SimpleNamespace(name="Positions")
    calc.result_keys = ["max_peak"]
    calc.result_series_keys = ["angle", "adf"]
    calc.minibatch = False
    calc._run_dependency_check = lambda: None
    queued = []
    calc.queue_data = lambda data, subjects: queued.append((list(subjects), data))

    def _prepare_managers(path_list):
        calc.batch_size = max(1, number_of_configurations // n_batches)
        calc.memory_manager = SimpleNamespace()
    calc._prepare_managers = _prepare_managers
    seen_batches = []

    def get_batch_dataset(subject_list, loop_array, correct):
        for frames in loop_array:
            seen_batches.append([int(f) for f in frames])
            yield {str.encode(f"{name}/Positions"):
                   pos_fm[frames][:, offs[k]:offs[k + 1]].transpose(1, 0, 2).astype(np.float64)
                   for k, name in enumerate(subject_list)}
    calc.get_batch_dataset = get_batch_dataset
    with np.errstate(all="ignore"):
        calc.run_calculator()
    return calc, queued, seen_batches


def positions_for(case):
    rng = np.random.Generator(np.random.Philox(case["seed"]))
    n = sum(case["counts"])
    box = np.asarray(case["box"], dtype=np.float64)
    pos = rng.random((case["frames"], n, 3)) * box
    if case.get("unwrapped"):
        pos += rng.integers(-2, 3, size=pos.shape) * box
    return pos.astype(np.float32)


CASES = [
    dict(name="two_species_defaults", seed=401, counts=[14, 10], names=["Na", "Cl"], frames=7,
         box=[9.0, 9.0, 9.0], cutoff=3.5, nbins=50, norm_power=4, number_of_configurations=5,
         n_batches=5),
    dict(name="three_species_counts_only", seed=402, counts=[6, 9, 5], names=["O", "H", "C"],
         frames=4, box=[7.0, 8.0, 7.5], cutoff=3.0, nbins=36, norm_power=0,
         number_of_configurations=3, n_batches=1, start=0),
    dict(name="cutoff_not_a_float16_and_unwrapped", seed=403, counts=[20], names=["Ar"], frames=6,
         box=[8.0, 8.0, 8.0], cutoff=3.2, nbins=100, norm_power=2, number_of_configurations=4,
         n_batches=2, unwrapped=True),
]


def main():
    adf_mod = load_reference_adf()
    out = []
    for case in CASES:
        pos = positions_for(case)
        species = dict(zip(case["names"], case["counts"]))
        calc, queued, batches = run_reference(
            adf_mod, pos, species, case["box"], case["nbins"], case["cutoff"],
            case["number_of_configurations"], case["norm_power"], case["n_batches"],
            start=case.get("start", 1))
        rec = dict(case)
        rec["sample_configurations"] = [int(v) for v in calc.sample_configurations]
        rec["batches"] = batches
        rec["results"] = [{"subjects": s, "max_peak": float(d["max_peak"]),
                           "angle": [float(v).hex() for v in d["angle"]],
                           "adf": [float(np.float32(v)).hex() for v in d["adf"]]}
                          for s, d in queued]
        out.append(rec)
        print(case["name"], batches, [("-".join(s), round(d["max_peak"], 2),
                                      float(np.sum(d["adf"]))) for s, d in queued])
    path = os.path.join(HERE, "adf_golden.json")
    with open(path, "w") as fh:
        json.dump({"generator": "tests/golden/make_adf_golden.py",
                   "source": "the reference's AngularDistributionFunction.run_calculator executed "
                             "over a NumPy stand-in for the TensorFlow leaf ops",
                   "cases": out}, fh)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
