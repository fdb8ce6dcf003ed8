"""Generate tests/golden/reference_golden.json by EXECUTING THE REFERENCE'S OWN CODE.

Runs only where /root/reference exists (this container; not the GPU box).  The reference package
cannot be imported as a whole (tensorflow, h5py, sqlalchemy, ... are not installed), so this script
loads just the files on the RDF path

    mdsuite/calculators/radial_distribution_function.py      (the calculator, unmodified)
    mdsuite/utils/linalg.py                                  (get_partial_triu_indices, ...)
    mdsuite/utils/meta_functions.py                          (join_path, split_array)

from where they lie, under
  * a NumPy-backed stand-in for the ~20 TensorFlow leaf ops these files call (``TfShim`` below),
  * stubs for the data plumbing of the parent class (batches come from an in-memory array).

What this PINS (the reference's own Python runs, statement by statement): check_input defaults,
configuration sampling, the batch / atom-minibatch loops, get_partial_triu_indices, the per-pair
index masks with their strict '>' (quirk Q1), the accumulation, particles_list, the prefactor, the
ideal-gas correction with its > L/2 branches, the x axis and unit scaling, the queued payload.
What it CANNOT pin: TensorFlow's arithmetic itself.  The shim computes the leaf ops in float32
NumPy exactly as SURVEY.md §8c states the contract (every op rounded on its own; histogram bins by
the CPU kernel formula in double) — the same statement the oracle makes.  So: control flow, masks
and normalisation pinned; float32 leaf arithmetic still a restatement.

Re-run:  python tests/golden/make_reference_golden.py
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
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/mdsuite"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import make_golden  # noqa: E402  (seeded inputs shared with the self-generated vectors)


# ---- NumPy stand-in for the TensorFlow calls on the RDF path ------------------------------------
class _Dataset:
    def __init__(self, tensor):
        self.tensor = tensor
        self.size = None

    def batch(self, size):
        self.size = int(size)
        return self

    def prefetch(self, _n):
        return self

    def __iter__(self):
        for lo in range(0, self.tensor.shape[0], self.size):
            yield self.tensor[lo:lo + self.size]


class _IntTensor:
    """The int32 histogram tensors of the reference (tf.zeros(..., tf.int32) and the result of
    tf.histogram_fixed_width): not an ndarray, so that ``ndarray += tensor`` defers to the tensor's
    __radd__ as it does with a real EagerTensor (__array_priority__ 100) and the global accumulator
    becomes an int32 tensor with a .numpy() method (reference :884-885, :358; quirk Q5)."""
    __array_priority__ = 100

    def __init__(self, values):
        self.values = np.asarray(values).astype(np.int32)

    def numpy(self):
        return self.values

    def __array__(self, dtype=None, copy=None):
        return self.values if dtype is None else self.values.astype(dtype)

    def __add__(self, other):
        return _IntTensor(self.values + np.asarray(other).astype(np.int32))

    __radd__ = __add__


def _histogram_fixed_width(values, value_range, nbins):
    """TF CPU kernel (histogram_op.cc) as SURVEY.md §8c restates it."""
    lo, hi = np.float32(value_range[0]), np.float32(value_range[1])
    nbins = int(nbins)
    step = np.float64(hi - lo) / np.float64(nbins)
    v = np.asarray(values, dtype=np.float32)
    clipped = np.maximum(v, lo) - lo
    idx = np.minimum(clipped.astype(np.float64) / step, np.float64(nbins - 1)).astype(np.int32)
    return _IntTensor(np.bincount(idx, minlength=nbins))


def _norm(x, axis=-1):
    assert axis == -1 and x.shape[-1] == 3 and x.dtype == np.float32
    sq = x * x
    return np.sqrt((sq[..., 0] + sq[..., 1]) + sq[..., 2])


def _band_part(mat, lower, upper):
    m, n = mat.shape
    i, j = np.indices((m, n))
    keep = np.ones((m, n), dtype=bool)
    if lower >= 0:
        keep &= (i - j) <= lower
    if upper >= 0:
        keep &= (j - i) <= upper
    return mat & keep


def make_tf_shim():
    tf = types.ModuleType("tensorflow")
    tf.float32, tf.float64, tf.int32, tf.int64, tf.bool = np.float32, np.float64, np.int32, np.int64, bool
    tf.Tensor = np.ndarray
    tf.dtypes = SimpleNamespace(int32=np.int32, float32=np.float32)  # default arguments elsewhere
    tf.function = lambda *a, **k: (a[0] if a and callable(a[0]) else (lambda f: f))
    tf.cast = lambda x, dtype: np.asarray(x).astype(dtype)
    tf.shape = lambda x: np.asarray(x).shape
    # scalars become NumPy scalars, which are immutable like tensors (`stop += n` must not alias)
    tf.constant = lambda v, dtype=None: (np.asarray(v, dtype=dtype) if np.ndim(v) else np.int32(v))
    tf.zeros = lambda shape, dtype=np.float32: _IntTensor(np.zeros(int(shape), dtype=np.int32))  # only use: histograms
    tf.ones = lambda shape, dtype=np.float32: np.ones([int(s) for s in shape], dtype=dtype)
    tf.stack = lambda vals, axis=0: np.stack([np.asarray(v) for v in vals], axis=axis)
    tf.concat = lambda vals, axis=0: np.concatenate(vals, axis=axis)
    tf.transpose = lambda x: np.transpose(x)
    tf.where = lambda cond: np.argwhere(cond)
    tf.gather = lambda params, indices, axis=0: np.take(params, indices, axis=axis)
    tf.less = lambda a, b: np.less(a, b)
    tf.boolean_mask = lambda t, mask, axis=None: np.asarray(t)[np.asarray(mask)]
    tf.histogram_fixed_width = lambda values, value_range, nbins: _histogram_fixed_width(values, value_range, nbins)
    tf.math = SimpleNamespace(rint=np.rint)
    tf.linalg = SimpleNamespace(norm=_norm, band_part=_band_part)
    tf.data = SimpleNamespace(Dataset=SimpleNamespace(from_tensor_slices=_Dataset), AUTOTUNE=-1)
    return tf


def load_reference_rdf_class():
    """Import the three reference files with stand-ins for everything they import but do not need."""
    tf = make_tf_shim()
    stubs = {"tensorflow": tf}
    for name in ("GPUtil", "psutil", "tensorflow_probability"):
        stubs[name] = mock.MagicMock()
    for pkg in ("mdsuite", "mdsuite.calculators", "mdsuite.utils", "mdsuite.database"):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REF, *pkg.split(".")[1:])]
        stubs[pkg] = m
    calc = types.ModuleType("mdsuite.calculators.calculator")
    calc.call = lambda f: f
    stubs["mdsuite.calculators.calculator"] = calc
    traj = types.ModuleType("mdsuite.calculators.trajectory_calculator")

    class TrajectoryCalculator:  # the parent's data plumbing, replaced by in-memory batches
        def __init__(self, **kwargs):
            pass
    traj.TrajectoryCalculator = TrajectoryCalculator
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
    rdf_mod = load("mdsuite.calculators.radial_distribution_function",
                   "calculators/radial_distribution_function.py")
    return rdf_mod


def run_reference(rdf_mod, pos_fm, species, box, number_of_bins, cutoff, number_of_configurations,
                  minibatch, start=0, stop=None, frames_per_batch=1, length_unit=1e-10):
    """The reference's run_calculator over in-memory positions.  species: {name: n_particles} in
    storage order; pos_fm: (F, N, 3) float32, atoms species-contiguous."""
    cls = rdf_mod.RadialDistributionFunction
    calc = object.__new__(cls)
    offs = np.concatenate([[0], np.cumsum(list(species.values()))])
    calc.experiment = SimpleNamespace(
        box_array=[float(b) for b in box], number_of_configurations=pos_fm.shape[0],
        species={n: SimpleNamespace(n_particles=c) for n, c in species.items()}, molecules={},
        volume=float(np.prod(box)), units=SimpleNamespace(length=length_unit))
    calc.args = rdf_mod.Args(number_of_bins=number_of_bins, cutoff=cutoff, start=start, stop=stop,
                             atom_selection=np.s_[:], data_range=1, correlation_time=1,
                             molecules=False, species=None,
                             number_of_configurations=number_of_configurations)
    calc.rdf_minibatch = minibatch
    calc.dtype = rdf_mod.tf.float32
    calc.loaded_property = SimpleNamespace(name="Positions")
    calc.result_series_keys = ["x", "y"]
    calc.tqdm_limit = 0
    calc.override_n_batches = None
    calc.minibatch = False
    calc.selected_species = None
    queued = []
    calc.queue_data = lambda data, subjects: queued.append((list(subjects), data))

    def _prepare_managers(path_list):
        calc.batch_size = frames_per_batch
        calc.memory_manager = SimpleNamespace()
    calc._prepare_managers = _prepare_managers

    def get_batch_dataset(subject_list, loop_array, correct):
        for frames in loop_array:  # the DataManager generator: {b"Na/Positions": (n, C, 3) float64}
            yield {str.encode(f"{name}/Positions"):
                   pos_fm[frames][:, offs[k]:offs[k + 1]].transpose(1, 0, 2).astype(np.float64)
                   for k, name in enumerate(subject_list)}
    calc.get_batch_dataset = get_batch_dataset

    raw = {}
    real_normalise = calc._calculate_radial_distribution_functions

    def capture_then_normalise():
        raw.update({k: np.array(v.numpy()).astype(np.int64) for k, v in calc.rdf.items()})
        real_normalise()
    calc._calculate_radial_distribution_functions = capture_then_normalise
    with np.errstate(all="ignore"):
        calc.run_calculator()
    return calc, raw, queued


CASES = [dict(c) for c in make_golden.CASES] + [
    dict(name="q1_demo_5_4", kind="uniform", seed=201, counts=[5, 4], frames=3, box=[10.0, 10.0, 10.0],
         cutoff=9.0, nbins=20, minibatch=2),
    dict(name="defaults_cutoff_and_bins", kind="uniform", seed=202, counts=[30, 20], frames=6,
         box=[8.0, 8.0, 8.0], cutoff=None, nbins=None, minibatch=-1, number_of_configurations=4),
    dict(name="cutoff_into_arctan_branch", kind="uniform", seed=203, counts=[25], frames=2,
         box=[4.0, 4.0, 4.0], cutoff=3.3, nbins=66, minibatch=10, frames_per_batch=2),
]


def main():
    rdf_mod = load_reference_rdf_class()
    out = []
    for case in CASES:
        pos, box = make_golden.case_inputs({**case, "cutoff": case["cutoff"] or 1.0,
                                            "nbins": case["nbins"] or 1})
        names = ["Na", "Cl", "K"][:len(case["counts"])]
        species = dict(zip(names, case["counts"]))
        nconf = case.get("number_of_configurations", case["frames"])
        calc, raw, queued = run_reference(rdf_mod, pos, species, box, case["nbins"], case["cutoff"],
                                          nconf, case["minibatch"],
                                          frames_per_batch=case.get("frames_per_batch", 1))
        rec = dict(case)
        rec["species"] = names
        rec["resolved"] = {"cutoff": float(calc.args.cutoff), "nbins": int(calc.args.number_of_bins),
                           "stop": int(calc.args.stop),
                           "number_of_configurations": int(calc.args.number_of_configurations),
                           "sample_configurations": [int(v) for v in calc.sample_configurations],
                           "key_list": list(calc.key_list)}
        rec["histograms"] = {k: [int(v) for v in raw[k]] for k in calc.key_list}
        rec["results"] = [{"subjects": s, "x": [float(v).hex() for v in d["x"]],
                           "y": [float(v).hex() for v in d["y"]]} for s, d in queued]
        out.append(rec)
        print(case["name"], {k: int(np.sum(v)) for k, v in raw.items()})
    path = os.path.join(HERE, "reference_golden.json")
    with open(path, "w") as fh:
        json.dump({"generator": "tests/golden/make_reference_golden.py",
                   "source": "the reference's run_calculator executed over a NumPy shim of its "
                             "TensorFlow leaf ops (see the generator's docstring for what this pins)",
                   "cases": out}, fh)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
