"""Generate tests/golden/consumer_golden.json by EXECUTING THE REFERENCE'S OWN CODE for the three
calculators that consume RDF output:

    mdsuite/calculators/coordination_number_calculation.py
    mdsuite/calculators/potential_of_mean_force.py
    mdsuite/calculators/kirkwood_buff_integrals.py
    mdsuite/utils/meta_functions.py            (apply_savgol_filter, golden_section_search)
    mdsuite/utils/units.py                     (boltzmann_constant, golden_ratio)

loaded unmodified from /root/reference with stand-ins for what they import but do not need (bokeh,
the calculator base class, the SQL scheme).  These calculators are NumPy / SciPy only, so — unlike
the RDF itself — nothing is restated: the fixtures are the reference's numbers.

Runs only where /root/reference exists.  Re-run:  python tests/golden/make_consumer_golden.py
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


def synthetic_rdf(case):
    """A liquid-like g(r) on the RDF calculator's grid (x in nm, index 0 singular)."""
    rng = np.random.Generator(np.random.Philox(case["seed"]))
    nbins, cutoff_nm = case["nbins"], case["cutoff_nm"]
    x = np.linspace(0.0, cutoff_nm, nbins)
    out = {}
    for k, pair in enumerate(case["pairs"]):
        r0 = case["sigma_nm"] * (1.0 + 0.15 * k)
        osc = 1.0 + case["amplitude"] * np.exp(-(x - r0) / (2.5 * r0)) * np.cos(
            2 * np.pi * (x - r0) / (0.95 * r0))
        core = 1.0 / (1.0 + np.exp(-(x - 0.92 * r0) / (0.02 * r0)))
        y = core * osc + rng.normal(0.0, case["noise"], nbins)
        y = np.maximum(y, 1e-6)
        y[0] = np.nan
        out[pair] = {"x": x.tolist(), "y": y.tolist()}
    return out


CASES = [
    dict(name="two_species_melt", seed=301, nbins=400, cutoff_nm=1.2, sigma_nm=0.26, amplitude=1.6,
         noise=0.004, pairs=["Na_Na", "Na_Cl", "Cl_Cl"], species={"Na": 500, "Cl": 500},
         box=[31.5, 31.5, 31.5], temperature=1400.0, shells=2, savgol=(2, 17)),
    dict(name="single_species_fine_grid", seed=302, nbins=700, cutoff_nm=1.5, sigma_nm=0.34,
         amplitude=1.2, noise=0.001, pairs=["Ar_Ar"], species={"Ar": 4000}, box=[55.0, 55.0, 55.0],
         temperature=94.4, shells=1, savgol=(3, 31)),
]


def load_reference():
    stubs = {}
    for name in ("bokeh", "bokeh.models", "bokeh.models.ranges", "bokeh.plotting", "GPUtil", "psutil",
                 "tensorflow"):
        stubs[name] = mock.MagicMock()
    for pkg in ("mdsuite", "mdsuite.calculators", "mdsuite.utils", "mdsuite.database"):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REF, *pkg.split(".")[1:])]
        stubs[pkg] = m
    calc = types.ModuleType("mdsuite.calculators.calculator")
    calc.call = lambda f: f
    calc.Calculator = type("Calculator", (), {"__init__": lambda self, **kw: None})
    stubs["mdsuite.calculators.calculator"] = calc
    scheme = types.ModuleType("mdsuite.database.scheme")
    scheme.Computation = type("Computation", (), {})
    stubs["mdsuite.database.scheme"] = scheme
    exc = types.ModuleType("mdsuite.utils.exceptions")
    exc.CannotPerformThisAnalysis = type("CannotPerformThisAnalysis", (Exception,), {})
    exc.NoGPUInSystem = type("NoGPUInSystem", (Exception,), {})
    stubs["mdsuite.utils.exceptions"] = exc
    sys.modules.update(stubs)

    def load(modname, relpath):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, relpath))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod

    units = load("mdsuite.utils.units", "utils/units.py")
    sys.modules["mdsuite"].utils = sys.modules["mdsuite.utils"]
    sys.modules["mdsuite.utils"].units = units
    load("mdsuite.utils.meta_functions", "utils/meta_functions.py")
    return (load("mdsuite.calculators.coordination_number_calculation",
                 "calculators/coordination_number_calculation.py"),
            load("mdsuite.calculators.potential_of_mean_force",
                 "calculators/potential_of_mean_force.py"),
            load("mdsuite.calculators.kirkwood_buff_integrals",
                 "calculators/kirkwood_buff_integrals.py"), units)


def run(module, cls_name, case, rdf, units, with_shells=True):
    cls = getattr(module, cls_name)
    calc = object.__new__(cls)
    real = units.units_dict["real"] if hasattr(units, "units_dict") else None
    length = 1e-10
    calc.experiment = SimpleNamespace(
        volume=float(np.prod(case["box"])), temperature=case["temperature"],
        units=SimpleNamespace(volume=length ** 3, length=length),
        species={n: SimpleNamespace(n_particles=c) for n, c in case["species"].items()})
    calc.rdf_data = SimpleNamespace(data_dict=rdf)
    fields = dict(savgol_order=case["savgol"][0], savgol_window_length=case["savgol"][1],
                  number_of_bins=case["nbins"], number_of_configurations=100,
                  cutoff=case["cutoff_nm"] * 10.0)
    if with_shells:
        fields["number_of_shells"] = case["shells"]
    calc.args = module.Args(**fields)
    calc.result_keys = []
    if with_shells:
        prefix = "CN" if cls_name == "CoordinationNumbers" else "POMF"
        for i in range(case["shells"]):
            calc.result_keys += [f"{prefix}_{i + 1}", f"{prefix}_{i + 1}_error"]
    calc.result_series_keys = {"CoordinationNumbers": ["r", "cn"],
                               "PotentialOfMeanForce": ["r", "pomf"],
                               "KirkwoodBuffIntegral": ["r", "kb_integral"]}[cls_name]
    if cls_name == "CoordinationNumbers":
        calc._compute_nm_volume()
    queued = []
    calc.queue_data = lambda data, subjects: queued.append({"subjects": list(subjects), "data": data})
    with np.errstate(all="ignore"):
        calc.run_calculator()
    return queued


def encode(queued):
    out = []
    for q in queued:
        data = {}
        for k, v in q["data"].items():
            data[k] = [float(x).hex() for x in v] if isinstance(v, list) else float(v).hex()
        out.append({"subjects": q["subjects"], "data": data})
    return out


def main():
    cn_mod, pomf_mod, kb_mod, units = load_reference()
    cases = []
    for case in CASES:
        rdf = synthetic_rdf(case)
        rec = dict(case)
        rec["coordination_numbers"] = encode(run(cn_mod, "CoordinationNumbers", case, rdf, units))
        rec["potential_of_mean_force"] = encode(run(pomf_mod, "PotentialOfMeanForce", case, rdf, units))
        rec["kirkwood_buff"] = encode(run(kb_mod, "KirkwoodBuffIntegral", case, rdf, units, False))
        cases.append(rec)
        print(case["name"], {q["subjects"][0] + "_" + q["subjects"][1]:
                             float.fromhex(q["data"]["CN_1"]) for q in rec["coordination_numbers"]})
    path = os.path.join(HERE, "consumer_golden.json")
    with open(path, "w") as fh:
        json.dump({"generator": "tests/golden/make_consumer_golden.py",
                   "source": "the reference's CoordinationNumbers / PotentialOfMeanForce / "
                             "KirkwoodBuffIntegral run_calculator, executed unmodified",
                   "cases": cases}, fh)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
