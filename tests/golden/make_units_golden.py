"""Generate tests/golden/units_golden.json from the reference's own unit systems
(mdsuite/utils/units.py, loaded unmodified).  Runs only where /root/reference exists.
Re-run:  python tests/golden/make_units_golden.py"""
import dataclasses
import importlib.util
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    spec = importlib.util.spec_from_file_location("reference_units",
                                                  "/root/reference/mdsuite/utils/units.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = {"generator": "tests/golden/make_units_golden.py",
           "systems": {name: dict(dataclasses.asdict(u), volume=u.volume)
                       for name, u in mod.units_dict.items()},
           "boltzmann_constant": mod.boltzmann_constant, "golden_ratio": mod.golden_ratio}
    with open(os.path.join(HERE, "units_golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)
        fh.write("\n")
    print({k: v["length"] for k, v in out["systems"].items()})


if __name__ == "__main__":
    main()
