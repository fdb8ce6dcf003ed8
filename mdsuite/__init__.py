"""``import mdsuite as mds`` — the reference's import name, served by mdsuite_b200.

A user of zincware/MDSuite writes ``import mdsuite as mds; mds.Project(...)`` and
``project.run.RadialDistributionFunction(...)``; this alias package makes those lines work
unchanged on top of the B200-native implementation.  Submodules resolve to their mdsuite_b200
counterparts (``mdsuite.calculators.radial_distribution_function`` is
``mdsuite_b200.calculators.radial_distribution_function``); there is one copy of every module.
Only what SURVEY.md §8 scopes exists — asking for anything else raises ImportError as usual."""
import importlib
import importlib.abc
import importlib.util
import sys

import mdsuite_b200
from mdsuite_b200 import *  # noqa: F401,F403
from mdsuite_b200 import Experiment, Project, __version__, config  # noqa: F401


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """mdsuite.<x> -> mdsuite_b200.<x>, the same module object under both names."""

    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith("mdsuite."):
            return None
        real = "mdsuite_b200." + fullname[len("mdsuite."):]
        try:
            if importlib.util.find_spec(real) is None:
                return None
        except (ImportError, ValueError):
            return None
        return importlib.util.spec_from_loader(fullname, self)

    def create_module(self, spec):
        return importlib.import_module("mdsuite_b200." + spec.name[len("mdsuite."):])

    def exec_module(self, module):  # already executed under its real name
        pass


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
