"""``import mdsuite`` is the reference's import name; here it is an alias of mdsuite_b200."""
import numpy as np


def test_reference_import_name_resolves_to_this_package(tmp_path):
    import mdsuite as mds
    import mdsuite_b200
    import mdsuite.calculators.radial_distribution_function as by_alias
    import mdsuite_b200.calculators.radial_distribution_function as by_name
    from mdsuite.database.simulation_database import Database
    from mdsuite_b200.database import simulation_database
    assert mds.Project is mdsuite_b200.Project and mds.Experiment is mdsuite_b200.Experiment
    assert by_alias is by_name and Database is simulation_database.Database
    assert by_name.__name__ == "mdsuite_b200.calculators.radial_distribution_function"
    project = mds.Project("alias", storage_path=str(tmp_path))
    exp = project.add_experiment("NaCl", units="real")
    pos = np.random.default_rng(0).random((2, 6, 3)).astype(np.float32) * 5
    exp.add_positions(pos, {"Na": 3, "Cl": 3}, np.array([5.0, 5.0, 5.0]))
    assert callable(exp.run.RadialDistributionFunction)
    try:
        import mdsuite.not_a_module  # noqa: F401
    except ImportError:
        pass
    else:
        raise AssertionError("unknown submodules must not resolve")
