"""
``import dfdm`` compatibility: scripts written against the reference import ``dfdm.equilibrium``,
``dfdm.goals`` ...; ``install()`` registers this package's mirrors under those names so they run
unchanged (apart from COMPAS-only helpers such as viewers and mesh builders).

    import dfdm_b200.compat; dfdm_b200.compat.install()
    from dfdm.equilibrium import fdm            # -> dfdm_b200.equilibrium.fdm
"""
import importlib
import sys

SUBMODULES = ("datastructures", "equilibrium", "goals", "losses", "constraints", "optimization")


def install(name="dfdm"):
    import dfdm_b200
    if name in sys.modules and sys.modules[name] is not dfdm_b200:
        raise ImportError("a different package named %r is already imported" % name)
    sys.modules[name] = dfdm_b200
    for sub in SUBMODULES:
        module = importlib.import_module("dfdm_b200." + sub)
        sys.modules[name + "." + sub] = module
        setattr(dfdm_b200, sub, module)
    return dfdm_b200
