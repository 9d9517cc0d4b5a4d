"""
Golden results for the example scripts: every ``examples/*.py`` port is executed HERE with the reference's own
``constrained_fdm`` / optimisers / goals / losses / constraints (/root/reference/src/dfdm, unmodified, under the shims
of make_golden.py) substituted for the dfdm_b200 modules of the same names.  The problem definitions (geometry,
goals, bounds, optimiser settings) are therefore literally the ones the GPU examples run; what is recorded is the
optimal force densities, the equilibrium coordinates and the loss of every ``constrained_fdm`` call, in call order.

    python tests/golden/make_example_golden.py          # writes tests/golden/examples.npz (build container only)

``tests/test_gpu_examples.py`` compares the GPU runs of the same scripts with these arrays.
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples"))

import make_golden as MG                                   # noqa: E402  (shims for autograd / compas)

# example -> keyword arguments of its main(); the GPU test passes the same
RUNS = {
    "arches": dict(vertical_comps=(0.2, 0.8), verbose=False),
    "pillow": dict(maxiter=40, verbose=False),
    "pringle": dict(maxiter=60, verbose=False),
    "dome_constraints": dict(maxiter=40, verbose=False),
    "monkey_saddle_constraints": dict(maxiter=40, verbose=False),
    "dome": dict(maxiter=300, tol=1e-3, verbose=False),
    "monkey_saddle": dict(maxiter=200, tol=1e-3, verbose=False),
}


def main():
    assert os.path.isdir(MG.REFERENCE_SRC), "the reference tree is only present in the build container"
    MG._install_shims()
    sys.path.insert(0, MG.REFERENCE_SRC)
    import dfdm.equilibrium as REQ
    import dfdm.goals as RG
    import dfdm.losses as RL
    import dfdm.constraints as RC
    import dfdm.optimization as RO
    from dfdm.datastructures import FDNetwork as RefFDNetwork

    calls = []

    def to_ref(network):
        return network if isinstance(network, RefFDNetwork) else RefFDNetwork.from_data(network.data)

    def fdm(network, *args, **kwargs):
        MG.MODE["value"] = "numpy"
        return REQ.fdm(to_ref(network))

    import torch

    class TorchLoss:
        """The reference's Loss, callable with SciPy's NumPy iterates while its model holds torch arrays."""
        def __init__(self, loss):
            self.loss = loss

        def __call__(self, q, model):
            if isinstance(q, torch.Tensor):
                return self.loss(q, model)
            return float(self.loss(torch.as_tensor(np.asarray(q, dtype=np.float64)), model))

    class TorchConstraint:
        def __init__(self, constraint):
            self.constraint = constraint
            self.bound_low, self.bound_up = constraint.bound_low, constraint.bound_up

        def __call__(self, q, model):
            if isinstance(q, torch.Tensor):
                return self.constraint(q, model)
            value = self.constraint(torch.as_tensor(np.asarray(q, dtype=np.float64)), model)
            return np.atleast_1d(np.asarray(value, dtype=np.float64))

    def constrained_fdm(network, optimizer, loss, bounds=(None, None), constraints=None, maxiter=100, tol=1e-6, callback=None):
        net = to_ref(network)
        MG.MODE["value"] = "torch"                           # one model serves loss (values) and autograd.grad (torch tape)
        try:
            # reference fdm.py:36-41, with the final _fdm in NumPy mode
            wrapped = [TorchConstraint(c) for c in constraints] if constraints else constraints
            q_opt = optimizer.minimize(net, TorchLoss(loss), bounds, wrapped, maxiter, tol, callback=callback)
        finally:
            MG.MODE["value"] = "numpy"
        q = np.asarray(q_opt, dtype=np.float64)
        model = REQ.EquilibriumModel(net)
        out = REQ.network_updated(net, model(q))
        xyz = np.asarray([out.node_coordinates(k) for k in out.nodes()], dtype=np.float64)
        calls.append({"q": q, "xyz": xyz, "loss": float(loss(q, model))})
        return out

    eq = types.ModuleType("dfdm_b200.equilibrium")
    eq.fdm, eq.constrained_fdm = fdm, constrained_fdm
    eq.EquilibriumModel, eq.network_update, eq.network_updated = REQ.EquilibriumModel, REQ.network_update, REQ.network_updated
    aliases = {"dfdm_b200.equilibrium": eq, "dfdm_b200.goals": RG, "dfdm_b200.losses": RL, "dfdm_b200.constraints": RC,
               "dfdm_b200.optimization": RO}
    import dfdm_b200                                        # noqa: F401  the real package (workloads, datastructures stay ours)
    import dfdm_b200.workloads                              # noqa: F401
    out = {}
    for name, kwargs in RUNS.items():
        saved = {k: sys.modules.get(k) for k in aliases}
        sys.modules.update(aliases)
        del calls[:]
        try:
            spec = importlib.util.spec_from_file_location("golden_example_" + name, os.path.join(ROOT, "examples", name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            try:
                mod.main(**kwargs)
            except Exception as exc:                         # reporting that needs the batched API comes after the solves
                print("  (%s: post-processing skipped under the reference: %s: %s)" % (name, type(exc).__name__, str(exc)[:80]))
        finally:
            for k, v in saved.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v
        for i, c in enumerate(calls):
            for key, val in c.items():
                out["%s.%d.%s" % (name, i, key)] = np.asarray(val)
        print(name, "->", len(calls), "constrained_fdm calls; losses", [round(c["loss"], 6) for c in calls], flush=True)
    np.savez_compressed(os.path.join(HERE, "examples.npz"), **out)
    print("wrote", os.path.join(HERE, "examples.npz"))


if __name__ == "__main__":
    main()
