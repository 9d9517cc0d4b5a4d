"""
The equilibrium calculator as a differentiable torch function.

The reference differentiates ANY Python function of the equilibrium state through HIPS autograd's tape
(src/dfdm/optimization.py:61, goals/goal.py:40-68, losses/losses.py:22-31): a user may subclass ``Goal`` or
``LossTerm`` and the optimiser still gets a gradient.  The library goals and loss terms of this package never touch a
tape - their cotangents are closed forms inside the CUDA kernels - so user-written ones need another route to the
GPU path: ``EquilibriumFunction`` is a ``torch.autograd.Function`` whose forward is ``dfdm_forward`` and whose backward
is ``dfdm_vjp`` (the closed-form adjoint for arbitrary cotangents of xyz, residuals, vectors, lengths and forces).
Anything written with torch operations on the state it returns is differentiated by torch down to the state, and by
the CUDA adjoint from the state to the force densities.

    q = torch.tensor(q0, dtype=torch.float64, device="cuda", requires_grad=True)
    state = equilibrium(q, model)                 # EquilibriumState of torch tensors, on the GPU
    value = my_function(state)                    # any differentiable torch code
    value.backward()                              # q.grad through dfdm_vjp

``Loss`` uses it by itself for user-defined goals and loss terms (losses.py).
"""
import numpy as np

from dfdm_b200.engine import Evaluator
from dfdm_b200.equilibrium.state import EquilibriumState

__all__ = ["EquilibriumFunction", "equilibrium"]


def _function():
    import torch

    class EquilibriumFunction(torch.autograd.Function):
        """(q[B, E]) -> (xyz[B, n, 3], residuals[B, n, 3], vectors[B, E, 3], lengths[B, E], forces[B, E])"""

        @staticmethod
        def forward(ctx, q, evaluator):
            out = evaluator.forward(q.detach(), rescue=True)
            Evaluator.raise_on_status(out["status"])
            ctx.evaluator = evaluator
            ctx.save_for_backward(q.detach())
            return tuple(out[k] for k in Evaluator.STATE)

        @staticmethod
        def backward(ctx, g_xyz, g_res, g_vec, g_len, g_for):
            q, = ctx.saved_tensors
            cots = [None if g is None else g.contiguous() for g in (g_xyz, g_res, g_vec, g_len, g_for)]
            dq, status = ctx.evaluator.vjp(q, *cots, rescue=True)
            Evaluator.raise_on_status(status)
            return dq, None

    return EquilibriumFunction


_FUNCTION = None


def EquilibriumFunction():
    """The ``torch.autograd.Function`` class (created on first use: torch is imported lazily, as everywhere in the package)."""
    global _FUNCTION
    if _FUNCTION is None:
        _FUNCTION = _function()
    return _FUNCTION


def equilibrium(q, model):
    """
    Differentiable ``EquilibriumModel.__call__``: ``q`` is a torch tensor ``[E]`` or ``[B, E]`` (moved to the model's GPU);
    returns an ``EquilibriumState`` whose arrays are torch tensors attached to the autograd graph of ``q``.
    """
    import torch
    ev = model.evaluator
    if not isinstance(q, torch.Tensor):
        q = torch.as_tensor(np.asarray(q, dtype=np.float64))
    q = q.to(device=ev.device, dtype=torch.float64)
    single = q.dim() == 1
    qb = q[None, :] if single else q
    xyz, res, vec, length, force = EquilibriumFunction().apply(qb.contiguous(), ev)
    if single:
        xyz, res, vec, length, force = xyz[0], res[0], vec[0], length[0], force[0]
    return EquilibriumState(xyz=xyz, residuals=res, vectors=vec, lengths=length, forces=force, force_densities=q)
