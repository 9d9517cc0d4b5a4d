"""
Host-side handles over the C ABI: ``Plan`` (compiled topology), ``GoalProgram`` (flattened goals and
loss terms) and ``Evaluator`` (batched forward / loss+gradient / VJP on one GPU).

PyTorch is used for plumbing only — device buffers, the current stream, pinned host memory; every
number is produced by the kernels in ``csrc/``.  Without a CUDA device the evaluation methods raise:
there is no CPU fallback.
"""
import ctypes as C

import numpy as np

from dfdm_b200 import _lib

__all__ = ["Plan", "GoalProgram", "Evaluator", "LinAlgError"]

LinAlgError = np.linalg.LinAlgError


def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Plan:
    """
    Compiled topology: replaces ``EquilibriumStructure`` (/root/reference/src/dfdm/equilibrium/structure.py:11-106)
    and the per-call slicing of model.py:44-47.  Immutable; shareable across evaluators and devices.
    """

    def __init__(self, n, edges, is_fixed):
        edges = np.ascontiguousarray(np.asarray(edges, dtype=np.int32).reshape(-1, 2))
        is_fixed = np.ascontiguousarray(np.asarray(is_fixed).astype(np.uint8))
        if is_fixed.shape[0] != int(n):
            raise ValueError("is_fixed must have one entry per node")
        self._handle = C.c_void_p()
        _lib.check(_lib.lib.dfdm_plan_create(int(n), edges.shape[0], _np_ptr(edges), _np_ptr(is_fixed), C.byref(self._handle)))
        self.edges = edges
        info = _lib.PlanInfo()
        _lib.check(_lib.lib.dfdm_plan_get_info(self._handle, C.byref(info)))
        self.info = {name: getattr(info, name) for name, _ in _lib.PlanInfo._fields_ if name != "reserved"}
        self.n, self.n_edges, self.n_free, self.n_fixed = info.n, info.n_edges, info.n_free, info.n_fixed
        self.nnz = info.nnz
        self._arrays = {}

    def __del__(self):
        handle = getattr(self, "_handle", None)
        if handle is not None and handle.value:
            try:
                _lib.lib.dfdm_plan_destroy(handle)
            except Exception:
                pass
            self._handle = None

    @property
    def handle(self):
        return self._handle

    def array(self, which):
        if which not in self._arrays:
            self._arrays[which] = _lib.plan_array(self._handle, which)
        return self._arrays[which]

    free = property(lambda self: self.array(_lib.ARR_FREE))
    fixed = property(lambda self: self.array(_lib.ARR_FIXED))
    freefixed = property(lambda self: self.array(_lib.ARR_FREEFIXED))
    rowptr = property(lambda self: self.array(_lib.ARR_ROWPTR))
    colidx = property(lambda self: self.array(_lib.ARR_COLIDX))
    edge_slots = property(lambda self: self.array(_lib.ARR_EDGE_SLOTS).reshape(-1, 4))
    solver_perm = property(lambda self: self.array(_lib.ARR_SOLVER_PERM))

    @property
    def auto_solver(self):
        return "direct" if self.info["auto_solver"] == _lib.SOLVER_DIRECT else "pcg"


class GoalProgram:
    """
    Flat goal / loss table (replaces Loss -> LossTerm -> [Goal], losses.py:10-87, goals/helpers.py:19-37).

    ``goals``: iterable of ``(kind name, element row index | None, term index, weight, params[<=6])``.
    ``terms``: iterable of ``(term kind name, alpha)``.  ``l2_alpha``: sum of L2Regularizer alphas.
    """

    def __init__(self, plan, goals, terms, l2_alpha=0.0):
        goals = list(goals)
        terms = list(terms)
        G = len(goals)
        kind = np.zeros(G, dtype=np.int32)
        elem = np.zeros(G, dtype=np.int32)
        term = np.zeros(G, dtype=np.int32)
        weight = np.zeros(G, dtype=np.float64)
        params = np.zeros((G, _lib.GOAL_NPARAMS), dtype=np.float64)
        for g, (k, e, t, w, p) in enumerate(goals):
            kind[g] = _lib.GOAL_KINDS[k]
            elem[g] = -1 if e is None else int(e)
            term[g] = int(t)
            weight[g] = float(w)
            p = np.asarray(p, dtype=np.float64).ravel()
            params[g, :p.size] = p
        tkind = np.asarray([_lib.TERM_KINDS[k] for k, _ in terms], dtype=np.int32)
        talpha = np.asarray([float(a) for _, a in terms], dtype=np.float64)
        self.plan = plan
        self.n_goals, self.n_terms, self.l2_alpha = G, len(terms), float(l2_alpha)
        self._handle = C.c_void_p()
        _lib.check(_lib.lib.dfdm_goalprog_create(plan.handle, G, _np_ptr(kind), _np_ptr(elem), _np_ptr(term), _np_ptr(weight),
                                                 _np_ptr(params), len(terms), _np_ptr(tkind), _np_ptr(talpha),
                                                 float(l2_alpha), C.byref(self._handle)))

    def __del__(self):
        handle = getattr(self, "_handle", None)
        if handle is not None and handle.value:
            try:
                _lib.lib.dfdm_goalprog_destroy(handle)
            except Exception:
                pass
            self._handle = None

    @property
    def handle(self):
        return self._handle


class Evaluator:
    """
    Batched evaluation of one network on one GPU.

    ``loads``: n x 3 nodal loads in node order (model.py:18); ``xyz_fixed``: n_fixed x 3 support
    coordinates in fixed-node order (model.py:19).
    """

    STATE = ("xyz", "residuals", "vectors", "lengths", "forces")

    def __init__(self, plan, loads, xyz_fixed, device=None, solver="auto", pcg_rtol=0.0, pcg_maxiter=0, pcg_check_every=0,
                 pcg_precond="auto", mg_omega=0.0, mg_over=0.0, mg_w_levels=0, mg_over_w=0.0, flags=0, pcg_rtol_adjoint=0.0):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("dfdm_b200 evaluates on a CUDA device and there is none here (no CPU fallback)")
        self.torch = torch
        self.plan = plan
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.options = _lib.make_options(solver, pcg_maxiter, pcg_rtol, pcg_check_every, pcg_precond, mg_omega, mg_over, mg_w_levels, mg_over_w, flags, pcg_rtol_adjoint)
        loads = np.ascontiguousarray(np.asarray(loads, dtype=np.float64).reshape(plan.n, 3))
        xyz_fixed = np.ascontiguousarray(np.asarray(xyz_fixed, dtype=np.float64).reshape(plan.n_fixed, 3))
        self.loads_host, self.xyz_fixed_host = loads, xyz_fixed
        self.loads = torch.from_numpy(loads).to(self.device)
        self.xyz_fixed = torch.from_numpy(xyz_fixed).to(self.device) if plan.n_fixed else torch.zeros(1, 3, dtype=torch.float64, device=self.device)
        self._ws = None
        self._ws_batch = 0
        self._solver_name = solver
        self._twin = None

    # -- designs the PCG path cannot solve -----------------------------------------------------------------
    # CG needs a definite K: force densities of one sign.  The reference's LU (model.py:55) also solves the indefinite K of
    # a mixed-sign q, which the optimisers do wander into when q is unbounded (pillow.py:70, dome_constraints.py:80).  With
    # rescue=True the entry points below look at the verdicts (one device-to-host read) and hand the designs the PCG path
    # gave up on to the direct solver's global-memory variant (DFDM_SOLVER_DIRECT beyond the shared-memory limit) - slow,
    # exact, and flagged itself should a tiny pivot spoil it.  The API layer (EquilibriumModel, Loss, constraints) asks
    # for it; batched device loops (bench.py, the lock-step optimisers) do not and keep their hands off the host.

    def _failed_rows(self, status):
        if self._solver_name != "auto" or self.plan.auto_solver != "pcg":
            return None
        bad = self.torch.nonzero((status == _lib.STATUS_NOT_CONVERGED) | (status == _lib.STATUS_NONFINITE)).flatten()
        return bad if bad.numel() else None

    def _direct_twin(self):
        if self._twin is None:
            self._twin = Evaluator(self.plan, self.loads_host, self.xyz_fixed_host, device=self.device, solver="direct")
        return self._twin

    # -- plumbing ------------------------------------------------------------------------------------

    def _shapes(self, B):
        n, E = self.plan.n, self.plan.n_edges
        return {"xyz": (B, n, 3), "residuals": (B, n, 3), "vectors": (B, E, 3), "lengths": (B, E), "forces": (B, E)}

    def _workspace(self, B):
        if self._ws is None or self._ws_batch < B:
            nbytes = _lib.lib.dfdm_workspace_bytes(self.plan.handle, B, C.byref(self.options))
            if nbytes < 0:
                _lib.check(int(nbytes))
            self._ws = None
            self._ws = self.torch.empty(int(nbytes), dtype=self.torch.uint8, device=self.device)
            self._ws_batch = B
        return self._ws

    def _q(self, q):
        torch = self.torch
        if isinstance(q, torch.Tensor):
            t = q.to(device=self.device, dtype=torch.float64)
        else:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(q, dtype=np.float64))).to(self.device)
        if t.dim() == 1:
            t = t[None, :]
        if t.shape[1] != self.plan.n_edges:
            raise ValueError("q must have one entry per edge (%d), got %d" % (self.plan.n_edges, t.shape[1]))
        return t.contiguous()

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    # -- device entry points ---------------------------------------------------------------------------

    def forward(self, q, outputs=STATE, rescue=False):
        """EquilibriumModel.__call__ for a batch: dict of device tensors + ``status`` int32[B]."""
        torch = self.torch
        q = self._q(q)
        B = q.shape[0]
        shapes = self._shapes(B)
        out = {k: torch.empty(shapes[k], dtype=torch.float64, device=self.device) for k in outputs}
        status = torch.empty(B, dtype=torch.int32, device=self.device)
        ws = self._workspace(B)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib.dfdm_forward(self.plan.handle, B, _ptr(q), _ptr(self.loads), _ptr(self.xyz_fixed),
                                            *[_ptr(out.get(k)) for k in self.STATE], _ptr(status), _ptr(ws), ws.numel(),
                                            C.byref(self.options), self._stream()))
        out["status"] = status
        out["q"] = q
        bad = self._failed_rows(status) if rescue else None
        if bad is not None:
            sub = self._direct_twin().forward(q[bad], outputs=outputs)
            for k in tuple(outputs) + ("status",):
                out[k][bad] = sub[k]
        return out

    def loss_and_grad(self, prog, q, grad=True, outputs=(), rescue=False):
        """Loss.__call__ + grad(loss): ``loss[B]``, ``dq[B,E]`` (None if grad=False), ``status[B]``, optional state."""
        torch = self.torch
        q = self._q(q)
        B = q.shape[0]
        shapes = self._shapes(B)
        out = {k: torch.empty(shapes[k], dtype=torch.float64, device=self.device) for k in outputs}
        loss = torch.empty(B, dtype=torch.float64, device=self.device)
        dq = torch.empty_like(q) if grad else None
        status = torch.empty(B, dtype=torch.int32, device=self.device)
        ws = self._workspace(B)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib.dfdm_loss_and_grad(self.plan.handle, prog.handle, B, _ptr(q), _ptr(self.loads), _ptr(self.xyz_fixed),
                                                  _ptr(loss), _ptr(dq), *[_ptr(out.get(k)) for k in self.STATE], _ptr(status),
                                                  _ptr(ws), ws.numel(), C.byref(self.options), self._stream()))
        out.update({"loss": loss, "dq": dq, "status": status, "q": q})
        bad = self._failed_rows(status) if rescue else None
        if bad is not None:
            sub = self._direct_twin().loss_and_grad(prog, q[bad], grad=grad, outputs=outputs)
            for k in tuple(outputs) + ("loss", "status") + (("dq",) if grad else ()):
                out[k][bad] = sub[k]
        return out

    def vjp(self, q, xyz=None, residuals=None, vectors=None, lengths=None, forces=None, rescue=False):
        """q-bar [B,E] for cotangents of the state arrays (any subset)."""
        torch = self.torch
        q = self._q(q)
        B = q.shape[0]
        shapes = self._shapes(B)
        cots = []
        for name, c in zip(self.STATE, (xyz, residuals, vectors, lengths, forces)):
            if c is None:
                cots.append(None)
                continue
            t = c if isinstance(c, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(c, dtype=np.float64)))
            t = t.to(device=self.device, dtype=torch.float64).reshape(shapes[name]).contiguous()
            cots.append(t)
        dq = torch.empty_like(q)
        status = torch.empty(B, dtype=torch.int32, device=self.device)
        ws = self._workspace(B)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib.dfdm_vjp(self.plan.handle, B, _ptr(q), _ptr(self.loads), _ptr(self.xyz_fixed),
                                        *[_ptr(c) for c in cots], _ptr(dq), _ptr(status), _ptr(ws), ws.numel(),
                                        C.byref(self.options), self._stream()))
        bad = self._failed_rows(status) if rescue else None
        if bad is not None:
            sub_dq, sub_status = self._direct_twin().vjp(q[bad], *[None if c is None else c[bad] for c in cots])
            dq[bad] = sub_dq
            status[bad] = sub_status
        return dq, status

    # -- host entry points (copies inside the call) ------------------------------------------------------

    def loss_and_grad_host(self, prog, q, loss, dq, status):
        """NumPy (ideally pinned) buffers in, NumPy buffers out; includes H2D and D2H copies."""
        B = q.shape[0]
        with self.torch.cuda.device(self.device):
            _lib.check(_lib.lib.dfdm_loss_and_grad_host(self.plan.handle, prog.handle, B, _np_ptr(q), _np_ptr(self.loads_host),
                                                       _np_ptr(self.xyz_fixed_host), _np_ptr(loss),
                                                       C.c_void_p(0) if dq is None else _np_ptr(dq), _np_ptr(status),
                                                       C.byref(self.options)))

    def loss_and_grad_host_many(self, prog, batches, threads=2):
        """A stream of batches with host buffers: `batches` is a list of (q, loss, dq, status) NumPy buffers (pinned for the
        overlap), one dfdm_loss_and_grad_host call each.  Batch i is evaluated by host thread i % threads, every thread
        working through its batches in order, so batches i and i + threads may reuse the same buffers and no others may.
        The library gives every call in flight an arena and streams of its own and lets the calls take turns with the
        kernels (include/dfdm_b200.h, host-buffer variants): one call's copies run under another call's kernels.
        Returns when every batch is back on the host."""
        from concurrent.futures import ThreadPoolExecutor
        threads = max(1, min(int(threads), len(batches)))
        if threads == 1:
            for b in batches:
                self.loss_and_grad_host(prog, *b)
            return
        # one long-lived worker thread per slot (the library keeps per-thread pinned words and events: threads are reused)
        workers = getattr(self, "_host_workers", None)
        if workers is None or len(workers) < threads:
            for w in workers or ():
                w.shutdown(wait=True)
            workers = self._host_workers = [ThreadPoolExecutor(max_workers=1, thread_name_prefix="dfdm-host-%d" % t) for t in range(threads)]

        def work(t):
            for b in batches[t::threads]:
                self.loss_and_grad_host(prog, *b)

        futures = [workers[t].submit(work, t) for t in range(threads)]
        errors = []
        for f in futures:
            try:
                f.result()
            except BaseException as exc:                                 # wait for every slot before raising
                errors.append(exc)
        if errors:
            raise errors[0]

    def forward_host(self, q, xyz=None, residuals=None, vectors=None, lengths=None, forces=None, status=None):
        B = q.shape[0]
        args = [C.c_void_p(0) if a is None else _np_ptr(a) for a in (xyz, residuals, vectors, lengths, forces, status)]
        with self.torch.cuda.device(self.device):
            _lib.check(_lib.lib.dfdm_forward_host(self.plan.handle, B, _np_ptr(q), _np_ptr(self.loads_host),
                                                 _np_ptr(self.xyz_fixed_host), *args, C.byref(self.options)))

    # -- helpers -------------------------------------------------------------------------------------------

    @staticmethod
    def raise_on_status(status):
        """Map per-design solver verdicts onto the exception the reference raises (model.py:55)."""
        status = np.asarray(status.cpu() if hasattr(status, "cpu") else status)
        bad = np.nonzero(status != _lib.STATUS_OK)[0]
        if bad.size:
            names = {1: "singular stiffness matrix", 2: "PCG did not converge", 3: "non-finite result"}
            raise LinAlgError("%s (design %d of %d)" % (names.get(int(status[bad[0]]), "solver failure"), int(bad[0]), status.size))
