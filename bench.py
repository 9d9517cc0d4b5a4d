#!/usr/bin/env python
"""
bench.py — equilibrium+VJP evaluations per second on synthetic networks of the named shapes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload grid256|pillow|pringle|monkey_saddle|dome|arch]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference algorithm on the host cores (oracle port)

One "step" = one pass of the hot path over one batch: for every design of the batch, equilibrium
state + scalar loss + full dL/dq.  Default workload: BASELINE.json configs[4] (256x256 free-node quad
cable net, batch 256 per GPU, PCG path) — the largest single-GPU configuration and the one whose
kernels the HBM roofline applies to.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "equilibrium+VJP evals/sec"
UNIT = "evals/s"
DEFAULT_BATCH = {"grid256": 256, "pillow": 1024, "pringle": 1024, "monkey_saddle": 4096, "dome": 512, "arch": 64}
L2_BYTES = 126 * 1024 * 1024


# --------------------------------------------------------------------------------------
# workloads
# --------------------------------------------------------------------------------------

def make_workload(name, batch):
    from dfdm_b200 import workloads as W
    import cases
    if name.startswith("grid"):
        nfree = int(name[4:])
        edges, fixed, xyz, loads, q0 = W.grid_arrays(nfree)
        q = W.sample_q(q0, batch, seed=6, spread=0.2, relative=True)
        lengths0 = np.linalg.norm(xyz[edges[:, 1]] - xyz[edges[:, 0]], axis=1)
        goals = [("EdgeLength", e, 0, 1.0, [lengths0[e]]) for e in range(len(edges))]
        return {"name": name, "edges": edges, "is_fixed": fixed, "xyz": xyz, "loads": loads, "q": q,
                "program": (goals, [("SquaredError", 1.0)], 0.1), "sparse": {"targets": lengths0, "weight": 1.0, "l2": 0.1},
                "loss": "SquaredError(EdgeLengthGoal for every edge, target = initial length) + L2Regularizer(0.1)"}
    builders = {"pillow": (cases.case_pillow, 2, 0.1, False, None, None),
                "pringle": (cases.case_pringle, 3, 0.1, False, -5.0, -0.1),
                "monkey_saddle": (cases.case_monkey_saddle, 4, 1.0, True, -20.0, -0.01),
                "dome": (cases.case_dome, 5, 0.2, True, None, None),
                "arch": (lambda: _arch_case(), 1, 0.1, False, None, None)}
    build, seed, spread, relative, lo, hi = builders[name]
    case = build()
    edges, fixed, xyz, loads = cases.case_arrays(case)
    q0 = np.asarray(case["network"].edges_forcedensities(), dtype=np.float64)
    q = W.sample_q(q0, batch, seed=seed, spread=spread, relative=relative, lo=lo, hi=hi)
    return {"name": name, "edges": edges, "is_fixed": fixed, "xyz": xyz, "loads": loads, "q": q,
            "program": cases.program_inputs(case), "terms": cases.oracle_terms(case), "sparse": None,
            "loss": "+".join(t[0] for t in case["terms"])}


def _arch_case():
    from dfdm_b200 import workloads as W
    net = W.arch(5000)
    goals = [("NodeResidualDirection", 0, [-1.0, 0.0, -0.4], 1.0), ("NodeResidualDirection", 5000, [1.0, 0.0, -0.4], 1.0)]
    return {"name": "arch", "network": net, "q": None, "terms": [("SquaredError", goals, 1.0)]}


# --------------------------------------------------------------------------------------
# CPU arm: the reference algorithm restated (oracle), all host cores, bounded sample
# --------------------------------------------------------------------------------------

_CPU = {}


def _cpu_init(wl):
    from oracle import fdm_oracle as O
    import torch
    torch.set_num_threads(1)
    fixed_rows = np.nonzero(wl["is_fixed"])[0]
    if wl["sparse"]:
        _CPU["model"] = O.SparseModel(len(wl["is_fixed"]), wl["edges"], wl["is_fixed"], wl["loads"], wl["xyz"][fixed_rows])
    else:
        s = O.Structure(len(wl["is_fixed"]), wl["edges"], wl["is_fixed"])
        _CPU["model"] = O.Model(s, wl["loads"], wl["xyz"][fixed_rows])
    _CPU["wl"] = wl


def _cpu_eval(q):
    from oracle import fdm_oracle as O
    wl, model = _CPU["wl"], _CPU["model"]
    if wl["sparse"]:
        sp = wl["sparse"]
        loss, dq, _ = model.loss_and_grad_edge_length(q, sp["targets"], sp["weight"], sp["l2"])
    else:
        loss, dq = O.loss_and_grad(wl["terms"], q, model)
    return float(loss)


def cpu_arm(wl, budget_s=20.0, steps=1):
    """Returns (evals/s, cores, sample description, ms per step).  Each step evaluates `sample` designs on all cores."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    ctx = mp.get_context("spawn")          # workers must not inherit a parent that already ran autograd / CUDA
    with ctx.Pool(1, initializer=_cpu_init, initargs=(wl,)) as probe:
        probe.map(_cpu_eval, [wl["q"][0]])
        t0 = time.perf_counter()
        probe.map(_cpu_eval, [wl["q"][0]])
        one = time.perf_counter() - t0
    sample = int(max(cores, min(len(wl["q"]), cores * max(1.0, budget_s / max(one, 1e-6)) / max(steps, 1))))
    sample = max(cores, (sample // cores) * cores)
    sample = min(sample, max(cores, (len(wl["q"]) // cores) * cores)) if len(wl["q"]) >= cores else len(wl["q"])
    qs = [wl["q"][k % len(wl["q"])] for k in range(sample)]
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(wl,)) as pool:
        pool.map(_cpu_eval, qs[:cores])                      # warm-up: imports, first factorisation
        times = []
        for _ in range(steps):
            t0 = time.perf_counter()
            pool.map(_cpu_eval, qs, chunksize=max(1, sample // cores))
            times.append(time.perf_counter() - t0)
    total = sum(times)
    kind = "sparse CPU restatement (scipy CSR + SuperLU + hand adjoint; dense reference path infeasible at this size)" \
        if wl["sparse"] else "oracle port of the reference (dense NumPy forward + torch-f64 tape gradient)"
    desc = "%d designs of workload %s per step, %d processes x 1 thread, %s" % (sample, wl["name"], cores, kind)
    return sample * steps / total, cores, desc, 1e3 * total / steps


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.tmp.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(self.NAMES, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.tmp.name)
        if not sm:                                   # the loop never got a sample in: one direct query
            try:
                line = subprocess.run(["nvidia-smi", "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits"],
                                      capture_output=True, text=True, timeout=10).stdout.splitlines()[0]
                parts = [p.strip() for p in line.split(",")]
                sm, mx = [float(parts[0])], [float(parts[1])]
                reasons = {n for n, v in zip(self.NAMES, parts[2:6]) if v.lower().startswith("active")}
            except Exception:
                pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------

def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fp:
            return float(json.load(fp)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_for(kernel, workload):
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as fp:
            return json.load(fp).get(workload, {}).get(kernel)
    return None


def gpu_arm(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from dfdm_b200 import _lib
    from dfdm_b200.engine import Plan, GoalProgram, Evaluator
    from dfdm_b200.distributed import ShardedBatch, gather_designs, bind_to_gpu_cpus

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    cpus = None
    if world > 1:
        if os.environ.get("DFDM_BIND_NUMA", "1") != "0":             # host buffers of a rank on the NUMA node of its GPU
            cpus = bind_to_gpu_cpus(local_rank)
        dist.init_process_group("nccl", device_id=device)
    batch = args.batch or DEFAULT_BATCH.get(args.workload, 256)
    if args.scaling == "strong":                                   # fixed global batch, cut across the ranks
        if batch % world:
            raise SystemExit("--scaling strong needs a batch divisible by the number of GPUs")
        batch //= world
    wl = make_workload(args.workload, batch * world)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, cores, desc, _ = cpu_arm(wl, budget_s=args.cpu_budget)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}

    q_all = wl["q"]
    lo, hi = rank * batch, (rank + 1) * batch                      # weak scaling: fixed work per GPU
    q_host = torch.from_numpy(np.ascontiguousarray(q_all[lo:hi])).pin_memory()
    plan = Plan(len(wl["is_fixed"]), wl["edges"], wl["is_fixed"])
    fixed_rows = np.nonzero(wl["is_fixed"])[0]
    ev = Evaluator(plan, wl["loads"], wl["xyz"][fixed_rows], device=device, solver=args.solver,
                   pcg_rtol=args.pcg_rtol, pcg_rtol_adjoint=args.pcg_rtol_adjoint, pcg_check_every=args.pcg_check_every, pcg_maxiter=args.pcg_maxiter, pcg_precond=args.precond, mg_omega=args.mg_omega, mg_over=args.mg_over, mg_w_levels=args.mg_w_levels, mg_over_w=args.mg_over_w, flags=(_lib.FLAG_SPMV_MATRIX_FREE if args.spmv_matrix_free else 0) | (_lib.FLAG_MG_NO_TAIL if args.no_tail else 0))
    prog = GoalProgram(plan, *wl["program"])
    q_dev = q_host.to(device)
    E = plan.n_edges
    solver = plan.auto_solver if args.solver == "auto" else args.solver
    if args.steps <= 0:
        args.steps = 3 if solver == "pcg" else 50

    # the one exchange of a batched evaluation, named by what the consumer needs (--exchange):
    #   allgather  losses and gradients of every design on every rank - peer stores over NVLink, every piece of the local
    #              slice pushed under the kernels of the next piece (--chunks), NCCL when the ranks cannot map each other
    #              (DFDM_PEER_GATHER=0 forces NCCL)
    #   root       rank 0 only receives them (a host-side optimiser on one rank)
    #   losses     gradients stay with their designs, all losses everywhere (a sharded optimiser: B doubles on the wire) - default
    #   mean       ensemble mean of loss and gradient: one all-reduce of E + 1 doubles
    def evaluate(q_local):
        out = ev.loss_and_grad(prog, q_local)
        last["out"] = out
        return out["loss"], out["dq"]

    last = {}
    sharded = ShardedBatch(evaluate, peer=os.environ.get("DFDM_PEER_GATHER", "1") != "0", exchange=args.exchange,
                           zero_copy=True, chunks=args.chunks if world > 1 else 1)

    def step():
        loss_x, dq_x = sharded.loss_and_grad_local(q_dev, batch * world)
        return last["out"], loss_x, dq_x

    working_set = _lib.lib.dfdm_workspace_bytes(plan.handle, batch, None) if solver == "pcg" else batch * E * 8 * 12
    flush = working_set < 2 * L2_BYTES
    flush_buf = torch.empty(2 * L2_BYTES // 8, dtype=torch.float64, device=device) if flush else None

    sampler = ClockSampler(local_rank) if rank == 0 and os.environ.get("DFDM_BENCH_SAMPLER", "1") != "0" else None      # sampled from the warm-up on: short steps last < 200 ms
    for _ in range(args.warmup):
        out, loss_x, dq_x = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    out = ev.loss_and_grad(prog, q_dev)                            # the whole local slice in one call: status, reference for e2e
    status = out["status"].cpu().numpy()
    gather_ok = None
    if world > 1:
        gather_ok = exchange_check(args.exchange, world, rank, batch, sharded.last_local, loss_x, dq_x, device)
    iters = _lib.last_pcg_iterations() if solver == "pcg" else (0, 0)
    direct_kernel = _lib.last_direct_kernel() if solver == "direct" else None

    # ---- timed regions: device-resident inputs -------------------------------------------------------
    # (1) K steps with nothing but the step's own launches on the stream: `value`.
    # (2) K more steps with a CUDA event after every launch of the PCG path: the per-kernel table and the roofline.
    #     Events between kernels cost ~3 % (they also defeat the programmatic dependent launches), so they stay out of (1).
    def timed_region(profile):
        _lib.profile_enable(profile)
        launches0 = _lib.launch_count()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        for ev_ in starts + ends:                       # events are created lazily: do it outside the timed steps
            ev_.record()
        step()                                          # untimed pre-roll: the device is busy (clocks up) when the first timed step starts
        step()
        host_ms = []
        for k in range(args.steps):
            if flush:
                flush_buf.fill_(float(k))
            starts[k].record()
            h0 = time.perf_counter()
            step()
            host_ms.append((time.perf_counter() - h0) * 1e3)
            ends[k].record()
        torch.cuda.synchronize()
        last["ms_steps_host"] = host_ms
        if world > 1:
            dist.barrier()
        count = _lib.launch_count() - launches0
        prof = _lib.profile_read()
        _lib.profile_enable(False)
        per_step = [s.elapsed_time(e) for s, e in zip(starts, ends)]
        total = sum(per_step)
        last["ms_steps"] = per_step
        t = torch.tensor([total, float(count)], dtype=torch.float64, device=device)
        if world > 1:
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            total, count = float(tmax[0]), int(t[1])
        return total, count, prof

    ms_total, launches, _ = timed_region(False)
    ms_steps = [round(v, 3) for v in last["ms_steps"]]
    ms_steps_host = [round(v, 3) for v in last["ms_steps_host"]]
    value = batch * world * args.steps / (ms_total * 1e-3)
    prof_ms, prof_count, ms_profiled = [0.0] * 8, [0] * 8, 0.0
    if solver == "pcg" and not args.no_profile:
        ms_profiled, _, (prof_ms, prof_count) = timed_region(True)
    clocks = sampler.stop() if sampler else None

    # ---- end to end: host buffers through the C ABI, copies inside the timed region --------------------
    # Every step copies its q in from pinned host memory and its losses, gradients and verdicts back out.  Measured twice:
    # one call at a time (`serial`: copy in, kernels, copy out, strictly in sequence) and as a stream of batches issued by
    # two host threads, whose calls take turns with the kernels so that one step's copies run under the other's kernels
    # (Evaluator.loss_and_grad_host_many; each thread has output buffers of its own).
    def pinned(*shape, dtype=torch.float64):
        return torch.empty(*shape, dtype=dtype).pin_memory().numpy()
    n_threads = 2
    outs = [(pinned(batch), pinned(batch, E), pinned(batch, dtype=torch.int32)) for _ in range(n_threads)]
    loss_h, dq_h, status_h = outs[0]
    q_np = q_host.numpy()

    def e2e_region(threads):
        batches = [(q_np,) + outs[k % threads] for k in range(args.steps)]
        ev.loss_and_grad_host_many(prog, batches[:threads], threads=threads)     # warm the cached arenas
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        ev.loss_and_grad_host_many(prog, batches, threads=threads)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return float(te[0])

    e2e_serial_s = e2e_region(1)
    e2e_stream_s = e2e_region(n_threads) if args.steps >= n_threads else e2e_serial_s
    e2e_s = min(e2e_serial_s, e2e_stream_s)
    e2e_api = ("dfdm_loss_and_grad_host on pinned host buffers, one call per step, issued by %d host threads: the calls take turns with the "
               "kernels and one step's copies run under the other's kernels" % n_threads) if e2e_stream_s < e2e_serial_s else \
              "dfdm_loss_and_grad_host on pinned host buffers, one call at a time"
    h2d = q_np.nbytes + ev.loads_host.nbytes + ev.xyz_fixed_host.nbytes
    d2h = loss_h.nbytes + dq_h.nbytes + status_h.nbytes
    ref_loss = out["loss"].cpu().numpy()
    same = all(bool(np.allclose(o[0], ref_loss, rtol=1e-12, atol=0)) and bool(np.all(o[2] == status)) for o in outs)

    if rank != 0:
        return None
    peak, peak_src = measured_peak()
    nf, nnz = plan.n_free, plan.nnz
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "batch_per_gpu": batch, "global_batch": batch * world, "n_nodes": plan.n,
                       "n_edges": E, "n_free": nf, "nnz": nnz, "solver": solver, "loss": wl["loss"],
                       "pcg_iterations": {"forward": iters[0], "adjoint": iters[1]},
                       "pcg_rtol": args.pcg_rtol or _lib.PCG_RTOL_DEFAULT, "pcg_rtol_adjoint": args.pcg_rtol_adjoint or (_lib.PCG_RTOL_ADJOINT_DEFAULT if (plan.info.get("mg_levels", 0) > 0 and args.precond != "jacobi") else _lib.PCG_RTOL_ADJOINT_JACOBI),
                       "pcg_rtol_note": "stop tests calibrated against SuperLU / the dense oracle with a tenfold margin to the 1e-9 / 1e-7 contract (include/dfdm_b200.h, tools/rtol_sweep.py)",
                       "preconditioner": ("aggregation multigrid, %d levels, %s" % (plan.info.get("mg_levels", 0) + 1, "V-cycle" if args.mg_w_levels < 0 else "W-cycle on the first %d coarse levels" % (args.mg_w_levels or 3))) if (solver == "pcg" and args.precond != "jacobi" and plan.info.get("mg_levels", 0) > 0) else ("jacobi" if solver == "pcg" else None),
                       "parallelism": "dp%d (designs sharded; exchange per step: %s)" % (world, sharded.collective if world > 1 else "none"),
                       "exchange": args.exchange if world > 1 else None, "gather_ok": gather_ok,
                       "cache": "L2 flushed between steps" if flush else "per-step working set %.1f GB >> 126 MB L2" % (working_set / 1e9),
                       "status_ok": bool(np.all(status == 0)), "e2e_matches_device": same,
                       "cpu_binding": ("%d cores next to the GPU" % len(cpus)) if cpus else "none"},
            "clocks": clocks,
            "e2e": {"value": batch * world * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d) * world,
                    "d2h_bytes_per_step": int(d2h) * world,
                    "api": e2e_api, "serial": batch * world * args.steps / e2e_serial_s,
                    "streamed_%d_threads" % n_threads: batch * world * args.steps / e2e_stream_s},
            "gpu_launches": launches, "ms_steps": ms_steps, "ms_steps_host_enqueue": ms_steps_host}
    if solver == "pcg" and sum(prof_count) == 0:
        line["roofline"] = {"bound": "hbm", "kernel": "whole step (per-kernel table skipped: --no-profile)", "achieved": None, "peak": peak, "unit": "GB/s",
                            "frac": None, "traffic": None, "peak_source": peak_src}
    elif solver == "pcg":
        nc = plan.info.get("mg_n_coarse", 0)
        is_mg = plan.info.get("mg_levels", 0) > 0 and args.precond != "jacobi"
        # spmv: K (8 nnz), p in single precision (12 nf), Ap out (24 nf); update: r in/out, Ap (+ the single-precision copy of r the
        # cycle reads, or 1/diag for Jacobi); direction: p in, p out (12 nf each), z (float) or r + 1/diag, plus its share of
        # k_pcg_xupdate, which adds a window of 8 iterations to x in one pass: (48 + 8 x 12) nf / 8 = 18 nf per iteration
        alg = {"k_pcg_spmv": ((8.0 * E if args.spmv_matrix_free else 8.0 * nnz) + 36.0 * nf) * batch, "k_pcg_update": (84.0 if is_mg else 80.0) * nf * batch,
               "k_pcg_direction": 0.0,      # set below: depends on how many x windows fall inside the iterations
               # single-precision V-cycle: K D^-1 (4 nnz), r (12 nf), t out (12 nf), coarse rhs out (12 nc)
               "k_mg_down[level 0]": (4.0 * nnz + 24.0 * nf + 12.0 * nc) * batch,
               # K (4 nnz), r (12 nf), t (12 nf), 1/diag (4 nf), coarse solution (12 nc), z out (12 nf)
               "k_mg_up[level 0]": (4.0 * nnz + 40.0 * nf + 12.0 * nc) * batch, "multigrid coarse levels (per V-cycle)": 0.0}
        # direction: p in, z (or r + 1/diag) in, p out; the x update runs after every 8th iteration of a solve (48 + 8 x 12 n_f bytes) and
        # is timed with the direction pass before it; what is left of a window at the end of a solve is a pass of its own
        XW = 8
        n_it = max(1, iters[0] + iters[1])
        full_windows = iters[0] // XW + iters[1] // XW
        alg["k_pcg_direction"] = ((36.0 if is_mg else 56.0) + (48.0 + 12.0 * XW) * full_windows / n_it) * nf * batch
        x_tail_bytes = sum((48.0 + 12.0 * (k % XW)) * nf * batch for k in iters if k % XW)
        if is_mg:
            alg["multigrid coarse levels (per V-cycle)"] = coarse_cycle_bytes(plan, args.mg_w_levels) * batch
        alg["k_assemble"] = (8.0 * E + 8.0 * nnz + 32.0 * nf) * batch      # q in; K values, 1/diag, b (3 columns) out
        # k_positions 48 n; k_edge_state 48 E + 24 n; k_node_state 32 E + 96 n; k_edge_adjoint 80 E + 24 n; k_adjoint_rhs 24 E + 48 n
        alg["state, loss and reverse edge kernels (5 launches)"] = (184.0 * E + 240.0 * plan.n) * batch
        kernels = []
        for k, name in enumerate(("k_pcg_spmv", "k_pcg_update", "k_pcg_direction", "k_mg_down[level 0]", "k_mg_up[level 0]",
                                  "multigrid coarse levels (per V-cycle)", "k_assemble", "state, loss and reverse edge kernels (5 launches)")):
            if prof_count[k]:
                avg_ms = prof_ms[k] / prof_count[k]
                kernels.append({"kernel": name, "launches": prof_count[k], "avg_ms": avg_ms, "share_of_step": prof_ms[k] / ms_profiled,
                                "algorithmic_bytes": alg[name], "achieved_gbs": alg[name] / (avg_ms * 1e-3) / 1e9})
        once = ("k_assemble", "state, loss and reverse edge kernels (5 launches)")      # these run once per evaluation
        per_iteration = [kk for kk in kernels if kk["kernel"] not in once]
        it_ms = sum(kk["avg_ms"] for kk in per_iteration)
        it_bytes = sum(kk["algorithmic_bytes"] for kk in per_iteration)
        # the WHOLE step against the roofline: algorithmic bytes of everything a step launches / ms_per_step of the
        # un-instrumented region.  Per solve: k_pcg_init (r, 1/diag in; x, p, single-precision r out: 80 n_f), one preconditioner
        # cycle and one direction pass without the x update (36 n_f) before the first iteration, then the iterations.
        n_iter = iters[0] + iters[1]
        cyc = alg["k_mg_down[level 0]"] + alg["k_mg_up[level 0]"] + alg["multigrid coarse levels (per V-cycle)"] if is_mg else 0.0
        step_parts = {"pcg iterations (%d forward + %d adjoint)" % iters: n_iter * it_bytes,
                      "solver start (init, first cycle, first direction) x 2": 2 * (80.0 * nf * batch + cyc + 36.0 * nf * batch),
                      "k_assemble": alg["k_assemble"], "state, loss and reverse edge kernels": alg["state, loss and reverse edge kernels (5 launches)"],
                      "hierarchy set-up (single-precision copies, Galerkin sums)": mg_setup_bytes(plan) * batch if is_mg else 0.0,
                      "k_edge_dq": (40.0 * E + 24.0 * plan.n) * batch, "layout changes of q and dq": 32.0 * E * batch,
                      "x updates of the last, partial windows": x_tail_bytes}
        step_bytes = sum(step_parts.values())
        ms_step = ms_total / args.steps
        top = max((kk for kk in kernels if kk["algorithmic_bytes"] > 0), key=lambda kk: kk["share_of_step"])
        line["roofline"] = {"bound": "hbm", "kernel": "whole step; dominant group: " + top["kernel"], "achieved": step_bytes / (ms_step * 1e-3) / 1e9,
                            "peak": peak, "unit": "GB/s", "frac": step_bytes / (ms_step * 1e-3) / 1e9 / peak,
                            "traffic": traffic_for(top["kernel"], wl["name"]), "peak_source": peak_src,
                            "traffic_of": "the dominant group, DRAM bytes read + written per launch (group: per cycle) from ncu --set full (profiles/traffic.json); its algorithmic bytes are under 'kernels'",
                            "what": "algorithmic bytes of one whole step (step_bytes) / ms_per_step / peak; per-kernel and per-group numbers under 'kernels'",
                            "step_bytes": step_bytes, "step_bytes_parts": step_parts,
                            "dominant": {"kernel": top["kernel"], "share_of_step": top["share_of_step"], "achieved": top["achieved_gbs"],
                                         "frac": top["achieved_gbs"] / peak},
                            "timing": "value / ms_per_step: CUDA events around each step, nothing else on the stream; 'kernels': CUDA events after every PCG launch in a second timed region of the same K steps",
                            "ms_per_step_with_events": ms_profiled / args.steps}
        line["kernels"] = kernels
        line["roofline"]["iteration"] = {"what": "one whole PCG iteration = the kernels listed under 'kernels' except k_assemble and the state / loss group", "ms": it_ms,
                                         "algorithmic_bytes": it_bytes, "achieved": it_bytes / (it_ms * 1e-3) / 1e9,
                                         "frac": it_bytes / (it_ms * 1e-3) / 1e9 / peak, "share_of_step": sum(kk["share_of_step"] for kk in per_iteration)}
        line["roofline"]["frac_of_nominal_8000"] = line["roofline"]["achieved"] / 8000.0
    else:
        # direct path: one fused kernel per step; mandatory I/O only (q in, loss + dq out)
        alg = batch * (16.0 * E + 8.0)
        ms = ms_total / args.steps
        line["roofline"] = {"bound": "hbm", "kernel": direct_kernel, "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                            "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": traffic_for(direct_kernel, wl["name"]),
                            "peak_source": peak_src,
                            "note": "shared-memory factorisation: latency / FP64-issue bound, not HBM bound; fraction reported for completeness"}
        # FP64 view of the same kernel: banded LDL^T nf*w*(w+2) flops, two solves x 3 right-hand sides x (4 nf w + nf)
        w_band = plan.info["bandwidth_solver"]
        flops = batch * (nf * w_band * (w_band + 2.0) + 6.0 * (4.0 * nf * w_band + nf))
        dgemm = fp64_peak_gflops(device)
        line["roofline"]["fp64"] = {"flops_per_step": flops, "achieved_gflops": flops / (ms * 1e-3) / 1e9, "peak_gflops": dgemm,
                                    "frac": flops / (ms * 1e-3) / 1e9 / dgemm, "peak_source": "torch.matmul f64 4096^3 on this GPU, best of 5 (calibration only)"}
        line["roofline"]["frac_of_nominal_8000"] = alg / (ms * 1e-3) / 1e9 / 8000.0
    if cpu:
        line["cpu_baseline"] = cpu
    return line


def fp64_peak_gflops(device):
    """Measured FP64 GEMM rate of this GPU: the ceiling the banded factorisation is compared with (not part of the path)."""
    import torch
    a = torch.randn(4096, 4096, dtype=torch.float64, device=device)
    b = torch.randn(4096, 4096, dtype=torch.float64, device=device)
    torch.matmul(a, b)
    best = float("inf")
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        torch.matmul(a, b)
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return 2.0 * 4096 ** 3 / (best * 1e-3) / 1e9


def exchange_check(mode, world, rank, batch, pieces, loss_x, dq_x, device):
    """Every row of the exchanged arrays against its owner's copy, bit for bit: the owners' per-row integer checksums
    travel through NCCL (an independent path) and are compared with checksums of the rows as received."""
    import torch
    import torch.distributed as dist
    out = {"loss": torch.cat([p[0] for p in pieces]), "dq": torch.cat([p[1] for p in pieces])}     # this rank's own rows of that step

    def rowsum(t):                                                  # exact, order-independent: wrap-around integer sums
        return t.contiguous().view(torch.int64).reshape(t.shape[0], -1).sum(dim=1)

    own = torch.stack([rowsum(out["loss"].reshape(-1, 1)), rowsum(out["dq"])], dim=1)
    every = torch.empty((world * batch, 2), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(every, own)
    ok = True
    if mode == "allgather" or (mode == "root" and rank == 0):
        ok = bool(torch.equal(rowsum(loss_x.reshape(-1, 1)), every[:, 0])) and bool(torch.equal(rowsum(dq_x), every[:, 1]))
    elif mode == "losses":
        ok = bool(torch.equal(rowsum(loss_x.reshape(-1, 1)), every[:, 0])) and bool(torch.equal(rowsum(dq_x), own[:, 1]))
    elif mode == "mean":
        packed = torch.cat([out["dq"].sum(dim=0), out["loss"].sum().reshape(1)])
        dist.all_reduce(packed)
        packed /= float(world * batch)
        ok = bool(torch.allclose(dq_x, packed[:-1], rtol=1e-12, atol=0)) and bool(torch.allclose(loss_x, packed[-1], rtol=1e-12, atol=0))
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) != 1:
        raise SystemExit("bench.py: the exchanged rows differ from their owners' rows (exchange %s)" % mode)
    return True


def mg_setup_bytes(plan):
    """Algorithmic bytes per design of the numerical set-up of the hierarchy: level 0 reads K (8 nnz) and 1/diag (8 n) and
    writes two single-precision copies and 1/diag (8 nnz + 4 n); every coarse level sums its parent's values (4 nnz_parent
    in, 4 nnz out), takes 1/diag (4 n out) and the column-scaled copy (4 nnz in, 4 nnz out)."""
    from dfdm_b200 import _lib
    rows = [int(v) for v in plan.array(_lib.ARR_MG_ROWS)]
    nnzs = [int(v) for v in plan.array(_lib.ARR_MG_NNZ)]
    if len(rows) < 2:
        return 0.0
    total = 16.0 * nnzs[0] + 12.0 * rows[0]
    for l in range(1, len(rows)):
        total += 4.0 * nnzs[l - 1] + 12.0 * nnzs[l] + 4.0 * rows[l]
    return total


def coarse_cycle_bytes(plan, mg_w_levels):
    """Algorithmic bytes per design of the levels below level 0 in one preconditioner cycle (single precision):
    per visit of level l: down 4 nnz + 24 n + 12 n_next, up 4 nnz + 40 n + 12 n_next (+ 12 n when it adds to a previous
    correction), the W-cycle's residual 4 nnz + 36 n; the coarsest level is a dense product, 4 n^2 + 24 n."""
    from dfdm_b200 import _lib
    rows = [int(v) for v in plan.array(_lib.ARR_MG_ROWS)]
    nnzs = [int(v) for v in plan.array(_lib.ARR_MG_NNZ)]
    if len(rows) < 2:
        return 0.0
    w = 0 if mg_w_levels < 0 else (mg_w_levels or 3)
    last = len(rows) - 1
    total, visits = 0.0, 1
    for l in range(1, last + 1):
        twice = l <= w
        visits *= 2 if twice else 1
        if l == last:
            total += visits * (4.0 * rows[l] ** 2 + 24.0 * rows[l])
        else:
            down = 4.0 * nnzs[l] + 24.0 * rows[l] + 12.0 * rows[l + 1]
            up = 4.0 * nnzs[l] + 40.0 * rows[l] + 12.0 * rows[l + 1]
            total += visits * (down + up)
            if twice:                                   # second visit: residual of the first, correction added to it
                total += (visits // 2) * (4.0 * nnzs[l] + 36.0 * rows[l] + 12.0 * rows[l])
    return total


def reference_arm(args, rank, world):
    if rank != 0:
        return None
    batch = args.batch or DEFAULT_BATCH.get(args.workload, 256)
    wl = make_workload(args.workload, batch)
    if args.steps <= 0:
        args.steps = 3
    v, cores, desc, ms = cpu_arm(wl, budget_s=args.cpu_budget, steps=args.steps)
    return {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "batch_per_gpu": batch, "global_batch": batch * world, "loss": wl["loss"],
                       "note": "CPU arm: rank 0 only, host cores; the reference has no GPU or multi-process path"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed steps (default: 3; 50 for the sub-millisecond steps of the direct path)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="grid256")
    ap.add_argument("--batch", type=int, default=0, help="designs per GPU (default: the configuration's own)")
    ap.add_argument("--solver", default="auto", choices=["auto", "direct", "pcg"])
    ap.add_argument("--pcg-rtol", type=float, default=0.0)
    ap.add_argument("--pcg-rtol-adjoint", type=float, default=0.0)
    ap.add_argument("--pcg-check-every", type=int, default=0)
    ap.add_argument("--exchange", default="losses", choices=["allgather", "root", "losses", "mean"],
                    help="N > 1: what travels between the ranks after an evaluation (see gpu_arm).  Default: what the repo's own "
                    "consumer of a sharded batch - the device-resident lock-step optimisers - needs: gradients stay with their "
                    "designs, every rank learns every loss.  SURVEY.md 8(e)'s all-gather / gather-to-root of losses AND gradients: "
                    "--exchange allgather / root")
    ap.add_argument("--chunks", type=int, default=1, help="N > 1, allgather by peer stores: pieces of the local slice, each pushed under the next one's kernels "
                    "(measured: evaluating the slice in two pieces costs more than the overlap returns on the PCG path - 70.2 against 60.1 ms per step at N = 2)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="weak: --batch designs per GPU; strong: --batch designs in all")
    ap.add_argument("--pcg-maxiter", type=int, default=0, help="cap PCG iterations (kernel tuning runs only: results not converged)")
    ap.add_argument("--precond", default="auto", choices=["auto", "jacobi", "multigrid"])
    ap.add_argument("--mg-omega", type=float, default=0.0)
    ap.add_argument("--mg-over", type=float, default=0.0)
    ap.add_argument("--mg-over-w", type=float, default=0.0)
    ap.add_argument("--spmv-matrix-free", action="store_true", help="K p from q over the incidence lists instead of the stored CSR values")
    ap.add_argument("--no-tail", action="store_true", help="coarse multigrid levels on the launch-per-level path instead of the tail kernel")
    ap.add_argument("--mg-w-levels", type=int, default=0, help="coarse levels visited twice per cycle (0: default 3, -1: V-cycle)")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for the baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="skip the second, event-instrumented timed region (no per-kernel table)")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: library chatter (NCCL's version banner, ...) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        line = reference_arm(args, rank, world)
    else:
        if world != args.gpus and rank == 0:
            print("note: --gpus %d but WORLD_SIZE=%d; using WORLD_SIZE" % (args.gpus, world), file=sys.stderr)
        line = gpu_arm(args, rank, world, local_rank)
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
