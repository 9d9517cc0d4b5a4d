"""Experiment: does driving two half batches on two streams (two host threads) hide the latency-bound
multigrid levels behind the other half's bandwidth-bound kernels?  Prints evals/s for 1, 2 and 4 streams."""
import sys, os, time, threading
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_workload
from dfdm_b200.engine import Plan, GoalProgram, Evaluator

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
wl = make_workload("grid256", B)
dev = torch.device("cuda:0")
plan = Plan(len(wl["is_fixed"]), wl["edges"], wl["is_fixed"])
fixed_rows = np.nonzero(wl["is_fixed"])[0]
prog = GoalProgram(plan, *wl["program"])
q = torch.from_numpy(np.ascontiguousarray(wl["q"])).to(dev)


def run(nstreams, reps=3):
    per = B // nstreams
    evs = [Evaluator(plan, wl["loads"], wl["xyz"][fixed_rows], device=dev) for _ in range(nstreams)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(nstreams)]
    outs = [None] * nstreams

    def work(i, n):
        with torch.cuda.stream(streams[i]):
            for _ in range(n):
                outs[i] = evs[i].loss_and_grad(prog, q[i * per:(i + 1) * per])
            streams[i].synchronize()

    def go(n):
        th = [threading.Thread(target=work, args=(i, n)) for i in range(nstreams)]
        for t in th: t.start()
        for t in th: t.join()

    go(2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    go(reps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    loss = torch.cat([o["loss"] for o in outs]).cpu().numpy()
    return B * reps / dt, loss


ref = None
for ns in (1, 2, 4, 1):
    v, loss = run(ns)
    if ref is None: ref = loss
    print("streams %d: %.1f evals/s  max rel loss diff vs 1 stream %.2e" % (ns, v, np.max(np.abs(loss - ref) / np.abs(ref))), flush=True)
