"""Statistics the reference's example scripts print at the end (e.g. examples/pillow.py:278-295)."""
import numpy as np


def report(name, network):
    q = list(network.edges_forcedensities())
    f = list(network.edges_forces())
    l = list(network.edges_lengths())
    print(f"\nDesign {name}")
    print(f"Load path: {round(network.loadpath(), 3)}")
    for label, values in (("FDs", q), ("Forces", f), ("Lengths", l)):
        print(f"{label}\t\tMin: {round(min(values), 3)}\tMax: {round(max(values), 3)}\tMean: {round(float(np.mean(values)), 3)}")
    return {"loadpath": network.loadpath(), "q": q, "forces": f, "lengths": l}
