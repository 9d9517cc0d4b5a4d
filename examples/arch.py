"""
Form-finding of a 5000-segment arch (reference: examples/arch.py, without the viewer).
Uniform chain, q = -1, p_z = -0.1: z_i = p_z i (N - i) / (2 q), so z_mid = 312 500.
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from dfdm_b200 import workloads
from dfdm_b200.equilibrium import fdm


def main(num_segments=5000, verbose=True):
    network = workloads.arch(num_segments, length=5.0, q_init=-1.0, pz=-0.1)
    eq_network = fdm(network)
    z_mid = eq_network.node_attribute(num_segments // 2, "z")
    reaction = eq_network.node_residual(0)
    if verbose:
        print(f"nodes {eq_network.number_of_nodes()}, edges {eq_network.number_of_edges()}")
        print(f"z at mid span: {z_mid}   (closed form {-0.1 * (num_segments / 2) ** 2 / (2 * -1.0)})")
        print(f"reaction at the first support: {reaction}")
    return z_mid, reaction


if __name__ == "__main__":
    main()
