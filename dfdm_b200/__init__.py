"""
dfdm_b200 — B200-native force-density equilibrium + reverse-mode gradient behind dfdm's Python API.

Host code mirrors /root/reference/src/dfdm (same entry points, goal / loss / optimizer objects); the
arithmetic runs in hand-written sm_100a CUDA reached through the C-ABI of ``include/dfdm_b200.h``.
There is no CPU fallback: evaluating a model without the CUDA library and a GPU raises.
"""
__version__ = "0.1.0"
