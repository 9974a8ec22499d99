"""voxcraft-sim_b200 — B200-native drop-in for voxcraft-sim's VX3_VoxelyzeKernel step loop.

The product is the C-ABI shared library built from ``csrc/`` (``include/vx3_abi.h``); this Python
package is plumbing around it: the in-tree build, a ctypes binding and synthetic workloads for the
tests and the benchmark.  The directory name carries a hyphen (it is named after the reference repo),
so import it through ``__graft_entry__.load_package()`` which registers it as ``voxcraft_sim_b200``.
"""
__version__ = "0.1.0"
