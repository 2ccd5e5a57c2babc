"""blockcg_b200: B200-native (sm_100a) block-CG hot path of lkeegan/blockCG.

Only what the path needs lives here:
  csrc/    hand-written CUDA kernels + the C-ABI  (libblockcg_b200.so)
  host/    C++ headers mirroring the reference's fields.hpp / dirac_op.hpp /
           block_solvers.hpp on top of the C-ABI (the drop-in for C++ callers)
  capi.py  ctypes binding of the C-ABI (used by tests and bench.py)
  solvers.py  the reference's call signatures for Python callers
"""
from .capi import BcgError, Context, SolveInfo, load  # noqa: F401
from .solvers import BCG, BCGrQ, CG, SBCGrQ, SCG, block_fermion_field, dirac_op  # noqa: F401
