"""simuscop_b200 -- B200-native drop-in for the read-generation hot path of SimuSCoP's simuReads.

The product is the C++/CUDA code under csrc/ and host/ (built by __graft_entry__.build()):
  libsimuscop_cuda.so   C-ABI of include/simuscop.h (kernels + device runtime)
  simuReads             drop-in CLI (host front end -> C-ABI)
This Python package only holds measurement / test plumbing (ctypes bindings, synthetic
genomes, plan-file reader).  It never falls back to a CPU implementation.
"""
from . import paths  # noqa: F401
