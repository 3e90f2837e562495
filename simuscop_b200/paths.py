"""Filesystem locations of the built artefacts (all in-tree)."""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "simuscop_b200")
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
LIB_CUDA = os.environ.get("SIMUSCOP_CUDA_LIB") or os.path.join(PKG, "libsimuscop_cuda.so")
LIB_HOST = os.path.join(PKG, "libsimuscop_host.so")
SIMUREADS = os.path.join(PKG, "simuReads")
DATA = os.path.join(ROOT, "data")
ORACLE = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE, "libssc_oracle.so")
REF_PHILOX = os.path.join(ORACLE, "_ref", "simuReads_philox")
REF_PLAIN = os.path.join(ORACLE, "_ref", "simuReads_ref")
