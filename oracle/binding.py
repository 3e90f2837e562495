"""ctypes binding of the CPU oracle (oracle/libssc_oracle.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg.  It lives outside the product package (simuscop_b200/) on purpose: nothing
on the product path imports, links or executes anything under oracle/.
"""
import ctypes as C

import numpy as np

from simuscop_b200 import abi
from simuscop_b200.paths import ORACLE_LIB


class OraclePlan(C.Structure):
    _fields_ = [("prof", C.POINTER(abi.ProfileTables)), ("genome", C.c_void_p), ("genome_len", C.c_uint64),
                ("bins", C.c_void_p), ("n_bins", C.c_int64), ("segs", C.c_void_p), ("n_segs", C.c_int64),
                ("names", C.c_char_p), ("genome_first", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(ORACLE_LIB)
        _lib.ssco_generate.restype = C.c_int64
        _lib.ssco_generate.argtypes = [C.POINTER(OraclePlan), C.c_uint64, C.c_int64, C.c_int64,
                                       C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                       C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                       C.c_void_p, C.c_int64, C.POINTER(C.c_uint64)]
        _lib.ssco_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.ssco_draw_real.restype = C.c_double
        _lib.ssco_draw_real.argtypes = [C.c_uint32, C.c_double, C.c_double]
        _lib.ssco_draw_int.restype = C.c_long
        _lib.ssco_draw_int.argtypes = [C.c_uint32, C.c_long, C.c_long]
        _lib.ssco_rand_indx.restype = C.c_int
        _lib.ssco_rand_indx.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
    return _lib


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().ssco_philox(c.ctypes.data, k.ctypes.data, o.ctypes.data)
    return o


def generate(plan, seed, pair_lo=0, pair_hi=None, trace=False, genome=None, genome_first=0):
    """Run the oracle over a simuscop_b200.planfile.Plan. Returns (fq1, fq2, info).
    genome / genome_first: a window of the haplotype store (ASCII, store index of its first base) instead of plan.genome,
    for plans whose store is too large to hold as ASCII (bench-scale parity tests)."""
    L = lib()
    if pair_hi is None:
        pair_hi = plan.planned_pairs()
    prof = plan.profile_struct()
    op = OraclePlan()
    op.prof = C.pointer(prof)
    genome = np.ascontiguousarray(plan.genome if genome is None else genome)
    bins = np.ascontiguousarray(plan.bins)
    segs = np.ascontiguousarray(plan.segs)
    op.genome = genome.ctypes.data
    op.genome_len = genome.size
    op.genome_first = int(genome_first)
    op.bins = bins.ctypes.data
    op.n_bins = bins.size
    op.segs = segs.ctypes.data
    op.n_segs = segs.size
    op.names = plan.names
    npairs = max(0, pair_hi - pair_lo)
    rl = plan.hdr["read_length"]
    cap = int(npairs * (2 * rl * 2 + 400) + 4096)
    out1 = np.empty(cap, dtype=np.uint8)
    out2 = np.empty(cap if plan.paired else 1, dtype=np.uint8)
    l1, l2, nb = C.c_size_t(0), C.c_size_t(0), C.c_uint64(0)
    tr = np.zeros((npairs, 4), dtype=np.int64) if trace else None
    n = L.ssco_generate(C.byref(op), seed, pair_lo, pair_hi, out1.ctypes.data, out1.size, C.byref(l1),
                        out2.ctypes.data if plan.paired else None, out2.size if plan.paired else 0, C.byref(l2),
                        tr.ctypes.data if trace else None, npairs if trace else 0, C.byref(nb))
    if n == -3:
        raise RuntimeError("oracle: a fragment lies outside the genome window handed over")
    if n < 0:
        raise RuntimeError("oracle output buffer too small (%d)" % n)
    info = dict(emitted=int(n), bases=int(nb.value), trace=tr[:n] if trace else None)
    return out1[:l1.value].tobytes(), out2[:l2.value].tobytes() if plan.paired else b"", info
