"""ctypes binding of libsimuscop_cuda.so (the C ABI of include/simuscop.h).

Plumbing for tests and bench.py only: every call goes to the CUDA library; a missing
library or a missing GPU raises -- there is no fallback path.
"""
import ctypes as C

import numpy as np

from . import abi
from .paths import LIB_CUDA

SYMBOLS = ["ssc_last_error", "ssc_version", "ssc_create", "ssc_destroy", "ssc_set_option", "ssc_set_profile",
           "ssc_genome_reserve", "ssc_genome_append", "ssc_genome_size", "ssc_reference_upload", "ssc_reference_upload_fasta", "ssc_reference_prefetch_fasta", "ssc_reference_adopt_prefetched", "ssc_genome_append_ref", "ssc_genome_poke", "ssc_genome_read", "ssc_gc_census", "ssc_set_plan", "ssc_generate",
           "ssc_generate_device", "ssc_get_stats", "ssc_reset_stats", "ssc_table_lookup_host", "ssc_sub_lookup_host", "ssc_gzip_member_host", "ssc_issue_floor"]

_lib = None


class SscError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_CUDA)   # raises OSError when the extension is not built: no fallback
        L.ssc_last_error.restype = C.c_char_p
        L.ssc_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.ssc_destroy.argtypes = [C.c_void_p]
        L.ssc_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
        L.ssc_set_profile.argtypes = [C.c_void_p, C.POINTER(abi.ProfileTables)]
        L.ssc_genome_reserve.argtypes = [C.c_void_p, C.c_uint64]
        L.ssc_genome_append.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.ssc_genome_size.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.ssc_reference_upload.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
        L.ssc_reference_upload_fasta.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32,
                                                 C.POINTER(C.c_uint64)]
        L.ssc_reference_prefetch_fasta.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
        L.ssc_reference_adopt_prefetched.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.ssc_genome_append_ref.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int32, C.POINTER(C.c_uint64)]
        L.ssc_genome_poke.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_int64]
        L.ssc_genome_read.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
        L.ssc_gc_census.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.ssc_set_plan.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                   C.c_char_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.ssc_generate.argtypes = [C.c_void_p, C.c_int64, C.c_int64, abi.SINK_FN, C.c_void_p]
        L.ssc_generate_device.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_uint64),
                                          C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
        L.ssc_get_stats.argtypes = [C.c_void_p, C.POINTER(abi.Stats)]
        L.ssc_reset_stats.argtypes = [C.c_void_p]
        L.ssc_table_lookup_host.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
        L.ssc_sub_lookup_host.argtypes = [C.c_void_p, C.c_uint32]
        L.ssc_gzip_member_host.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.ssc_gzip_member_host.restype = C.c_int64
        L.ssc_issue_floor.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def _ck(rc):
    if rc != 0:
        raise SscError("ssc error %d: %s" % (rc, lib().ssc_last_error().decode(errors="replace")))


class Generator:
    """One handle on one GPU."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        _ck(lib().ssc_create(device, C.byref(self.h)))
        self.planned = 0
        self.emitted = 0
        self._keep = []

    def close(self):
        if self.h:
            lib().ssc_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key, value):
        _ck(lib().ssc_set_option(self.h, key.encode(), int(value)))

    def load_plan(self, plan, seed):
        """Upload profile tables, haplotype store and bins of a planfile.Plan."""
        prof = plan.profile_struct()
        _ck(lib().ssc_set_profile(self.h, C.byref(prof)))
        g = np.ascontiguousarray(plan.genome)
        _ck(lib().ssc_genome_reserve(self.h, g.size))
        first = C.c_uint64()
        _ck(lib().ssc_genome_append(self.h, g.ctypes.data, g.size, C.byref(first)))
        assert first.value == 0
        self.set_plan(plan.bins, plan.segs, plan.names, seed)

    def genome_read(self, start, n):
        """Decode store bases [start, start+n) back to upper-case ASCII (diagnostic)."""
        buf = C.create_string_buffer(int(n))
        _ck(lib().ssc_genome_read(self.h, int(start), int(n), buf))
        return buf.raw[:n]

    def genome_size(self):
        v = C.c_uint64()
        _ck(lib().ssc_genome_size(self.h, C.byref(v)))
        return v.value

    def gc_census(self, starts, lens):
        """G/C and non-ACGT base counts of haplotype-store intervals (ssc_gc_census)."""
        starts = np.ascontiguousarray(starts, dtype=np.int64)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        gc = np.zeros(len(starts), np.int32)
        nn = np.zeros(len(starts), np.int32)
        _ck(lib().ssc_gc_census(self.h, starts.ctypes.data, lens.ctypes.data, len(starts), gc.ctypes.data, nn.ctypes.data))
        return gc, nn

    def set_plan(self, bins, segs, names, seed):
        bins = np.ascontiguousarray(bins)
        segs = np.ascontiguousarray(segs)
        pp, ep = C.c_int64(), C.c_int64()
        _ck(lib().ssc_set_plan(self.h, seed, bins.ctypes.data, bins.size, segs.ctypes.data, segs.size,
                               names, len(names), C.byref(pp), C.byref(ep)))
        self.planned, self.emitted = pp.value, ep.value
        return self.planned, self.emitted

    def generate(self, lo=0, hi=None, sink=None, user=None):
        """Stream pairs [lo, hi) through the pinned host slabs. Without `sink` returns (fq1, fq2) bytes.  `sink`: a Python
        callable or a C function (an abi.SINK_FN instance, e.g. the file writer of libsimuscop_host) called with `user`."""
        if hi is None:
            hi = self.planned
        parts1, parts2 = [], []

        def collect(user, b1, l1, b2, l2, first, n):
            parts1.append(C.string_at(b1, l1))
            if b2:
                parts2.append(C.string_at(b2, l2))
            return 0
        cb = sink if isinstance(sink, abi.SINK_FN) else abi.SINK_FN(sink if sink is not None else collect)
        _ck(lib().ssc_generate(self.h, lo, hi, cb, user))
        if sink is None:
            return b"".join(parts1), b"".join(parts2)

    def generate_device(self, lo=0, hi=None):
        if hi is None:
            hi = self.planned
        b1, b2, nb, ms = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_double()
        _ck(lib().ssc_generate_device(self.h, lo, hi, C.byref(b1), C.byref(b2), C.byref(nb), C.byref(ms)))
        return dict(bytes1=b1.value, bytes2=b2.value, bases=nb.value, device_ms=ms.value)

    def issue_floor(self, mode, read_length, n_pairs, reps=3):
        """Mean launch time (ms) of the issue-rate microbenchmark (ssc_issue_floor)."""
        ms = C.c_double()
        _ck(lib().ssc_issue_floor(self.h, mode, read_length, n_pairs, reps, C.byref(ms)))
        return ms.value

    def stats(self):
        s = abi.Stats()
        _ck(lib().ssc_get_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in abi.Stats._fields_}

    def reset_stats(self):
        _ck(lib().ssc_reset_stats(self.h))
