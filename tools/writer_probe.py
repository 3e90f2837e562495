#!/usr/bin/env python3
"""Rate of the drop-in CLI's FASTQ file writer (libsimuscop_host ssh_writer_*) alone: 4 x 2 x 256 MB of slabs from host
memory into two files, per directory and number of writer threads.  Tells what the end-to-end legs that write files can
reach on this box at most."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from simuscop_b200 import abi, host_binding  # noqa: E402

hl = host_binding.lib()
hl.ssh_writer_open.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
hl.ssh_writer_sink.restype = C.c_void_p
hl.ssh_writer_close.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
sink = C.cast(hl.ssh_writer_sink(), abi.SINK_FN)
buf = np.random.default_rng(1).integers(0, 256, 256 << 20, dtype=np.uint8).tobytes()
for d in sys.argv[1:] or ["/dev/shm", "/tmp"]:
    for rep, mode, ths in ((0, "pwrite", (1, 2)), (1, "pwrite", (1, 2)), (0, "mmap", (4, 8, 16)), (0, "hybrid", (3, 4, 6, 8, 12, 16)),
                           (1, "hybrid", (3, 4, 6, 8, 12, 16))):
        os.environ["SIMUSCOP_WRITER_MODE"] = mode          # pwrite: one stream per file; mmap: chunk pool into mappings; hybrid: both
        for th in ths:
            w = C.c_void_p()
            p1, p2 = os.path.join(d, "wprobe_1"), os.path.join(d, "wprobe_2")
            assert hl.ssh_writer_open(p1.encode(), p2.encode(), th, C.byref(w)) == 0
            t = time.perf_counter()
            for i in range(4):
                assert sink(w, buf, len(buf), buf, len(buf), 0, 0) == 0
            hl.ssh_writer_close(w, None, None)
            dt = time.perf_counter() - t
            print(json.dumps({"dir": d, "mode": mode, "pass": rep, "threads": th, "GBps": round(8 * len(buf) / dt / 1e9, 2)}), flush=True)
            os.remove(p1); os.remove(p2)
