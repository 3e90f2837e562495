#!/usr/bin/env python3
"""Runs the reference's three shipped test configurations (tests/helpers.SHIPPED: the key/value pairs of
configFiles/config_test_{wgs,wes,tumor}.txt, the shipped variations / SNP / BED / abundance / profile files, a seeded synthetic
63 025 520-bp chr20 for the missing ref.fa.gz) through the Philox-instrumented reference (oracle/_ref/simuReads_philox) and
records size + sha256 of every FASTQ file in tests/golden/shipped.json.  Run by hand in the build container (about 15 minutes
of CPU); tests/test_gpu_shipped_configs.py asserts that the CUDA CLI reproduces these files on the GPU box."""
import hashlib
import json
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from simuscop_b200 import paths  # noqa: E402


def sha_file(p):
    h = hashlib.sha256()
    n = 0
    with open(p, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b:
                break
            h.update(b)
            n += len(b)
    return h.hexdigest(), n


def main():
    wd = sys.argv[1] if len(sys.argv) > 1 else tempfile.mkdtemp(prefix="shipped_")
    root = helpers.build_shipped_tree(wd)
    out = {"seed": helpers.SHIPPED_SEED, "chr20_bp": helpers.SHIPPED_CHR20, "configs": {}}
    for which in ("wes", "wgs", "tumor"):
        t0 = time.time()
        files = helpers.run_shipped(paths.REF_PHILOX, root, which, "ref")
        ent = {}
        for fn, p in files.items():
            s, n = sha_file(p)
            ent[fn] = {"sha256": s, "bytes": n}
        out["configs"][which] = ent
        print(which, "%.0f s" % (time.time() - t0), {k: v["bytes"] for k, v in ent.items()}, flush=True)
        with open(os.path.join(HERE, "shipped.json"), "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
