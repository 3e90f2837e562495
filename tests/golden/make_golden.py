#!/usr/bin/env python3
"""Regenerates tests/golden/ from the Philox-instrumented reference (oracle/_ref/simuReads_philox).

Run in the build container (needs /root/reference to have been compiled by oracle/build_ref.py).
Writes, per scenario of tests/helpers.py:
  golden.json              sha256 of every plan dump and FASTQ file the instrumented reference produced
  <small scenario>.plan.gz / _1.fq.gz / _2.fq.gz   full fixtures for the two smallest scenarios
"""
import gzip
import hashlib
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from simuscop_b200 import planfile  # noqa: E402

FULL = ["pe_tiny", "se_mini"]
helpers.SCENARIOS.setdefault("se_mini", dict(lengths=[30000], profile="GAIIx", layout="SE", coverage=2, insertSize=250, n_runs=1))


def sha(b):
    return hashlib.sha256(b).hexdigest()


def main():
    out = {}
    with tempfile.TemporaryDirectory() as wd:
        for name in sorted(helpers.SCENARIOS):
            if name in ("smoke", "qb"):
                continue
            scn = helpers.build_scenario(name, wd)
            plans, odir = helpers.run_reference_philox(scn)
            ent = {"seed": scn["seed"], "samples": []}
            for i, pf in enumerate(plans):
                plan = planfile.read_plan(pf)
                f1, f2 = helpers.sample_files(odir, plan, i, scn)
                b1, b2 = helpers.read_file(f1), helpers.read_file(f2)
                pb = open(pf, "rb").read()
                ent["samples"].append({"plan_sha256": sha(pb), "fq1_sha256": sha(b1), "fq2_sha256": sha(b2),
                                       "fq1_bytes": len(b1), "fq2_bytes": len(b2), "planned_pairs": plan.planned_pairs()})
                if name in FULL and i == 0:
                    for suffix, data in ((".plan.gz", pb), ("_1.fq.gz", b1), ("_2.fq.gz", b2)):
                        if data:
                            with gzip.GzipFile(os.path.join(HERE, name + suffix), "wb", mtime=0) as g:
                                g.write(data)
            out[name] = ent
            print(name, ent["samples"][0]["fq1_bytes"])
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
