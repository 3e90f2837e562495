"""The C oracle pinned against (a) committed golden fixtures made by the instrumented reference
(tests/golden/make_golden.py) and (b) the instrumented reference run live when its binary exists."""
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

import helpers
from oracle import binding as oracle_binding
from simuscop_b200 import planfile
from simuscop_b200.paths import REF_PHILOX

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gunzip(path, dst):
    with gzip.open(path, "rb") as f, open(dst, "wb") as g:
        g.write(f.read())
    return dst


def test_philox_known_answers(built):
    # Random123 kat_vectors, philox4x32-10
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        assert [int(x) for x in oracle_binding.philox(ctr, key)] == want


@pytest.mark.parametrize("name", ["pe_tiny", "se_mini"])
def test_oracle_matches_golden_fixture(name, built, tmp_path):
    gold = json.load(open(os.path.join(GOLD, "golden.json")))[name]
    plan = planfile.read_plan(_gunzip(os.path.join(GOLD, name + ".plan.gz"), str(tmp_path / "p.plan")))
    f1, f2, info = oracle_binding.generate(plan, gold["seed"])
    want1 = gzip.open(os.path.join(GOLD, name + "_1.fq.gz"), "rb").read()
    assert f1 == want1
    if plan.paired:
        assert f2 == gzip.open(os.path.join(GOLD, name + "_2.fq.gz"), "rb").read()
    s = gold["samples"][0]
    assert hashlib.sha256(f1).hexdigest() == s["fq1_sha256"]
    assert plan.planned_pairs() == s["planned_pairs"]
    # abandon rule exercised: fewer pairs emitted than planned in the tiny scenario
    if name == "pe_tiny":
        assert info["emitted"] < plan.planned_pairs()


def test_oracle_pair_ranges_concatenate(built, tmp_path):
    plan = planfile.read_plan(_gunzip(os.path.join(GOLD, "pe_tiny.plan.gz"), str(tmp_path / "p.plan")))
    whole1, whole2, _ = oracle_binding.generate(plan, 7)
    n = plan.planned_pairs()
    cuts = [0, 1, n // 2, n - 3, n]
    parts = [oracle_binding.generate(plan, 7, a, b) for a, b in zip(cuts[:-1], cuts[1:])]
    assert b"".join(p[0] for p in parts) == whole1
    assert b"".join(p[1] for p in parts) == whole2


@pytest.mark.skipif(not os.path.exists(REF_PHILOX), reason="instrumented reference binary not built")
@pytest.mark.parametrize("name", ["pe_xten", "se_gaiix", "pe_tiny", "pe_wes", "pe_ploidy3", "se_ploidy1"])
def test_oracle_matches_live_instrumented_reference(name, built, workdir):
    scn = helpers.build_scenario(name, workdir)
    plans, out = helpers.run_reference_philox(scn, tag="pin")
    gold = json.load(open(os.path.join(GOLD, "golden.json")))[name]
    for i, pf in enumerate(plans):
        plan = planfile.read_plan(pf)
        r1p, r2p = helpers.sample_files(out, plan, i, scn)
        r1, r2 = helpers.read_file(r1p), helpers.read_file(r2p)
        f1, f2, _ = oracle_binding.generate(plan, scn["seed"])
        assert f1 == r1 and f2 == r2
        # and the live run reproduces the committed hashes (the instrumented build is deterministic)
        assert hashlib.sha256(r1).hexdigest() == gold["samples"][i]["fq1_sha256"]
        assert hashlib.sha256(open(pf, "rb").read()).hexdigest() == gold["samples"][i]["plan_sha256"]


@pytest.mark.skipif(not os.path.exists(REF_PHILOX), reason="instrumented reference binary not built")
@pytest.mark.parametrize("seed", [2, 16, 19, 25, 31, 44])
def test_oracle_matches_live_instrumented_reference_on_random_jobs(seed, built, workdir):
    """Fuzz pin: seeded random jobs (capture targets, SNPs, tumour mixtures, CNV / indel / SNV variations, 1..3 chromosomes, PE and
    SE, all four shipped profiles): every FASTQ byte of the oracle equals the instrumented reference's, for every sample."""
    scn, mode = helpers.build_random_job_scenario(seed, workdir)
    plans, out = helpers.run_reference_philox(scn, tag="pinfz")
    assert plans, mode
    for i, pf in enumerate(plans):
        plan = planfile.read_plan(pf)
        r1p, r2p = helpers.sample_files(out, plan, i, scn)
        f1, f2, info = oracle_binding.generate(plan, scn["seed"])
        assert f1 == helpers.read_file(r1p) and f2 == helpers.read_file(r2p), (seed, mode, i)


@pytest.mark.skipif(not os.path.exists(REF_PHILOX), reason="instrumented reference binary not built")
@pytest.mark.parametrize("name", sorted(helpers.STRESS))
def test_oracle_and_host_on_synthetic_profiles(name, built, workdir):
    """Synthetic profiles with odd k-mer sizes, short/long reads, heavy indels, degenerate rows: the oracle and the
    C++ front end's plan (tables included) against the live instrumented reference."""
    import glob
    import subprocess
    from simuscop_b200 import paths, synth
    scn = helpers.build_stress(name, workdir)
    plans, out = helpers.run_reference_philox(scn, tag="st")
    plan = planfile.read_plan(plans[0])
    r1p, r2p = helpers.sample_files(out, plan, 0, scn)
    f1, f2, info = oracle_binding.generate(plan, scn["seed"])
    assert f1 == helpers.read_file(r1p) and f2 == helpers.read_file(r2p) and info["emitted"] > 100
    d = scn["dir"]
    cfg = os.path.join(d, "cfg_plan.txt")
    synth.write_config(cfg, output=os.path.join(d, "out_plan"), **scn["kw"])
    env = dict(os.environ, SIMUSCOP_SEED=str(scn["seed"]), SIMUSCOP_DUMP_PLAN=os.path.join(d, "plan_ours"), SIMUSCOP_PLAN_ONLY="1")
    r = subprocess.run([paths.SIMUREADS, cfg], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert open(os.path.join(d, "plan_ours.0.plan"), "rb").read() == open(plans[0], "rb").read()
