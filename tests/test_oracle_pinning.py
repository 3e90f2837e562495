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
@pytest.mark.parametrize("name", ["pe_xten", "se_gaiix", "pe_tiny", "pe_wes", "pe_ploidy3", "se_ploidy1", "pe_iupac"])
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


def test_windowed_oracle_and_flat_plan_dump(built, workdir):
    """The two pieces the bench-scale GPU parity test stands on, checked on a small job against the instrumented reference:
    (1) the flat dump of the C++ front end (ssc_bin / ssc_segment / name arrays, no haplotype strings) equals the flattened
    full dump; (2) the oracle run on a pair sub-range with only the window of the store those pairs can touch, cut from the
    FASTA file, gives exactly the bytes of that range in the reference's files."""
    import ctypes as C
    import os
    import numpy as np
    from oracle import binding as oracle_binding
    from simuscop_b200 import host_binding, synth
    scn = helpers.build_scenario("pe_xten", workdir)
    plans, out = helpers.run_reference_philox(scn, tag="win")
    full = planfile.read_plan(plans[0])
    r1p, r2p = helpers.sample_files(out, full, 0, scn)
    r1, r2 = helpers.read_file(r1p), helpers.read_file(r2p)
    d = scn["dir"]
    cfg = os.path.join(d, "cfg_flat.txt")
    synth.write_config(cfg, output=os.path.join(d, "out_flat"), **scn["kw"])
    job = host_binding.Job(cfg, scn["seed"])
    flat_path = os.path.join(d, "plan.flat")
    rc = host_binding.lib().ssh_prepare_sample(job.j, 0, None, flat_path.encode(), C.byref(C.c_int64()), C.byref(C.c_int64()))
    assert rc == 0
    job.close()
    flat = planfile.read_plan(flat_path)
    assert flat.genome is None
    assert (flat.bins == full.bins).all() and (flat.segs == full.segs).all() and flat.names == full.names
    store = helpers.FastaStoreWindow(flat, scn["kw"]["ref"])
    assert (store.window(0, len(full.genome)) == np.frombuffer(bytes(full.genome), np.uint8)).all()
    n = flat.planned_pairs()
    whole1, whole2, _ = oracle_binding.generate(full, scn["seed"])
    assert (whole1, whole2) == (r1, r2)
    cuts = [0, 5, n // 3, n // 3 + 700, n - 3, n]
    got1, got2 = b"", b""
    for a, b in zip(cuts[:-1], cuts[1:]):
        wlo, whi, _, _ = helpers.pair_window(flat, a, b)
        assert whi - wlo < len(full.genome) or b - a > n // 2
        o1, o2, info = oracle_binding.generate(flat, scn["seed"], a, b, genome=store.window(wlo, whi), genome_first=wlo)
        got1 += o1; got2 += o2
    assert (got1, got2) == (r1, r2)
    with pytest.raises(RuntimeError):      # a window that misses the fragments is an error, never a silent read elsewhere
        oracle_binding.generate(flat, scn["seed"], 0, 50, genome=store.window(0, 100), genome_first=0)


def test_shipped_wes_config_host_plan_and_oracle(built, workdir):
    """BASELINE.json configs[1] on the CPU side: the reference's shipped config_test_wes.txt (4 677 capture targets at 100x,
    SNPs + variations, HiSeq2500 PE125, synthetic 63 Mb chr20) through the C++ front end in plan-only mode and the C oracle
    reproduces the FASTQ files the instrumented reference wrote for it (tests/golden/shipped.json) -- the same files the
    CUDA CLI must reproduce on the GPU box (tests/test_gpu_shipped_configs.py)."""
    from simuscop_b200 import paths
    gold = json.load(open(os.path.join(GOLD, "shipped.json")))["configs"]["wes"]
    root = helpers.build_shipped_tree(workdir)
    prefix = os.path.join(workdir, "shipped_wes_plan")
    helpers.run_shipped(paths.SIMUREADS, root, "wes", "plan", {"SIMUSCOP_PLAN_ONLY": "1", "SIMUSCOP_DUMP_PLAN": prefix})
    plan = planfile.read_plan(prefix + ".0.plan")
    o1, o2, _ = oracle_binding.generate(plan, helpers.SHIPPED_SEED)
    assert len(o1) == gold["test_1.fq"]["bytes"] and hashlib.sha256(o1).hexdigest() == gold["test_1.fq"]["sha256"]
    assert len(o2) == gold["test_2.fq"]["bytes"] and hashlib.sha256(o2).hexdigest() == gold["test_2.fq"]["sha256"]
