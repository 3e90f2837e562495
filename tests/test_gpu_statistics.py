"""Statistical parity against the UNMODIFIED reference (its own mt19937 streams, wall-clock seeds), SURVEY.md 8c:
everything Profile::predict and Segment::yieldReads sample must follow the same distributions in both programs.

Per profile (all four shipped ones), on a 400 kb chromosome at 30x PE (1.2e7 bases per run), tests/stat_helpers.py counts
  * substitutions per (k-mer context row x position bin x called base), both mates   -- 100 chi-square tests
  * quality symbols per (reference base, called base, position bin)                   -- 212 tests
  * reads without / with net insertion / with net deletion, insertion and deletion length histograms, per mate
  * insert sizes (histogram chi-square + two-sample KS)
and compares our run (CUDA CLI) with the reference's cell by cell (two-sample chi-square homogeneity, sparse cells pooled).
Tolerance: every p > 0.001 / number of tests (Bonferroni, family-wise level 0.001).  Coverage: the per-kb (WGS) and per-target
(WES) read counts follow the GC-weight model -- compared per GC percentage by two-sample KS; and our WES run places exactly
the planned number of fragments on every capture target."""
import os
import subprocess

import numpy as np
import pytest
from scipy import stats

import helpers
import stat_helpers
from simuscop_b200 import paths, planfile, synth, testdata

pytestmark = pytest.mark.gpu

PROFILE_JOBS = {"XTen": (151, 300), "GAIIx": (74, 250), "HiSeq2000": (75, 250), "HiSeq2500": (125, 200)}


def _genome(d):
    return np.frombuffer(b"".join(l.strip() for l in open(os.path.join(d, "ref.fa"), "rb").read().split(b"\n")[1:]), np.uint8)


def _run(binp, scn, tag, env, threads=1, extra=None):
    d = scn["dir"]
    out = os.path.join(d, "out_stat_" + tag)
    cfg = os.path.join(d, "cfg_stat_%s.txt" % tag)
    # threads = 1 for the reference: its worker threads share mutable Profile state without a lock (e.g. Profile::getKmerIndx
    # uses map::operator[], lib/profile/Profile.cpp:223); a multi-threaded run now and then skews the mate-2 error rate
    synth.write_config(cfg, output=out, **dict(scn["kw"], threads=threads, **(extra or {})))
    r = subprocess.run([binp, cfg], env=dict(os.environ, **env), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return out


@pytest.mark.parametrize("profile", sorted(PROFILE_JOBS))
def test_profile_tables_match_unmodified_reference(profile, built, workdir):
    if not os.path.exists(paths.REF_PLAIN):
        pytest.skip("unmodified reference binary not built")
    RL, insert = PROFILE_JOBS[profile]
    name = "stat_" + profile
    helpers.SCENARIOS[name] = dict(lengths=[400000], names=["chr1"], profile=profile, layout="PE", coverage=30, insertSize=insert)
    scn = helpers.build_scenario(name, workdir)
    genome = _genome(scn["dir"])

    def stats_of(out):
        return stat_helpers.collect(os.path.join(out, "test_1.fq"), os.path.join(out, "test_2.fq"), genome, RL, 50)
    ours = stats_of(_run(paths.SIMUREADS, scn, "ours", {"SIMUSCOP_SEED": "99"}))
    assert ours["sub"].sum() > 0.6 * 2 * 400000 * 30 / 2 and len(ours["isize"]) > 10000
    # The reference seeds itself from the wall clock: every run of it is a fresh sample, ours (seed 99) a fixed one.  At a
    # family-wise level of 0.001 a true-null comparison fails once in a thousand runs; a failing comparison is repeated
    # against up to two more reference samples (a real defect fails them all).
    last = None
    for attempt in range(3):
        ref = stats_of(_run(paths.REF_PLAIN, scn, "ref%d" % attempt, {}))
        assert abs(ref["pairs"] - ours["pairs"]) <= 0.02 * ours["pairs"]
        last = stat_helpers.compare(ours, ref)
        if not last:
            return
    raise AssertionError("statistics differ from the unmodified reference: %s" % last[:5])


def _gc_percent_windows(genome, starts, lens):
    up = genome
    gc = np.concatenate(([0], np.cumsum((up == ord("G")) | (up == ord("C")))))
    nn = np.concatenate(([0], np.cumsum(up == ord("N"))))
    g = gc[starts + lens] - gc[starts]
    n = nn[starts + lens] - nn[starts]
    return np.where(n > 0, -1, 100 * g // np.maximum(lens, 1))


def _ks_by_gc(counts_a, counts_b, gcp, min_windows=8):
    """per GC percentage: two-sample KS of the per-window counts; returns (number of groups, failures at Bonferroni 0.001)"""
    groups = [g for g in np.unique(gcp) if g >= 0 and (gcp == g).sum() >= min_windows]
    fails = []
    for g in groups:
        p = stats.ks_2samp(counts_a[gcp == g], counts_b[gcp == g])[1]
        if not (p > 1e-3 / len(groups)):
            fails.append((int(g), float(p)))
    return len(groups), fails


def test_wgs_coverage_follows_the_gc_weight_model(built, workdir):
    """Reads per 1 kb window: floor(weight x reads / total weight) with weight = the sum over the haplotypes of a
    normal_distribution(gcMeans[gc], gcStd) draw (Segment::getWeightedLength, Segment.cpp:567-600; Profile::getGCFactor,
    Profile.cpp:1507-1517).  The draws are independent between the programs, so windows are compared per GC percentage."""
    if not os.path.exists(paths.REF_PLAIN):
        pytest.skip("unmodified reference binary not built")
    # GC content varying along the chromosome (20 % ... 80 %) so that many GC percentages occur
    name = "stat_cov"
    helpers.SCENARIOS[name] = dict(lengths=[1000000], names=["chr1"], profile="XTen", layout="PE", coverage=30, insertSize=300)
    scn = helpers.build_scenario(name, workdir)
    d = scn["dir"]
    rng = np.random.default_rng(8)
    n = 1000000
    pgc = np.repeat(np.clip(0.5 + 0.3 * np.sin(np.arange(n // 1000) / 40.0) + rng.normal(0, 0.02, n // 1000), 0.15, 0.85), 1000)
    u = rng.random(n)
    seq = np.where(u < pgc / 2, ord("G"), np.where(u < pgc, ord("C"), np.where(u < pgc + (1 - pgc) / 2, ord("A"), ord("T")))).astype(np.uint8)
    synth.write_fasta(os.path.join(d, "ref.fa"), [("chr1", seq)])
    for f in (os.path.join(d, "ref.fa.fai"),):
        if os.path.exists(f):
            os.remove(f)
    starts = np.arange(0, n, 1000)
    gcp = _gc_percent_windows(seq, starts, np.full(len(starts), 1000))

    def counts(out):
        names, _, _ = stat_helpers.read_fastq(os.path.join(out, "test_1.fq"))
        pos = np.array([int(x.split(b"#")[2]) for x in names], np.int64)
        return np.bincount(pos // 1000, minlength=len(starts))[:len(starts)]
    ours = counts(_run(paths.SIMUREADS, scn, "ours", {"SIMUSCOP_SEED": "5"}))
    last = None
    for attempt in range(3):
        ref = counts(_run(paths.REF_PLAIN, scn, "ref%d" % attempt, {}))
        assert abs(int(ours.sum()) - int(ref.sum())) <= 0.02 * ours.sum()
        ng, last = _ks_by_gc(ours, ref, gcp)
        assert ng >= 20
        # and the GC response itself: mean count per GC percentage agrees within 6 standard errors
        for g in np.unique(gcp):
            m = gcp == g
            if m.sum() >= 8:
                se = np.sqrt(ours[m].var() / m.sum() + ref[m].var() / m.sum()) + 1e-9
                if abs(ours[m].mean() - ref[m].mean()) > 6 * se:
                    last = last + [("mean", int(g), float(ours[m].mean()), float(ref[m].mean()))]
        if not last:
            return
    raise AssertionError("per-window coverage differs from the unmodified reference: %s" % last[:5])


def test_wes_coverage_per_target(built, workdir):
    """Capture mode: (1) our run places exactly the planned number of fragments on every target piece (ceil(readCount / 2) pairs
    per bin and haplotype, Segment.cpp:741-848), counted from the read names against the plan dump; (2) fragments per target
    and base follow the same GC-weight model as the unmodified reference's (two-sample KS per GC percentage)."""
    if not os.path.exists(paths.REF_PLAIN):
        pytest.skip("unmodified reference binary not built")
    name = "stat_wes"
    helpers.SCENARIOS[name] = dict(lengths=[900000], names=["chr20"], profile="HiSeq2500", layout="PE", coverage=120, insertSize=200)
    scn = helpers.build_scenario(name, workdir)
    d = scn["dir"]
    genome = _genome(d)
    rng = np.random.default_rng(4)
    # well separated targets (padding 50 on both sides, pieces of at most 1000 bases: Genome.cpp:270-279, 684-739)
    bed, p = [], 2000
    while p < 890000:
        ln = int(rng.integers(120, 900))
        bed.append((p, p + ln))
        p += ln + int(rng.integers(1500, 4000))
    with open(os.path.join(d, "targets.bed"), "w") as f:
        f.write("".join("chr20\t%d\t%d\n" % b for b in bed))
    scn["kw"]["target"] = os.path.join(d, "targets.bed")
    dump = os.path.join(d, "wes_plan")
    out = _run(paths.SIMUREADS, scn, "ours", {"SIMUSCOP_SEED": "11", "SIMUSCOP_DUMP_PLAN": dump})
    plan = planfile.read_plan(dump + ".0.plan")
    b = plan.bins[plan.bins["read_count"] > 0]
    pieces = sorted(set(zip(b["spos"].tolist(), b["epos"].tolist())))
    starts = np.array([s for s, _ in pieces]); ends = np.array([e for _, e in pieces])
    assert (starts[1:] > ends[:-1]).all()
    want = np.zeros(len(pieces), np.int64)
    where = {pc: i for i, pc in enumerate(pieces)}
    for s, e, rc in zip(b["spos"].tolist(), b["epos"].tolist(), b["read_count"].tolist()):
        want[where[(s, e)]] += (rc + 1) // 2

    def per_piece(o):
        names, _, _ = stat_helpers.read_fastq(os.path.join(o, "test_1.fq"))
        pos = np.array([int(x.split(b"#")[2]) for x in names], np.int64)
        k = np.searchsorted(starts, pos, side="right") - 1
        assert ((k >= 0) & (pos <= ends[np.maximum(k, 0)])).all(), "a fragment starts outside every capture target"
        return np.bincount(k, minlength=len(pieces))
    got = per_piece(out)
    assert (got == want).all()
    lens = ends - starts + 1
    gcp = _gc_percent_windows(genome, starts, lens)
    last = None
    for attempt in range(3):
        ref = per_piece(_run(paths.REF_PLAIN, scn, "ref%d" % attempt, {}))
        assert abs(int(got.sum()) - int(ref.sum())) <= 0.02 * got.sum()
        ng, last = _ks_by_gc(got / lens, ref / lens, gcp, min_windows=6)
        assert ng >= 5
        if not last:
            return
    raise AssertionError("per-target coverage differs from the unmodified reference: %s" % last[:5])
