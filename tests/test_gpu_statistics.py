"""Statistical parity against the UNMODIFIED reference (its own mt19937 RNG, wall-clock seeds): error rates per
cycle bin, quality histograms, read-length (indel) histogram, insert-size distribution and coverage must agree.
Tolerances: chi-square p > 1e-4 per statistic after pooling cells with expected count >= 5 (two independent
samples of ~1.2e7 bases each); KS p > 1e-4 for the insert size."""
import os
import subprocess

import numpy as np
import pytest
from scipy import stats

import helpers
from simuscop_b200 import paths, synth

pytestmark = pytest.mark.gpu

RL = 151
COMP = bytes.maketrans(b"ACGTN", b"TGCAN")


def _parse(fq):
    lines = fq.split(b"\n")
    names, seqs, quals = lines[0::4], lines[1::4], lines[3::4]
    n = len(seqs) - (1 if seqs and seqs[-1] == b"" else 0)
    return names[:n], seqs[:n], quals[:n]


def _collect(out, genome):
    """Statistics of one run: mismatches per cycle bin (reads without indels), quality histogram per cycle bin,
    read-length histogram, insert sizes (pairs whose mates both kept RL), per-10kb coverage of read-1 starts."""
    n1, s1, q1 = _parse(helpers.read_file(os.path.join(out, "test_1.fq")))
    n2, s2, q2 = _parse(helpers.read_file(os.path.join(out, "test_2.fq")))
    assert len(s1) == len(s2) > 1000
    g = genome
    kmer = {}
    K = 24
    gb = g.tobytes()
    for i in range(0, len(gb) - K):
        kmer.setdefault(gb[i:i + K], i)
    bins = 10
    mism = np.zeros((2, bins), np.int64)
    tot = np.zeros((2, bins), np.int64)
    qh = np.zeros((2, bins, 94), np.int64)
    lenh = np.zeros((2, 64), np.int64)
    isz = []
    cov = np.zeros(len(gb) // 10000 + 1, np.int64)
    for name, a, b, qa, qb in zip(n1, s1, s2, q1, q2):
        pos = int(name.split(b"#")[2])
        cov[pos // 10000] += 1
        for mate, (s, q) in enumerate(((a, qa), (b, qb))):
            lenh[mate, min(63, max(0, len(s) - RL + 32))] += 1
            qq = np.frombuffer(q, np.uint8) - 33
            cyc = (np.arange(len(s)) * bins // len(s))
            np.add.at(qh[mate], (cyc, qq), 1)
        if len(a) == RL:
            ref = np.frombuffer(gb[pos:pos + RL], np.uint8)
            rd = np.frombuffer(a, np.uint8)
            if len(ref) == RL:
                cyc = np.arange(RL) * bins // RL
                np.add.at(tot[0], cyc, 1)
                np.add.at(mism[0], cyc, (ref != rd).astype(np.int64))
        if len(a) == RL and len(b) == RL:
            rc = b.translate(COMP)[::-1]            # read 2 back on the forward strand = last RL bases of the fragment
            hit = kmer.get(rc[-K:])                 # last K bases of the fragment (cycles 0..K-1 of read 2, low error)
            if hit is not None:
                end = hit + K
                isz.append(end - pos)
                ref = np.frombuffer(gb[end - RL:end], np.uint8)
                rd = np.frombuffer(rc, np.uint8)
                if len(ref) == RL:
                    cyc = (RL - 1 - np.arange(RL)) * bins // RL
                    np.add.at(tot[1], cyc, 1)
                    np.add.at(mism[1], cyc, (ref != rd).astype(np.int64))
    return dict(mism=mism, tot=tot, qh=qh, lenh=lenh, isz=np.array(isz), cov=cov, pairs=len(s1))


def _chi2(a, b):
    """Two-sample chi-square homogeneity test on count vectors, pooling sparse cells."""
    a, b = np.asarray(a, float).ravel(), np.asarray(b, float).ravel()
    keep = (a + b) >= 10
    a2 = np.append(a[keep], a[~keep].sum())
    b2 = np.append(b[keep], b[~keep].sum())
    m = (a2 + b2) > 0
    if m.sum() < 2:
        return 1.0
    return stats.chi2_contingency(np.vstack([a2[m], b2[m]]))[1]


def test_distributions_match_unmodified_reference(built, workdir):
    if not os.path.exists(paths.REF_PLAIN):
        pytest.skip("unmodified reference binary not built")
    helpers.SCENARIOS["stat"] = dict(lengths=[400000], names=["chr1"], profile="XTen", layout="PE", coverage=30, insertSize=300)
    scn = helpers.build_scenario("stat", workdir)
    d = scn["dir"]
    genome = np.frombuffer(b"".join(l.strip() for l in open(os.path.join(d, "ref.fa"), "rb").read().split(b"\n")[1:]), np.uint8)
    def run(tag, binp, env):
        out = os.path.join(d, "out_stat_" + tag)
        cfg = os.path.join(d, "cfg_stat_%s.txt" % tag)
        # threads = 1: the reference's worker threads share mutable Profile state without a lock (e.g. Profile::getKmerIndx
        # uses map::operator[], lib/profile/Profile.cpp:223); a multi-threaded run now and then skews the mate-2 error rate
        synth.write_config(cfg, output=out, **dict(scn["kw"], threads=1))
        r = subprocess.run([binp, cfg], env=dict(os.environ, **env), capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        return _collect(out, genome)
    ours = run("ours", paths.SIMUREADS, {"SIMUSCOP_SEED": "99"})
    # The reference seeds itself from the wall clock, so every run of it is a fresh sample, while ours (seed 99) is one
    # fixed sample: with ~30 chi-square tests at alpha = 1e-4 a true-null comparison fails about once in 300 runs.  A
    # failing comparison is therefore repeated against up to two more reference samples (a real defect fails them all).
    last = None
    for attempt in range(3):
        try:
            _compare(run("ref%d" % attempt, paths.REF_PLAIN, {}), ours)
            return
        except AssertionError as e:
            last = e
    raise last


def _compare(a, b):
    alpha = 1e-4
    assert abs(a["pairs"] - b["pairs"]) <= 0.02 * a["pairs"]
    # substitution (mismatch) counts per cycle bin, both mates
    for mate in range(2):
        p = _chi2(np.stack([a["mism"][mate], a["tot"][mate] - a["mism"][mate]]).T.ravel(),
                  np.stack([b["mism"][mate], b["tot"][mate] - b["mism"][mate]]).T.ravel())
        assert p > alpha, ("mismatch rate per cycle bin, mate %d" % (mate + 1), p)
        ra, rb = a["mism"][mate].sum() / a["tot"][mate].sum(), b["mism"][mate].sum() / b["tot"][mate].sum()
        assert abs(ra - rb) < 0.05 * max(ra, rb), (ra, rb)
    # quality histograms per (mate, cycle bin)
    for mate in range(2):
        for c in range(a["qh"].shape[1]):
            p = _chi2(a["qh"][mate, c], b["qh"][mate, c])
            assert p > alpha, ("quality histogram mate %d bin %d" % (mate + 1, c), p)
    # read-length histogram = indel length distribution
    for mate in range(2):
        assert _chi2(a["lenh"][mate], b["lenh"][mate]) > alpha
    # insert size
    assert len(a["isz"]) > 1000 and len(b["isz"]) > 1000
    assert stats.ks_2samp(a["isz"], b["isz"])[1] > alpha
    assert abs(a["isz"].mean() - b["isz"].mean()) < 1.0
    # coverage per 10 kb window: same GC-weighted plan model, independent draws -> compare dispersion-normalised totals
    assert abs(a["cov"].sum() - b["cov"].sum()) <= 0.02 * a["cov"].sum()
    corr = np.corrcoef(a["cov"][:-1], b["cov"][:-1])[0, 1]
    assert corr > -0.5   # independent GC-factor draws per window: no systematic anti-correlation / empty regions
    assert (b["cov"][:-1] > 0).all() and (a["cov"][:-1] > 0).all()
