"""Parity at benchmark scale: the 3 Gb x 30x PE151 XTen job of bench.py (BASELINE.json configs[3]) through the C ABI against the
C oracle, on sampled pair ranges and one full 2 097 152-pair batch.

What only this size exercises: store indices beyond 2^32 bases (the diploid store holds 6.0 G bases), 6 M bins, the 32-bit blob
cursors and window-word indices of the fast kernel, batches of 2 097 152 pairs in which every warp works through a dozen
tickets, and fragment counters in the hundred thousands.  The oracle gets the plan's own bin / segment arrays (the flat dump of
the C++ front end) and a window of the haplotype store cut from the FASTA file, never from the device."""
import hashlib
import os

import numpy as np
import pytest

import helpers
from simuscop_b200 import planfile

pytestmark = pytest.mark.gpu

BATCH = 1 << 21


@pytest.fixture(scope="module")
def bench_job(built):
    import bench
    from simuscop_b200 import cuda_binding, host_binding
    wd = os.environ.get("SIMUSCOP_BENCH_DIR", "/tmp/simuscop_bench")
    os.makedirs(wd, exist_ok=True)
    cfg, genome_len = bench.write_job(wd, 3000000000, 30, "XTen")
    job = host_binding.Job(cfg, 1)
    gen = cuda_binding.Generator(0)
    gen.set_option("batch_pairs", BATCH)
    flat = os.path.join(wd, "bench_plan.flat")
    planned, emitted = job.prepare(0, gen, flat)
    plan = planfile.read_plan(flat)
    os.remove(flat)
    store = helpers.FastaStoreWindow(plan, os.path.join(wd, "genome_3000000000.fa"))
    yield dict(gen=gen, plan=plan, store=store, planned=planned, emitted=emitted)
    gen.close()
    job.close()


def oracle_range(job, lo, hi):
    from oracle import binding as oracle_binding
    plan = job["plan"]
    wlo, whi, _, _ = helpers.pair_window(plan, lo, hi)
    win = job["store"].window(wlo, whi)
    return oracle_binding.generate(plan, 1, lo, hi, genome=win, genome_first=wlo)


def test_store_and_plan_are_bench_sized(bench_job):
    assert bench_job["gen"].genome_size() > (1 << 32)
    assert bench_job["planned"] == bench_job["plan"].planned_pairs() > 290000000
    assert len(bench_job["plan"].bins) > 5000000


def sampled_ranges(job):
    """first / last pairs, the pairs around store index 2^32, both sides of the chromosome boundary nearest to it, a range
    across a batch boundary of a whole-job run, and two seeded random ranges"""
    plan, planned = job["plan"], job["planned"]
    rc = np.maximum(plan.bins["read_count"].astype(np.int64), 0)
    base = np.concatenate(([0], np.cumsum((rc + 1) // 2)))
    out = [("first", 0, 4096), ("last", planned - 4096, planned)]
    first_store = plan.bins["hap_base"] + plan.bins["spos"]
    k = int(np.searchsorted(first_store, 1 << 32))               # bins are in store order for this job
    mid = int(base[k])
    out.append(("store_2^32", mid - 2048, mid + 2048))
    ce = plan.bins["contig_end"]
    edges = np.flatnonzero(ce[1:] != ce[:-1]) + 1                # first bin of every contig
    e = int(edges[np.argmin(np.abs(first_store[edges] - (1 << 32)))])
    out.append(("contig_boundary", int(base[e]) - 2048, int(base[e]) + 2048))
    out.append(("batch_boundary", 37 * BATCH - 2048, 37 * BATCH + 2048))
    rng = np.random.default_rng(5)
    for i in range(2):
        a = int(rng.integers(0, planned - 4096))
        out.append(("random%d" % i, a, a + 4096))
    return out


def test_sampled_ranges_bit_exact(bench_job):
    gen = bench_job["gen"]
    for tag, lo, hi in sampled_ranges(bench_job):
        o1, o2, info = oracle_range(bench_job, lo, hi)
        f1, f2 = gen.generate(lo, hi)
        assert info["emitted"] == hi - lo
        assert helpers.first_diff(f1, o1) == -1, "%s [%d, %d): file 1 differs at byte %d" % (tag, lo, hi, helpers.first_diff(f1, o1))
        assert helpers.first_diff(f2, o2) == -1, "%s [%d, %d): file 2 differs at byte %d" % (tag, lo, hi, helpers.first_diff(f2, o2))


_FORK_JOB = None          # plan + store of the forked oracle workers (inherited, not pickled)


def _oracle_part(args):
    lo, hi = args
    o1, o2, _ = oracle_range(_FORK_JOB, lo, hi)
    return o1, o2


def test_full_batch_bit_exact(bench_job):
    """One whole 2 097 152-pair batch exactly as bench.py launches it (batch 71 of the job: store indices around 2.9 G), cut
    out of a three-batch ssc_generate call so that the slab double buffering is on the path; the oracle runs the batch in
    parallel slices on the host cores."""
    import multiprocessing as mp
    gen = bench_job["gen"]
    lo, hi = 71 * BATCH, 72 * BATCH
    parts = []
    sizes = []

    def sink(user, b1, l1, b2, l2, first, n):
        import ctypes as C
        sizes.append((first, n, l1, l2))
        if len(sizes) == 2:
            parts.append((C.string_at(b1, l1), C.string_at(b2, l2)))
        return 0
    gen.generate(lo - BATCH, hi + BATCH, sink=sink)
    assert [s[1] for s in sizes] == [BATCH] * 3
    f1, f2 = parts[0]
    nproc = min(32, os.cpu_count() or 1)
    step = (hi - lo) // (4 * nproc)
    cuts = list(range(lo, hi, step)) + [hi]
    global _FORK_JOB
    _FORK_JOB = dict(plan=bench_job["plan"], store=bench_job["store"])
    ctx = mp.get_context("fork")                      # the workers touch the oracle library and the FASTA file only, never CUDA
    with ctx.Pool(nproc) as pool:
        res = pool.map(_oracle_part, list(zip(cuts[:-1], cuts[1:])), chunksize=1)
    _FORK_JOB = None
    o1 = b"".join(r[0] for r in res)
    o2 = b"".join(r[1] for r in res)
    assert helpers.first_diff(f1, o1) == -1, "file 1 differs at byte %d" % helpers.first_diff(f1, o1)
    assert helpers.first_diff(f2, o2) == -1, "file 2 differs at byte %d" % helpers.first_diff(f2, o2)
    print("full batch sha256:", hashlib.sha256(f1).hexdigest(), hashlib.sha256(f2).hexdigest())


def test_device_resident_batch_equals_streamed_batch(bench_job):
    """ssc_generate_device (the leg bench.py times) leaves the same number of bytes and bases in HBM as the streamed call."""
    gen = bench_job["gen"]
    lo, hi = 5 * BATCH, 6 * BATCH
    tot = [0, 0]

    def sink(user, b1, l1, b2, l2, first, n):
        tot[0] += l1; tot[1] += l2
        return 0
    gen.reset_stats()
    gen.generate(lo, hi, sink=sink)
    bases = gen.stats()["bases_emitted"]
    r = gen.generate_device(lo, hi)
    assert (r["bytes1"], r["bytes2"], r["bases"]) == (tot[0], tot[1], bases)
