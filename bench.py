#!/usr/bin/env python3
"""bench.py -- simulated bases/s of the simuReads read-generation hot path on N B200s.

Workload (BASELINE.json configs[3]): synthetic 3 Gb human-sized genome (24 chromosomes), 30x PE
WGS, Illumina_HiSeqXTen.profile (RL 151), insert size 300, no SNP/variation.  The C++ front end
builds the haplotype store and the GC-weighted read plan once (setup, untimed); a *step* is one batch of
`--batch-pairs` pairs of that job.  With N GPUs the pair-ID range of the whole job is split into N
contiguous shards (no collective); every rank generates K steps of its own shard ("weak": per-GPU
batch fixed).  `value` = bases all ranks emitted in the K timed steps / max-over-ranks time, outputs left
in HBM; `e2e` = the same through ssc_generate() with host buffers (pinned device->host copy of every
FASTQ byte inside the timed region).

  python bench.py --gpus N --steps K --warmup W            our arm
  python bench.py --impl reference ...                     the reference's own CPU simuReads
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

HUMAN_LENGTHS = [249250621, 243199373, 198022430, 191154276, 180915260, 171115067, 159138663, 146364022,
                 141213431, 135534747, 135006516, 133851895, 115169878, 107349540, 102531392, 90354753,
                 81195210, 78077248, 59128983, 63025520, 48129895, 51304566, 155270560, 59373566]


def scaled_lengths(total):
    s = sum(HUMAN_LENGTHS)
    return [max(200000, int(round(l * total / s))) for l in HUMAN_LENGTHS]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def write_job(workdir, genome_bases, coverage, profile_name="XTen", insert=300, threads=1, seed=20):
    """FASTA (cached per box), config in the reference's grammar; returns config path."""
    from simuscop_b200 import synth, testdata
    data = testdata.materialize(os.path.join(workdir, "data"))
    lengths = scaled_lengths(genome_bases)
    fa = os.path.join(workdir, "genome_%d.fa" % genome_bases)
    if not os.path.exists(fa) or not os.path.exists(fa + ".ok"):
        names = ["chr%d" % (i + 1) for i in range(22)] + ["chrX", "chrY"]
        synth.make_genome(fa, lengths, seed=seed, names=names)
        open(fa + ".ok", "w").write("ok")
    cfg = os.path.join(workdir, "bench_%d_%d_%s.txt" % (genome_bases, coverage, profile_name))
    synth.write_config(cfg, ref=fa, profile=os.path.join(data, testdata.PROFILES[profile_name]), name="test",
                       output=os.path.join(workdir, "out"), layout="PE", threads=threads, verbose=0, coverage=coverage,
                       insertSize=insert)
    return cfg, sum(lengths)


def count_bases(paths):
    """Sum of read lengths in FASTQ files (line 2 of every 4-line record), chunked numpy scan."""
    total = 0
    for p in paths:
        pos = []
        off = 0
        with open(p, "rb") as f:
            while True:
                buf = f.read(1 << 28)
                if not buf:
                    break
                a = np.frombuffer(buf, dtype=np.uint8)
                pos.append(np.flatnonzero(a == 10).astype(np.int64) + off)
                off += len(buf)
        nl = np.concatenate(pos) if pos else np.zeros(0, np.int64)
        starts = np.concatenate(([-1], nl[:-1]))
        lens = nl - starts - 1
        total += int(lens[1::4].sum())
    return total


def run_reference_sample(workdir, threads, genome_mb, coverage, tag):
    """One bounded sample of the workload on the unmodified reference binary; returns (bases, seconds)."""
    from simuscop_b200 import paths, synth, testdata
    data = testdata.materialize(os.path.join(workdir, "data"))
    fa = os.path.join(workdir, "cpu_%d.fa" % genome_mb)
    if not os.path.exists(fa + ".ok"):
        # chromosomes of >= 2 segments per thread keep the reference's 1 Mb task queue busy (Genome.cpp:876-886)
        synth.make_genome(fa, [genome_mb * 1000000], seed=21, names=["chr1"])
        open(fa + ".ok", "w").write("ok")
    out = os.path.join(workdir, "cpu_out_" + tag)
    shutil.rmtree(out, ignore_errors=True)
    cfg = os.path.join(workdir, "cpu_%s.txt" % tag)
    synth.write_config(cfg, ref=fa, profile=os.path.join(data, testdata.PROFILES["XTen"]), name="test", output=out,
                       layout="PE", threads=threads, verbose=0, coverage=coverage, insertSize=300)
    binp = paths.REF_PLAIN
    t0 = time.perf_counter()
    r = subprocess.run([binp, cfg], capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError("reference simuReads failed: " + r.stderr[-500:])
    bases = count_bases([os.path.join(out, "test_1.fq"), os.path.join(out, "test_2.fq")])
    shutil.rmtree(out, ignore_errors=True)
    return bases, dt


def cpu_sample_shape(threads, target_s):
    genome_mb = max(16, 2 * threads)
    bases = min(target_s * threads * 2.5e6, 1.2e9)
    coverage = max(1, int(round(bases / (genome_mb * 1e6))))
    return genome_mb, coverage


def impl_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from simuscop_b200 import paths
    threads = os.cpu_count() or 1
    workdir = a.workdir
    os.makedirs(workdir, exist_ok=True)
    if not os.path.exists(paths.REF_PLAIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/simuReads_ref not built"}))
        return 0
    genome_mb, coverage = cpu_sample_shape(threads, 4.0)
    vals, times = [], []
    for i in range(a.warmup + a.steps):
        bases, dt = run_reference_sample(workdir, threads, genome_mb, coverage, "ref%d" % i)
        if i >= a.warmup:
            vals.append(bases); times.append(dt)
    v = sum(vals) / sum(times)
    sample = "unmodified reference simuReads (oracle/_ref/simuReads_ref), %d threads, whole-process wall time per step on a " \
             "%d Mb synthetic chromosome at %dx PE151 XTen, FASTQ written to %s" % (threads, genome_mb, coverage, workdir)
    line = {"impl": "reference", "metric": "simulated_bases_per_sec", "value": v, "unit": "bases/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000.0 * sum(times) / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "3 Gb synthetic genome, 30x PE151 WGS, HiSeqXTen profile (bounded sample per step)"},
            "cpu_baseline": {"value": v, "unit": "bases/s", "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--genome-bases", type=int, default=3000000000)
    ap.add_argument("--coverage", type=int, default=30)
    ap.add_argument("--profile", default="XTen", choices=["GAIIx", "HiSeq2000", "HiSeq2500", "XTen"],
                    help="sequencing profile of data/ (BASELINE.json configs[4]: profile sweep)")
    ap.add_argument("--batch-pairs", type=int, default=1 << 21)
    ap.add_argument("--workdir", default=os.environ.get("SIMUSCOP_BENCH_DIR", "/tmp/simuscop_bench"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gzip", action="store_true", help="skip the gzip end-to-end leg")
    ap.add_argument("--no-affinity", action="store_true", help="do not bind the rank to the CPUs of its GPU's NUMA node")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    a = ap.parse_args()
    if a.impl == "reference":
        return impl_reference(a)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    # pin this rank to the CPUs next to its GPU before any pinned host buffer is allocated (first touch decides the NUMA
    # node of the FASTQ slabs; with 8 ranks the device->host stream is otherwise limited by cross-socket traffic)
    affinity = None
    all_cpus = os.sched_getaffinity(0)
    if not a.no_affinity:
        try:
            import pynvml
            pynvml.nvmlInit()
            pr = torch.cuda.get_device_properties(local)
            try:
                hnd = pynvml.nvmlDeviceGetHandleByPciBusId(("%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)).encode())
            except Exception:
                hnd = pynvml.nvmlDeviceGetHandleByIndex(local)
            pynvml.nvmlDeviceSetCpuAffinity(hnd)
            affinity = len(os.sched_getaffinity(0))
        except Exception as ex:  # informational only
            affinity = "unavailable: %s" % type(ex).__name__
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    import __graft_entry__ as ge
    from simuscop_b200 import cuda_binding, host_binding, sharding
    if rank == 0:
        ge.build()
    barrier()
    os.makedirs(a.workdir, exist_ok=True)
    t_setup0 = time.perf_counter()
    job = None
    if rank == 0:
        cfg, genome_len = write_job(a.workdir, a.genome_bases, a.coverage, a.profile)
        t_fasta = time.perf_counter() - t_setup0
        job = host_binding.Job(cfg, 1)          # also writes the .fai next to the FASTA, once
    barrier()
    if rank != 0:
        cfg, genome_len = write_job(a.workdir, a.genome_bases, a.coverage, a.profile)
        t_fasta = time.perf_counter() - t_setup0
        job = host_binding.Job(cfg, 1)

    gen = cuda_binding.Generator(local)
    gen.set_option("batch_pairs", a.batch_pairs)
    t0 = time.perf_counter()
    planned, emitted = job.prepare(0, gen)
    t_plan = time.perf_counter() - t0
    lo, hi = sharding.shard_range(planned, rank, world)
    nbatch = (hi - lo) // a.batch_pairs
    if nbatch < 4:
        raise SystemExit("shard of %d pairs is too small for batches of %d pairs" % (hi - lo, a.batch_pairs))

    def step_range(k):
        # consecutive batches of this rank's shard, wrapping around (a batch is revisited only after >= 3 others,
        # i.e. after several GB of other traffic: nothing of it is left in the 126 MB L2)
        s = lo + (k % nbatch) * a.batch_pairs
        return s, s + a.batch_pairs

    # ---------------- device-resident leg: outputs stay in HBM
    sampler = ClockSampler(local) if rank == 0 else None
    for k in range(a.warmup):
        gen.generate_device(*step_range(k))
    gen.reset_stats()
    barrier()
    t0 = time.perf_counter()
    bases = 0
    for k in range(a.warmup, a.warmup + a.steps):
        r = gen.generate_device(*step_range(k))
        bases += r["bases"]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    st = gen.stats()
    dev_ms = st["device_ms"]

    # ---------------- end-to-end leg: the reference-facing C-ABI call with host buffers (pinned D2H inside)
    state = {"bytes": 0}

    def sink(user, b1, l1, b2, l2, first, n):
        state["bytes"] += l1 + l2
        return 0
    # one ssc_generate() call over K consecutive batches -- the call a user makes covers the whole job, and only then does
    # the library's own pipeline (kernel of batch k+1 under the device->host copy of batch k) come into play
    base_k = a.warmup + a.steps

    def run_e2e(first, count):
        count = min(count, nbatch)
        first %= nbatch
        if first + count > nbatch:
            first = 0
        gen.generate(lo + first * a.batch_pairs, lo + (first + count) * a.batch_pairs, sink=sink)
    run_e2e(base_k, min(a.warmup, 3))
    gen.reset_stats()
    state["bytes"] = 0
    barrier()
    t1 = time.perf_counter()
    run_e2e(base_k + min(a.warmup, 3), a.steps)
    torch.cuda.synchronize()
    dt_e2e = time.perf_counter() - t1
    barrier()
    st2 = gen.stats()
    # ---------------- the same end-to-end call with the FASTQ compressed on the GPU (gzip members)
    gz = None
    if not a.no_gzip:
        gen.set_option("gzip", 1)
        run_e2e(base_k, min(a.warmup, 3))
        gen.reset_stats()
        barrier()
        t2 = time.perf_counter()
        run_e2e(base_k + min(a.warmup, 3), a.steps)
        torch.cuda.synchronize()
        dt_gz = time.perf_counter() - t2
        barrier()
        st3 = gen.stats()
        gen.set_option("gzip", 0)
        gz = (dt_gz, st3)
    clocks = sampler.stop() if sampler else None

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    T = allmax(dt)
    T_dev = allmax(dev_ms / 1000.0)
    T_e2e = allmax(dt_e2e)
    if gz is not None:
        T_gz = allmax(gz[0])
        gz_bases = allsum(float(gz[1]["bases_emitted"]))
        gz_d2h = allsum(float(gz[1]["d2h_bytes"]))
        gz_raw = allsum(float(gz[1]["fastq_bytes"]))
    tot_bases = allsum(float(bases))
    tot_bases_e2e = allsum(float(st2["bases_emitted"]))
    launches = int(allsum(float(st["launches"])))

    if rank == 0:
        peak, peak_src = peaks()
        # roofline of the dominant kernel (generate_slots_kernel): algorithmic bytes per launch / its CUDA-event time
        alg_bytes = st["fastq_bytes"] + st["hap_bytes"] + 64.0 * st["pairs_emitted"] / 50.0
        nb = max(1, st["timed_batches"])
        gen_ms = st["gen_kernel_ms"] / nb
        achieved = alg_bytes / nb / (gen_ms / 1000.0) / 1e9 if gen_ms > 0 else 0.0
        traffic, issue, ipp = None, None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["generate_slots_kernel"]
            issue, ipp = tr.get("issue_slot_utilisation"), tr.get("warp_instructions_per_pair")
            if tr["batch_pairs"] == a.batch_pairs:
                traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
        except Exception:
            pass
        line = {
            "metric": "simulated_bases_per_sec", "value": tot_bases / T, "unit": "bases/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000.0 * T / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "synthetic %.2f Gb genome (24 chromosomes), %dx PE%d WGS, %s profile, insertSize 300, "
                                   "diploid, no SNP/variation" % (genome_len / 1e9, a.coverage, job.read_length, a.profile),
                       "planned_pairs": planned, "batch_pairs": a.batch_pairs, "seed": 1, "cpu_affinity_cpus": affinity,
                       "l2": "inputs larger than L2: every step reads fresh fragments of a %.1f GB packed haplotype store and writes "
                             "a fresh %.1f GB slab" % (2 * genome_len * 0.375 / 1e9, st["fastq_bytes"] / a.steps / 1e9),
                       "setup_s": {"fasta": round(t_fasta, 2), "host_plan_and_upload": round(t_plan, 2)},
                       "device_event_bases_per_sec": tot_bases / T_dev if T_dev > 0 else None},
            "e2e": {"value": tot_bases_e2e / T_e2e, "unit": "bases/s", "h2d_bytes_per_step": 16,
                    "d2h_bytes_per_step": int(st2["d2h_bytes"] / a.steps),
                    "note": "one ssc_generate() call over the K batches: pair range in, FASTQ slabs out through pinned host buffers "
                            "(kernel of batch k+1 overlaps the two device->host copies of batch k); the haplotype store and plan were "
                            "uploaded once from host memory during setup (setup_s)"},
            "e2e_gzip": None if gz is None else {
                "value": gz_bases / T_gz, "unit": "bases/s", "d2h_bytes_per_step": int(gz_d2h / a.steps),
                "compression_ratio": gz_raw / gz_d2h if gz_d2h else None,
                "note": "optional output mode, not the headline: same call with ssc_set_option(gzip): every 32-record blob is "
                        "deflated on the GPU into one gzip member (literal-only dynamic Huffman, CRC-32 on the fly) before the "
                        "device->host copy; the host receives a valid .fq.gz stream of the same FASTQ bytes"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "generate_slots_kernel",
                         "kernel_ms_per_launch": gen_ms, "pass2_ms_per_launch": st["compact_kernel_ms"] / nb,
                         "algorithmic_bytes_per_launch": alg_bytes / nb,
                         "issue_slot_utilisation_ncu": issue, "warp_instructions_per_pair_ncu": ipp,
                         "note": "issue-bound kernel, not HBM-bound: one Philox4x32-10 block per base (4 draws) plus ~25 table / compare "
                                 "instructions; ncu figures from profiles/ (traffic = dram bytes of one launch of this kernel); pass 2 "
                                 "(scan + move of the per-ticket blobs to the dense ordered slab) adds ~2x the FASTQ bytes of HBM traffic "
                                 "per step at ~80 % of the copy bandwidth; see DESIGN.md"},
        }
        if world == 1 and not a.no_cpu_baseline:
            os.sched_setaffinity(0, all_cpus)          # the reference gets every host core again
            threads = os.cpu_count() or 1
            from simuscop_b200 import paths
            try:
                genome_mb, coverage = cpu_sample_shape(threads, a.cpu_seconds)
                cb, cdt = run_reference_sample(a.workdir, threads, genome_mb, coverage, "cpu")
                line["cpu_baseline"] = {"value": cb / cdt, "unit": "bases/s", "cores": threads, "kind": "reference",
                                        "sample": "unmodified reference simuReads (oracle/_ref/simuReads_ref), %d threads, %d Mb "
                                                  "synthetic chromosome at %dx PE151 XTen, whole-process wall %.1f s, %.0f Mbases"
                                                  % (threads, genome_mb, coverage, cdt, cb / 1e6)}
            except Exception as ex:  # the baseline is a reported figure; never fail the bench on it
                line["cpu_baseline"] = {"value": None, "unit": "bases/s", "cores": threads, "kind": "reference", "sample": "failed: %s" % ex}
        print(json.dumps(line))
    gen.close()
    job.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
