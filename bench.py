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
    """Sum of read lengths in FASTQ files (line 2 of every 4-line record): streaming numpy scan, constant memory."""
    total = 0
    for p in paths:
        line_no = 0          # index of the line that the next line feed ends
        prev = -1            # file offset of the previous line feed
        off = 0
        with open(p, "rb") as f:
            while True:
                buf = f.read(1 << 27)
                if not buf:
                    break
                nl = np.flatnonzero(np.frombuffer(buf, dtype=np.uint8) == 10).astype(np.int64) + off
                if len(nl):
                    lens = np.diff(np.concatenate(([prev], nl))) - 1
                    idx = (line_no + np.arange(len(nl))) & 3
                    total += int(lens[idx == 1].sum())
                    line_no += len(nl)
                    prev = int(nl[-1])
                off += len(buf)
    return total


WORKLOAD = "synthetic 3 Gb human-sized genome (24 chromosomes), 30x PE151 WGS, Illumina_HiSeqXTen.profile, insertSize 300, diploid, no SNP/variation"


def run_reference_once(workdir, threads, genome_mb, coverage, tag, profile="XTen"):
    """One run of the unmodified reference binary on a one-chromosome slice; returns (bases, wall seconds)."""
    from simuscop_b200 import paths, synth, testdata
    data = testdata.materialize(os.path.join(workdir, "data"))
    fa = os.path.join(workdir, "cpu_%d.fa" % genome_mb)
    if not os.path.exists(fa + ".ok"):
        synth.make_genome(fa, [genome_mb * 1000000], seed=21, names=["chr1"])
        open(fa + ".ok", "w").write("ok")
    out = os.path.join(workdir, "cpu_out_" + tag)
    shutil.rmtree(out, ignore_errors=True)
    cfg = os.path.join(workdir, "cpu_%s.txt" % tag)
    synth.write_config(cfg, ref=fa, profile=os.path.join(data, testdata.PROFILES[profile]), name="test", output=out,
                       layout="PE", threads=threads, verbose=0, coverage=coverage, insertSize=300)
    t0 = time.perf_counter()
    r = subprocess.run([paths.REF_PLAIN, cfg], capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError("reference simuReads failed: " + r.stderr[-500:])
    bases = count_bases([os.path.join(out, "test_1.fq"), os.path.join(out, "test_2.fq")])
    shutil.rmtree(out, ignore_errors=True)
    return bases, dt


def reference_marginal(workdir, threads, genome_mb, cov_a, cov_b, tag):
    """Throughput of the reference's read-generation path alone: two runs of the unmodified binary on the same slice at
    coverages cov_a < cov_b; everything the program does once per run (FASTA load, segmentation, haplotype strings, the
    single-threaded GC-weight pass, Genome.cpp:783-852) is the same in both and cancels in the difference, which is the
    time Segment::yieldReads needs for (cov_b - cov_a) x slice more bases on `threads` threads."""
    ba, ta = run_reference_once(workdir, threads, genome_mb, cov_a, tag + "a")
    bb, tb = run_reference_once(workdir, threads, genome_mb, cov_b, tag + "b")
    return dict(bases=bb - ba, seconds=tb - ta, run_a=dict(coverage=cov_a, bases=ba, wall_s=ta),
                run_b=dict(coverage=cov_b, bases=bb, wall_s=tb))


def impl_reference(a):
    """The reference arm: the UNMODIFIED reference binary (oracle/_ref/simuReads_ref, built from /root/reference by
    oracle/build_ref.py) on the box's host cores, threads = nproc.  BASELINE.md section 3: the 3 Gb x 30x job costs over an hour
    of CPU, so a slice is timed: ONE chromosome of >= 300 Mb and >= 16 Mb per thread (the reference's task grain is one 1 Mb
    segment with a barrier per chromosome), same profile / layout / insert size.  A step = one 1x pass over the slice
    (= slice-size bases, about 0.35 % of the job).  W warm-up steps + K timed steps are two process runs: run A at coverage W,
    run B at coverage W + K; the K timed steps are the difference B - A (see reference_marginal)."""
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
    genome_mb = a.ref_slice_mb or max(320, min(16 * threads, 1000))
    cov_a = max(1, a.warmup)
    m = reference_marginal(workdir, threads, genome_mb, cov_a, cov_a + a.steps, "ref")
    v = m["bases"] / m["seconds"]
    sample = ("unmodified reference simuReads (oracle/_ref/simuReads_ref), %d threads, one %d Mb synthetic chromosome, PE151 XTen, "
              "insertSize 300; step = one 1x pass over the slice; the %d timed steps = run at %dx (%.1f s wall, %.0f Mbases) minus "
              "run at %dx (%.1f s wall, %.0f Mbases): the per-run setup of the program cancels, what remains is its read-generation "
              "loop incl. FASTQ formatting and file write to %s; whole-process rate of the longer run: %.1f Mbases/s"
              % (threads, genome_mb, a.steps, m["run_b"]["coverage"], m["run_b"]["wall_s"], m["run_b"]["bases"] / 1e6,
                 m["run_a"]["coverage"], m["run_a"]["wall_s"], m["run_a"]["bases"] / 1e6, workdir,
                 m["run_b"]["bases"] / m["run_b"]["wall_s"] / 1e6))
    line = {"impl": "reference", "metric": "simulated_bases_per_sec", "value": v, "unit": "bases/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000.0 * m["seconds"] / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD},
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
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work of the cpu_baseline sample (per thread-second)")
    ap.add_argument("--no-gzip", action="store_true", help="skip the gzip end-to-end leg")
    ap.add_argument("--no-file", action="store_true", help="skip the end-to-end legs that write the FASTQ files")
    ap.add_argument("--no-affinity", action="store_true", help="do not bind the rank to the CPUs of its GPU's NUMA node")
    ap.add_argument("--ref-slice-mb", type=int, default=0, help="--impl reference: size of the one-chromosome slice (default max(320, 16 x threads))")
    ap.add_argument("--file-dirs", default="/dev/shm,workdir", help="directories the e2e_file legs write to (workdir = --workdir)")
    ap.add_argument("--writer-threads", type=int, default=2, help="threads of the file writer (its default mode uses one stream per file)")
    ap.add_argument("--opts", default="", help="experiments: comma-separated ssc_set_option settings, e.g. carry_pass2=0,prefetch_windows=0")
    ap.add_argument("--device-only", action="store_true", help="experiments: only the device-resident leg")
    ap.add_argument("--cli-wall", action="store_true", help="also time the drop-in CLI on the whole job (plain FASTQ to /dev/shm when it fits, gzip to --workdir)")
    a = ap.parse_args()
    if a.impl == "reference":
        return impl_reference(a)

    import ctypes as C
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
    from simuscop_b200 import abi, cuda_binding, host_binding, sharding
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
    for kv in filter(None, a.opts.split(",")):
        gen.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    t0 = time.perf_counter()
    planned, emitted = job.prepare(0, gen)
    t_plan = time.perf_counter() - t0
    lo, hi = sharding.shard_range(planned, rank, world)
    nbatch = (hi - lo) // a.batch_pairs
    if nbatch < 4:
        raise SystemExit("shard of %d pairs is too small for batches of %d pairs" % (hi - lo, a.batch_pairs))

    def spans(first, count):
        """pair ranges covering `count` consecutive batches of this rank's shard starting at batch `first`, wrapping to the
        start of the shard at its end (so usually one range; a batch is revisited only after every other batch of the shard,
        i.e. after tens of GB of other traffic -- nothing of it is left in the 126 MB L2)"""
        out = []
        first %= nbatch
        while count > 0:
            n = min(count, nbatch - first)
            out.append((lo + first * a.batch_pairs, lo + (first + n) * a.batch_pairs))
            count -= n
            first = (first + n) % nbatch
        return out

    # ---------------- device-resident leg: ONE ssc_generate_device call over the K timed batches, outputs stay in HBM
    sampler = ClockSampler(local) if rank == 0 else None
    for rg in spans(0, a.warmup):
        gen.generate_device(*rg)
    gen.reset_stats()
    barrier()
    t0 = time.perf_counter()
    bases = 0
    for rg in spans(a.warmup, a.steps):
        bases += gen.generate_device(*rg)["bases"]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    st = gen.stats()
    dev_ms = st["device_ms"]
    steps_done = a.steps

    if a.device_only:
        if rank == 0:
            nbt = max(1, st["timed_batches"])
            print(json.dumps({"opts": a.opts, "value": bases / dt, "ms_per_step": 1000.0 * dt / a.steps, "gen_ms": st["gen_kernel_ms"] / nbt,
                              "pass2_ms": st["compact_kernel_ms"] / nbt, "profile": a.profile, "host_plan_and_upload_s": round(t_plan, 3)}))
        gen.close(); job.close()
        return 0

    # ---------------- the issue ceiling of this GPU (csrc/floor.cu): Philox only, and Philox + the fast per-base path
    floor = None
    if rank == 0:
        rl = job.read_length
        f0 = min(gen.issue_floor(0, rl, a.batch_pairs, 3) for _ in range(2))
        f1 = min(gen.issue_floor(1, rl, a.batch_pairs, 3) for _ in range(2))
        floor = (2.0 * rl * a.batch_pairs / (f0 / 1000.0), 2.0 * rl * a.batch_pairs / (f1 / 1000.0), f0, f1)
    barrier()

    # ---------------- end-to-end legs: the reference-facing C-ABI call with host buffers (pinned D2H inside)
    state = {"bytes": 0}

    def count_sink(user, b1, l1, b2, l2, first, n):
        state["bytes"] += l1 + l2
        return 0
    base_k = a.warmup + a.steps

    def timed_e2e(sink=None, user=None, warm=None):
        """warm-up call, then one timed ssc_generate() call over K consecutive batches; returns (seconds, stats)"""
        for rg in spans(base_k, min(a.warmup, 3) if warm is None else warm):
            gen.generate(*rg, sink=sink, user=user)
        gen.reset_stats()
        barrier()
        t1 = time.perf_counter()
        for rg in spans(base_k + min(a.warmup, 3), a.steps):
            gen.generate(*rg, sink=sink, user=user)
        torch.cuda.synchronize()
        dte = time.perf_counter() - t1
        barrier()
        return dte, gen.stats()

    dt_e2e, st2 = timed_e2e(sink=count_sink)

    # ---------------- end to end INCLUDING the host write: the sink is the ordered file writer of the drop-in CLI
    # (libsimuscop_host: every slab pwrite()n by a small thread pool at its final offset), both FASTQ files
    file_legs = []
    # (one GPU only: the host write is a property of the box, not of the GPU count -- 4 to 7 GB/s on this pool -- and N ranks
    # writing N x 28 GB at once would fill the box's tmpfs / RAM)
    if not a.no_file and world == 1:
        hl = host_binding.lib()
        hl.ssh_writer_open.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
        hl.ssh_writer_sink.restype = C.c_void_p
        hl.ssh_writer_close.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        sink_ptr = C.cast(hl.ssh_writer_sink(), abi.SINK_FN)
        need = 2.3 * 2 * job.read_length * a.batch_pairs * (min(a.warmup, 3) + a.steps) * 1.05
        for d in a.file_dirs.split(","):
            d = a.workdir if d == "workdir" else d
            leg = {"dir": d}
            try:
                fs = os.statvfs(d)
                free = fs.f_bavail * fs.f_frsize
                if free < need:
                    raise RuntimeError("only %.1f GB free, %.1f GB needed" % (free / 1e9, need / 1e9))
                p1 = os.path.join(d, "simuscop_bench_r%d_1.fq" % rank)
                p2 = os.path.join(d, "simuscop_bench_r%d_2.fq" % rank)
                w = C.c_void_p()
                if hl.ssh_writer_open(p1.encode(), p2.encode(), a.writer_threads, C.byref(w)):
                    raise RuntimeError("cannot create %s" % p1)
                dtf, stf = timed_e2e(sink=sink_ptr, user=w, warm=0)
                t_close = time.perf_counter()
                b1, b2 = C.c_uint64(), C.c_uint64()
                rcw = hl.ssh_writer_close(w, C.byref(b1), C.byref(b2))
                dtf += time.perf_counter() - t_close
                size_ok = os.path.getsize(p1) == b1.value and os.path.getsize(p2) == b2.value
                os.remove(p1); os.remove(p2)
                if rcw or not size_ok:
                    raise RuntimeError("writer failed (rc %d, sizes ok %s)" % (rcw, size_ok))
                leg.update(seconds=dtf, bases=float(stf["bases_emitted"]), bytes=float(stf["d2h_bytes"]), fs=_fs_type(d))
            except Exception as ex:
                leg["skipped"] = str(ex)
            file_legs.append(leg)

    # ---------------- the same end-to-end call with the FASTQ compressed on the GPU (gzip members)
    gz = None
    if not a.no_gzip:
        gen.set_option("gzip", 1)
        dt_gz, st3 = timed_e2e(sink=count_sink)
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
    file_out = []
    for leg in file_legs:
        ok = allsum(0.0 if "skipped" in leg else 1.0) == world
        if ok:
            tf = allmax(leg["seconds"]); bf = allsum(leg["bases"]); byf = allsum(leg["bytes"])
            file_out.append({"dir": leg["dir"], "fs": leg["fs"], "value": bf / tf, "unit": "bases/s", "write_GBps": byf / tf / 1e9,
                             "writer_threads": a.writer_threads})
        else:
            file_out.append({"dir": leg["dir"], "skipped": leg.get("skipped", "skipped on another rank")})
    tot_bases = allsum(float(bases))
    tot_bases_e2e = allsum(float(st2["bases_emitted"]))
    launches = int(allsum(float(st["launches"])))

    if rank == 0:
        peak, peak_src = peaks()
        # roofline of the dominant kernel (generate_slots_kernel): algorithmic bytes per launch / its CUDA-event time
        alg_bytes = st["fastq_bytes"] + st["hap_bytes"] + st["bin_bytes"]
        nb = max(1, st["timed_batches"])
        gen_ms = st["gen_kernel_ms"] / nb
        achieved = alg_bytes / nb / (gen_ms / 1000.0) / 1e9 if gen_ms > 0 else 0.0
        kernel_bases = bases / nb / (gen_ms / 1000.0) if gen_ms > 0 else 0.0
        traffic, ncu = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["generate_slots_kernel"]
            ncu = {k: tr.get(k) for k in ("capture", "commit", "issue_slot_utilisation", "warp_instructions_per_pair")}
            if tr["batch_pairs"] == a.batch_pairs and tr.get("profile", "XTen") == a.profile:
                traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
        except Exception:
            pass
        e2e_file = None
        for leg in file_out:
            if "value" in leg:
                e2e_file = leg          # the first directory that worked is the headline of this leg (tmpfs by default)
                break
        line = {
            "metric": "simulated_bases_per_sec", "value": tot_bases / T, "unit": "bases/s", "n_gpus": world,
            "steps": steps_done, "warmup": a.warmup, "ms_per_step": 1000.0 * T / steps_done, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": WORKLOAD if (a.profile == "XTen" and a.genome_bases == 3000000000 and a.coverage == 30) else
                       "synthetic %.2f Gb genome (24 chromosomes), %dx PE%d WGS, %s profile, insertSize 300, diploid, no SNP/variation"
                       % (genome_len / 1e9, a.coverage, job.read_length, a.profile),
                       "genome_bases": genome_len, "planned_pairs": planned, "batch_pairs": a.batch_pairs, "seed": 1,
                       "cpu_affinity_cpus": affinity,
                       "l2": "inputs larger than L2: every step reads fresh fragments of a %.1f GB packed haplotype store and writes "
                             "a fresh %.1f GB slab" % (2 * genome_len * 0.375 / 1e9, st["fastq_bytes"] / steps_done / 1e9),
                       "setup_s": {"fasta": round(t_fasta, 2), "host_plan_and_upload": round(t_plan, 2)},
                       "device_event_bases_per_sec": tot_bases / T_dev if T_dev > 0 else None},
            "e2e": {"value": tot_bases_e2e / T_e2e, "unit": "bases/s", "h2d_bytes_per_step": 16,
                    "d2h_bytes_per_step": int(st2["d2h_bytes"] / steps_done),
                    "note": "one ssc_generate() call over the K batches: pair range in, FASTQ slabs out through pinned host buffers "
                            "(kernel of batch k+1 overlaps the two device->host copies of batch k); the haplotype store and plan were "
                            "uploaded once from host memory during setup (setup_s); the sink only counts bytes -- e2e_file is the "
                            "same call with the files written"},
            "e2e_file": None if e2e_file is None else dict(
                e2e_file, note="the same call with the drop-in CLI's file writer as the sink: both FASTQ files written (one pwrite "
                               "stream per file, the two files side by side: buffered writes to one file serialise on its inode "
                               "lock, tools/fs_probe.c; no fsync: the reference's SeqWriter does not sync either), files closed "
                               "inside the timed region"),
            "e2e_file_all": file_out,
            "e2e_gzip": None if gz is None else {
                "value": gz_bases / T_gz, "unit": "bases/s", "d2h_bytes_per_step": int(gz_d2h / steps_done),
                "compression_ratio": gz_raw / gz_d2h if gz_d2h else None,
                "note": "optional output mode, not the headline: same call with ssc_set_option(gzip): every 32-record blob is "
                        "deflated on the GPU into one gzip member (literal-only dynamic Huffman, CRC-32 on the fly) before the "
                        "device->host copy; the host receives a valid .fq.gz stream of the same FASTQ bytes"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "achieved_step": alg_bytes / nb / (T_dev / steps_done) / 1e9 if T_dev > 0 else None,
                         "frac_step": alg_bytes / nb / (T_dev / steps_done) / 1e9 / peak if T_dev > 0 else None,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "generate_slots_kernel",
                         "kernel_ms_per_launch": gen_ms, "pass2_ms_per_launch": st["compact_kernel_ms"] / nb,
                         "algorithmic_bytes_per_launch": alg_bytes / nb,
                         "issue": {"unit": "bases/s", "achieved": kernel_bases,
                                   "peak": floor[0], "frac": kernel_bases / floor[0],
                                   "peak_with_fast_path": floor[1], "frac_of_fast_path": kernel_bases / floor[1],
                                   "floor_ms_per_launch": {"philox_only": floor[2], "philox_plus_fast_path": floor[3]},
                                   "note": "measured in this run on this GPU (ssc_issue_floor, csrc/floor.cu), same launch shape and "
                                           "pairs per launch: peak = Philox4x32-10 alone, one block per base (the reference's four "
                                           "uniform draws per base); peak_with_fast_path adds the per-base table work and byte "
                                           "stores of an indel-free read"},
                         "ncu": ncu,
                         "note": "issue-bound kernel, not HBM-bound (DESIGN.md section 4).  kernel_ms_per_launch is the CUDA-event time "
                                 "of generate_slots_kernel (pass 1; on the SMs the concurrent mover leaves it: 140 of 148), "
                                 "pass2_ms_per_launch that of the scan of the blob lengths (+ the stand-alone move of the call's last "
                                 "batch, spread over the batches); the moves of the other batches run under the next generation "
                                 "launch on 8 SMs.  achieved_step / frac_step: the same bytes over the device time of a whole step"},
        }
        if world == 1 and not a.no_cpu_baseline:
            os.sched_setaffinity(0, all_cpus)          # the reference gets every host core again
            threads = os.cpu_count() or 1
            try:
                # bounded sample (about 20 s): the marginal rate of the reference's generation loop on a 64 Mb chromosome
                m = reference_marginal(a.workdir, threads, 64, 1, 1 + max(2, int(round(a.cpu_seconds * threads * 4.0e6 / 64e6 / 2))), "cpu")
                line["cpu_baseline"] = {"value": m["bases"] / m["seconds"], "unit": "bases/s", "cores": threads, "kind": "reference",
                                        "sample": "unmodified reference simuReads (oracle/_ref/simuReads_ref), %d threads, 64 Mb synthetic "
                                                  "chromosome, PE151 XTen: run at %dx (%.1f s wall) minus run at %dx (%.1f s wall) = %.0f Mbases "
                                                  "of read generation in %.1f s (per-run setup cancels); bench.py --impl reference does "
                                                  "the same on a >= 300 Mb slice"
                                                  % (threads, m["run_b"]["coverage"], m["run_b"]["wall_s"], m["run_a"]["coverage"],
                                                     m["run_a"]["wall_s"], m["bases"] / 1e6, m["seconds"])}
            except Exception as ex:  # the baseline is a reported figure; never fail the bench on it
                line["cpu_baseline"] = {"value": None, "unit": "bases/s", "cores": threads, "kind": "reference", "sample": "failed: %s" % ex}
        if a.cli_wall and world == 1:
            line["cli_wall_s"] = cli_wall(a, cfg, job.read_length, planned)
        print(json.dumps(line))
    gen.close()
    job.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def _fs_type(path):
    best, kind = "", "?"
    try:
        with open("/proc/mounts") as f:
            for line in f:
                p = line.split()
                if len(p) >= 3 and os.path.realpath(path).startswith(p[1]) and len(p[1]) >= len(best):
                    best, kind = p[1], p[2]
    except OSError:
        pass
    return kind


def cli_wall(a, cfg_path, read_length, planned):
    """Wall clock of the drop-in CLI (`simuscop_b200/simuReads <config>`) on the whole job: config in, FASTQ files out."""
    from simuscop_b200 import paths
    out = {}
    text = open(cfg_path).read()
    est = 2.3 * 2 * read_length * planned
    for tag, d, env in (("plain_tmpfs", "/dev/shm/simuscop_cli", {}), ("gzip_disk", os.path.join(a.workdir, "cli_gz"), {"SIMUSCOP_GZIP": "1"})):
        try:
            os.makedirs(d, exist_ok=True)
            fs = os.statvfs(d)
            need = est if not env else est / 2.3
            if fs.f_bavail * fs.f_frsize < 1.1 * need:
                raise RuntimeError("%.0f GB free in %s, %.0f GB needed" % (fs.f_bavail * fs.f_frsize / 1e9, d, need / 1e9))
            cfg = os.path.join(a.workdir, "cli_%s.txt" % tag)
            with open(cfg, "w") as f:
                f.write("\n".join(("output = " + d) if l.startswith("output") else l for l in text.splitlines()) + "\n")
            t0 = time.perf_counter()
            r = subprocess.run([paths.SIMUREADS, cfg], env=dict(os.environ, SIMUSCOP_SEED="1", SIMUSCOP_BATCH_PAIRS=str(a.batch_pairs),
                                                               SIMUSCOP_TIMING="1", SIMUSCOP_WRITER_THREADS=str(a.writer_threads), **env),
                               capture_output=True, text=True)
            dt = time.perf_counter() - t0
            if r.returncode != 0:
                raise RuntimeError(r.stderr[-300:])
            size = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d))
            out[tag] = {"wall_s": dt, "output_bytes": size, "timing": [l for l in r.stderr.splitlines() if "[simuscop timing]" in l][-6:]}
        except Exception as ex:
            out[tag] = {"skipped": str(ex)}
        finally:
            shutil.rmtree(d, ignore_errors=True)
    return out


if __name__ == "__main__":
    sys.exit(main())
