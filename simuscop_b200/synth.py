"""Seeded synthetic genomes (the reference's testData/ref.fa.gz is a missing blob).

i.i.d. uniform ACGT with optional N runs and lower-case stretches, written as 60-column
FASTA exactly as fastahack expects.
"""
import numpy as np

_LUT = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_bases(rng, n, n_runs=0, lower_runs=0, run_len=500):
    seq = _LUT[rng.integers(0, 4, size=n, dtype=np.uint8)]
    for _ in range(n_runs):
        s = int(rng.integers(0, max(1, n - run_len)))
        seq[s:s + run_len] = ord("N")
    for _ in range(lower_runs):
        s = int(rng.integers(0, max(1, n - run_len)))
        seq[s:s + run_len] |= 0x20
    return seq


def write_fasta(path, chroms, line=60):
    """chroms: list of (name, uint8 array)."""
    with open(path, "wb") as f:
        for name, seq in chroms:
            f.write(b">" + name.encode() + b"\n")
            n = len(seq)
            full = n // line
            if full:
                body = np.empty((full, line + 1), dtype=np.uint8)
                body[:, :line] = seq[:full * line].reshape(full, line)
                body[:, line] = 10
                f.write(body.tobytes())
            if n % line:
                f.write(seq[full * line:].tobytes() + b"\n")


def make_genome(path, lengths, seed=20, names=None, n_runs=0, lower_runs=0, run_len=500):
    rng = np.random.default_rng(seed)
    chroms = []
    for i, n in enumerate(lengths):
        name = names[i] if names else "chr%d" % (i + 1)
        chroms.append((name, random_bases(rng, n, n_runs, lower_runs, run_len)))
    write_fasta(path, chroms)
    return chroms


def write_config(path, **kw):
    """Write a config file in the reference's grammar (lib/config/Config.cpp:46-99)."""
    order = ["ref", "profile", "variation", "snp", "target", "abundance", "name", "output", "layout",
             "threads", "verbose", "coverage", "insertSize", "ploidy"]
    with open(path, "w") as f:
        for k in order:
            if k in kw and kw[k] is not None:
                f.write("%s = %s\n" % (k, kw[k]))
        for k in kw:
            if k not in order and kw[k] is not None:
                f.write("%s = %s\n" % (k, kw[k]))
