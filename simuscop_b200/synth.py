"""Seeded synthetic genomes (the reference's testData/ref.fa.gz is a missing blob).

i.i.d. uniform ACGT with optional N runs and lower-case stretches, written as 60-column
FASTA exactly as fastahack expects.
"""
import numpy as np

_LUT = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_bases(rng, n, n_runs=0, lower_runs=0, run_len=500, iupac=0):
    seq = _LUT[rng.integers(0, 4, size=n, dtype=np.uint8)]
    if iupac:
        # IUPAC ambiguity codes as hg19 / hg38 carry them: neither ACGT nor N (drawn first so that the other options
        # leave the random stream of existing fixtures untouched when iupac == 0)
        pos = rng.integers(0, n, size=iupac)
        seq[pos] = np.frombuffer(b"RYKMSWrymk", dtype=np.uint8)[rng.integers(0, 10, size=iupac)]
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


def make_genome(path, lengths, seed=20, names=None, n_runs=0, lower_runs=0, run_len=500, iupac=0):
    rng = np.random.default_rng(seed)
    chroms = []
    for i, n in enumerate(lengths):
        name = names[i] if names else "chr%d" % (i + 1)
        chroms.append((name, random_bases(rng, n, n_runs, lower_runs, run_len, iupac)))
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


def write_profile(path, seed=1, bases="ACTG", kmer=3, read_length=100, bins=20, ins_rate=5e-4, del_rate=5e-4,
                  n_ins=12, n_del=20, std_isize=30.0, n_live_qual=10, sub_rate=0.01, zero_rows=2, gc_std=0.15):
    """Random .profile in the format of Profile::saveResults (lib/profile/Profile.cpp:1240-1365).

    Deliberately nasty: zero-probability leading symbols, a few all-zero substitution rows (identity rule,
    Profile.cpp:848-860), an all-zero quality row (-> '~'), unnormalised rows, arbitrary k-mer size."""
    rng = np.random.default_rng(seed)
    N = len(bases)

    def kmers():
        out = []
        for pad in range(kmer - 1, -1, -1):
            valid = kmer - pad
            for v in range(N ** valid):
                digs = []
                x = v
                for _ in range(valid):
                    digs.append(x % N)
                    x //= N
                out.append("X" * pad + "".join(bases[d] for d in reversed(digs)))
        return out
    with open(path, "w") as f:
        f.write("#model created by simuscop_b200.synth.write_profile seed %d\n#reads: synthetic\n\n" % seed)
        f.write("bases: %s\nreadLength: %d\nbinCount: %d\nkmer: %d\n\n" % (bases, read_length, bins, kmer))
        f.write("\n[Insert Rate]\n%g\n[Insert Frequency]\n" % ins_rate)
        insf = np.concatenate([[0.0], rng.random(n_ins - 1)])
        f.write("\t".join("%g" % x for x in insf / insf.sum()) + "\n")
        f.write("\n[Deletion Rate]\n%g\n[Deletion Frequency]\n" % del_rate)
        delf = np.concatenate([[0.0], rng.random(n_del - 1)])
        f.write("\t".join("%g" % x for x in delf / delf.sum()) + "\n")
        f.write("\n[Substitution Probs]\n")
        names = kmers()
        zero = set(rng.choice(len(names), size=min(zero_rows, len(names)), replace=False).tolist())
        for ki, km in enumerate(names):
            f.write("kmer: %s\n" % km)
            last = bases.index(km[-1])
            for r in range(2 * bins):
                if ki in zero and r % 7 == 0:
                    row = np.zeros(N)
                else:
                    row = rng.random(N) * sub_rate
                    row[last] = 1.0 - row.sum() + row[last]
                    if rng.random() < 0.15:
                        row[int(rng.integers(0, N))] = 0.0       # zero-probability symbol (possibly a leading one)
                    row *= rng.uniform(0.5, 3.0)                  # unnormalised on purpose
                f.write("\t".join("%g" % x for x in row) + "\n")
        f.write("\n[Base Quality Distribution]\n")
        for bp in range(N * N):
            f.write("basePairIndx: %d\n" % bp)
            for r in range(bins):
                row = np.zeros(94)
                if not (bp == 5 and r == 1):                      # one all-zero row
                    lo = int(rng.integers(0, 94 - n_live_qual))
                    idx = rng.choice(np.arange(lo, min(94, lo + 2 * n_live_qual)), size=n_live_qual, replace=False)
                    row[idx] = rng.random(n_live_qual) * rng.uniform(1, 1000)
                f.write("\t".join("%g" % x for x in row) + "\n")
        f.write("\n[Insert Size Standard Deviation]\n%g\n" % std_isize)
        f.write("\n[Log Ratio Mean Value]\n")
        for g in range(101):
            f.write("%d\t%g\n" % (g, 1.0 + 0.3 * np.sin(g / 16.0) if 20 <= g <= 80 else 0.2))
        f.write("\n[Log Ratio Standard Deviation]\n%g\n" % gc_std)
