"""Shared test plumbing: seeded scenarios run through the Philox-instrumented reference."""
import glob
import os
import subprocess

from simuscop_b200 import planfile, synth, testdata
from simuscop_b200.paths import REF_PHILOX

# name -> dict(lengths, names, profile, layout, coverage, insertSize, extras)
SCENARIOS = {
    # paired-end XTen, N runs and lower-case stretches
    "pe_xten": dict(lengths=[300000, 120000], profile="XTen", layout="PE", coverage=5, insertSize=300,
                    n_runs=2, lower_runs=2),
    # single-end GAIIx (both strands, fragment = bin length)
    "se_gaiix": dict(lengths=[200000], profile="GAIIx", layout="SE", coverage=3, insertSize=250, n_runs=1),
    # tiny chromosomes: tail bin shorter than RL (abandon after 1000 fails), chromosome shorter than the insert
    "pe_tiny": dict(lengths=[1040, 300, 5200], profile="HiSeq2500", layout="PE", coverage=60, insertSize=200),
    # variations: CNV gain/loss, SNV, insertion, deletion, SNPs
    "pe_variants": dict(lengths=[2600000], names=["chr20"], profile="HiSeq2000", layout="PE", coverage=2,
                        insertSize=250, variation=True, snp=True, n_runs=2),
    # capture targets
    "pe_wes": dict(lengths=[900000], names=["chr20"], profile="HiSeq2500", layout="PE", coverage=20,
                   insertSize=200, target=True),
    # four populations mixed by two abundance rows, single-end
    "se_tumor": dict(lengths=[1500000], names=["chr20"], profile="GAIIx", layout="SE", coverage=2,
                     insertSize=250, tumor=True),
}

# triploid genome with copy-number changes, and a haploid single-end run
SCENARIOS["pe_ploidy3"] = dict(lengths=[2600000], names=["chr20"], profile="XTen", layout="PE", coverage=2, insertSize=300,
                               variation=True, ploidy=3, n_runs=1)
SCENARIOS["se_ploidy1"] = dict(lengths=[700000, 300000], profile="HiSeq2500", layout="SE", coverage=3, insertSize=200, ploidy=1)
# IUPAC ambiguity codes in the FASTA, in an inserted sequence and in an SNV allele: they behave as N on the read path but are
# ordinary non-GC bases for the GC weights (calculateGCPercent counts only a literal 'N', MyDefine.cpp:279-303)
SCENARIOS["pe_iupac"] = dict(lengths=[260000, 90000], names=["chr20", "chr21"], profile="XTen", layout="PE", coverage=3,
                             insertSize=300, n_runs=1, lower_runs=1, iupac=60, variation_text=(
                                 "i\ttest\tchr20\t45010\ttcgrytcg\thomo\n"
                                 "s\ttest\tchr20\t50010\tA\tY\thet\n"
                                 "d\ttest\tchr20\t120010\t8\thomo\n"
                                 "c\ttest\tchr20\t150001\t200000\t3\t2\n"))
SCENARIOS["se_mini"] = dict(lengths=[30000], profile="GAIIx", layout="SE", coverage=2, insertSize=250, n_runs=1)

# synthetic profiles (simuscop_b200.synth.write_profile): odd k-mer sizes, short / long reads, heavy indel rates,
# degenerate table rows, fixed insert size -- they drive the generic kernels and the indel path hard
STRESS = {
    "k2_short": dict(profile_kw=dict(kmer=2, read_length=40, bins=40, ins_rate=2e-3, del_rate=2e-3), layout="PE", insertSize=120),
    "k4_long": dict(profile_kw=dict(kmer=4, read_length=200, bins=25, ins_rate=3e-3, del_rate=3e-3, n_ins=40, n_del=50, n_live_qual=30), layout="PE", insertSize=400),
    "k3_indel_heavy": dict(profile_kw=dict(kmer=3, read_length=120, bins=30, ins_rate=6e-3, del_rate=8e-3, n_ins=12, n_del=30), layout="PE", insertSize=260),
    "k3_fixed_insert_se": dict(profile_kw=dict(kmer=3, read_length=64, bins=16, std_isize=0.0, n_live_qual=40), layout="SE", insertSize=150),
    "k3_fixed_insert_pe": dict(profile_kw=dict(kmer=3, read_length=90, bins=50, std_isize=0.0, n_live_qual=6, bases="GATC"), layout="PE", insertSize=200),
    "k1_tcga": dict(profile_kw=dict(kmer=1, read_length=151, bins=50, bases="TCGA", n_live_qual=7), layout="PE", insertSize=300),
    # fast-kernel edge cases: read lengths around the 32-cycle chunk boundaries (the record terminators ride on the idle
    # lanes of the last chunk only when RL % 32 is in 1..29), the longest / shortest supported reads, <= 7 live quality
    # symbols (all quality rows in shared memory) and a name that makes the record header longer than one / two warps
    "k3_rl95": dict(profile_kw=dict(kmer=3, read_length=95, bins=50, n_live_qual=7), layout="PE", insertSize=220),
    "k3_rl126": dict(profile_kw=dict(kmer=3, read_length=126, bins=33, n_live_qual=5, ins_rate=1e-3), layout="PE", insertSize=280),
    "k3_rl160": dict(profile_kw=dict(kmer=3, read_length=160, bins=50, n_live_qual=7, bases="CATG"), layout="PE", insertSize=330),
    "k3_rl33": dict(profile_kw=dict(kmer=3, read_length=33, bins=11, n_live_qual=7), layout="SE", insertSize=80),
    "k3_long_name": dict(profile_kw=dict(kmer=3, read_length=100, bins=20, n_live_qual=7), layout="PE", insertSize=250,
                         name="population_with_a_rather_long_name_0123456789"),
    "k3_longer_name": dict(profile_kw=dict(kmer=3, read_length=75, bins=50, n_live_qual=12), layout="PE", insertSize=200,
                           name="p" * 55),
}


# beyond the documented per-read limits: must fail loudly, never diverge silently
OVERFLOW = dict(profile_kw=dict(kmer=3, read_length=120, bins=30, ins_rate=2e-2, del_rate=2e-2, n_ins=30, n_del=30), layout="PE", insertSize=260)


def build_stress(name, workdir, seed=5):
    sc = OVERFLOW if name == "overflow" else STRESS[name]
    d = os.path.join(workdir, "stress_" + name)
    os.makedirs(d, exist_ok=True)
    synth.make_genome(os.path.join(d, "ref.fa"), [60000, 9000], seed=31, n_runs=2, lower_runs=1, run_len=200)
    prof = os.path.join(d, "synthetic.profile")
    synth.write_profile(prof, seed=sum(map(ord, name)), **sc["profile_kw"])
    kw = dict(ref=os.path.join(d, "ref.fa"), profile=prof, layout=sc["layout"], coverage=8, insertSize=sc["insertSize"],
              threads=1, verbose=0, name=sc.get("name", "s"))
    return dict(dir=d, kw=kw, seed=seed, name=name)


VARIATION_SMALL = """\
i\ttest\tchr20\t450010\ttcgagtcg\thomo
i\ttest\tchr20\t1100010\ttcgagtc\thomo
i\ttest\tchr20\t1200010\ttcgagt\thet
d\ttest\tchr20\t460010\t8\thomo
d\ttest\tchr20\t1300010\t12\thet
d\ttest\tchr20\t1400010\t5\thet
s\ttest\tchr20\t500010\tA\tC\thomo
s\ttest\tchr20\t600010\tG\tT\thet
s\ttest\tchr20\t700010\tC\tA\thet
c\ttest\tchr20\t800001\t1200000\t3\t2
c\ttest\tchr20\t1600001\t1900000\t1\t1
c\ttest\tchr20\t2000001\t2400000\t4\t2
"""

VARIATION_TUMOR = """\
i\tclone1\tchr20\t450010\ttcgagtcg\thomo
i\tclone2\tchr20\t450010\ttcgagtcg\thet
d\tclone3\tchr20\t460010\t8\thomo
s\tclone1\tchr20\t500010\tA\tC\thomo
s\tclone2\tchr20\t600010\tG\tT\thet
s\tclone4\tchr20\t700010\tC\tA\thet
c\tclone1\tchr20\t200001\t600000\t3\t2
c\tclone2\tchr20\t700001\t900000\t1\t1
c\tclone3\tchr20\t1000001\t1400000\t4\t3
c\tclone4\tchr20\t100001\t300000\t5\t3
"""


def build_scenario(name, workdir, seed=7):
    """Writes genome/config for SCENARIOS[name]; returns dict(cfg=..., dir=..., seed=...)."""
    sc = SCENARIOS[name]
    d = os.path.join(workdir, name)
    os.makedirs(d, exist_ok=True)
    data = testdata.materialize(os.path.join(workdir, "data"))
    synth.make_genome(os.path.join(d, "ref.fa"), sc["lengths"], seed=20, names=sc.get("names"),
                      n_runs=sc.get("n_runs", 0), lower_runs=sc.get("lower_runs", 0), run_len=300, iupac=sc.get("iupac", 0))
    kw = dict(ref=os.path.join(d, "ref.fa"), profile=os.path.join(data, testdata.PROFILES[sc["profile"]]),
              layout=sc["layout"], coverage=sc["coverage"], insertSize=sc["insertSize"], threads=1, verbose=0,
              name="test")
    if sc.get("ploidy"):
        kw["ploidy"] = sc["ploidy"]
    if sc.get("variation_text"):
        with open(os.path.join(d, "variations.txt"), "w") as f:
            f.write(sc["variation_text"])
        kw["variation"] = os.path.join(d, "variations.txt")
    if sc.get("variation"):
        with open(os.path.join(d, "variations.txt"), "w") as f:
            f.write(VARIATION_SMALL)
        kw["variation"] = os.path.join(d, "variations.txt")
    if sc.get("snp"):
        # SNPs of the shipped file that fall inside the synthetic chromosome
        with open(os.path.join(data, "snp.txt")) as f, open(os.path.join(d, "snp.txt"), "w") as g:
            for line in f:
                p = line.split("\t")
                if len(p) >= 3 and int(p[2]) <= sc["lengths"][0]:
                    g.write(line)
        kw["snp"] = os.path.join(d, "snp.txt")
    if sc.get("target"):
        with open(os.path.join(data, "exon_regions.bed")) as f, open(os.path.join(d, "targets.bed"), "w") as g:
            for line in f:
                p = line.split("\t")
                if len(p) >= 3 and int(p[2]) + 60 <= sc["lengths"][0]:
                    g.write(line)
        kw["target"] = os.path.join(d, "targets.bed")
    if sc.get("tumor"):
        with open(os.path.join(d, "variations.txt"), "w") as f:
            f.write(VARIATION_TUMOR)
        with open(os.path.join(d, "abundance.txt"), "w") as f:
            f.write("1.0\t0\t0\t0\n0.3\t0.25\t0.35\t0.1\n")
        kw["variation"] = os.path.join(d, "variations.txt")
        kw["abundance"] = os.path.join(d, "abundance.txt")
        kw["name"] = "clone1,clone2,clone3,clone4"
    return dict(dir=d, kw=kw, seed=seed, name=name)


def build_random_variation_scenario(seed, workdir):
    """A seeded random job for the plan-construction fuzz test: one chromosome, disjoint CNVs (copy numbers 1..5 with a random
    major copy number), insertions / deletions / SNVs (homo- and heterozygous) in shuffled file order, ploidy 1..3, PE or SE."""
    import random
    rng = random.Random(seed)
    L = rng.choice([400000, 900000, 1500000])
    d = os.path.join(workdir, "fuzz%d" % seed)
    os.makedirs(d, exist_ok=True)
    data = testdata.materialize(os.path.join(workdir, "data"))
    synth.make_genome(os.path.join(d, "ref.fa"), [L], seed=seed, names=["chr20"], n_runs=rng.randint(0, 2),
                      lower_runs=rng.randint(0, 2), run_len=300)
    with open(os.path.join(d, "ref.fa")) as f:
        ref = f.read().split("\n", 1)[1].replace("\n", "").upper()
    ploidy = rng.choice([1, 2, 2, 3])
    lines = []
    pos = 50001
    while pos + 150000 < L and rng.random() < 0.8:
        ln = rng.choice([60000, 100000, 250000])
        if pos + ln > L:
            break
        cn = rng.choice([1, 3, 4, 5])
        # a haploid genome cannot split a gain between two haplotype sets: the reference spins forever in
        # Segment.cpp:183-189 unless every copy is "major" (our front end rejects that input, test_host_logic)
        major = cn if ploidy == 1 else rng.randint((cn + 1) // 2, cn)
        lines.append("c\ttest\tchr20\t%d\t%d\t%d\t%d" % (pos, pos + ln - 1, cn, major))
        pos += ln + rng.choice([1, 40000, 120000])
    used = set()
    for _ in range(rng.randint(3, 12)):
        p = rng.randint(2000, L - 2000)
        if any(abs(p - u) < 200 for u in used):
            continue
        used.add(p)
        z = rng.choice(["homo", "het"])
        k = rng.random()
        if k < 0.35:
            lines.append("i\ttest\tchr20\t%d\t%s\t%s" % (p, "".join(rng.choice("acgt") for _ in range(rng.randint(1, 12))), z))
        elif k < 0.7:
            lines.append("d\ttest\tchr20\t%d\t%d\t%s" % (p, rng.randint(1, 15), z))
        elif ref[p - 1] in "ACGT":
            lines.append("s\ttest\tchr20\t%d\t%s\t%s\t%s" % (p, ref[p - 1], rng.choice([b for b in "ACGT" if b != ref[p - 1]]), z))
    rng.shuffle(lines)
    with open(os.path.join(d, "variations.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    kw = dict(ref=os.path.join(d, "ref.fa"), profile=os.path.join(data, testdata.PROFILES[rng.choice(["XTen", "GAIIx", "HiSeq2500"])]),
              layout=rng.choice(["PE", "SE"]), coverage=1, insertSize=rng.choice([200, 300]), threads=1, verbose=0, name="test",
              variation=os.path.join(d, "variations.txt"))
    if ploidy != 2:
        kw["ploidy"] = ploidy
    return dict(dir=d, kw=kw, seed=seed, name="fuzz%d" % seed)


def build_random_job_scenario(seed, workdir):
    """A seeded random job over the other inputs of the front end: 1..3 chromosomes, capture targets (BED intervals from 30 bp to
    2.6 kb, touching / close / far apart), SNP files on both strands, tumour mixtures (2..3 populations x 1..2 abundance rows) with
    per-population variations.  Returns (scenario, mode)."""
    import random
    rng = random.Random(seed)
    nchr = rng.choice([1, 2, 3])
    lens = [rng.choice([120000, 400000, 700000]) for _ in range(nchr)]
    names = ["chr%d" % (20 + i) for i in range(nchr)]
    d = os.path.join(workdir, "job%d" % seed)
    os.makedirs(d, exist_ok=True)
    data = testdata.materialize(os.path.join(workdir, "data"))
    synth.make_genome(os.path.join(d, "ref.fa"), lens, seed=seed, names=names, n_runs=rng.randint(0, 2), lower_runs=rng.randint(0, 1),
                      run_len=300)
    with open(os.path.join(d, "ref.fa")) as f:
        recs = f.read().split(">")[1:]
    refs = {t.split("\n", 1)[0].split()[0]: t.split("\n", 1)[1].replace("\n", "").upper() for t in recs}
    kw = dict(ref=os.path.join(d, "ref.fa"),
              profile=os.path.join(data, testdata.PROFILES[rng.choice(["XTen", "GAIIx", "HiSeq2500", "HiSeq2000"])]),
              layout=rng.choice(["PE", "SE"]), coverage=rng.choice([1, 2]), insertSize=rng.choice([200, 300]), threads=1, verbose=0,
              name="test")
    mode = rng.choice(["wes", "snp", "tumor", "wes+var"])
    pops = ["test"]
    if mode == "tumor":
        pops = ["c1", "c2", "c3"][:rng.choice([2, 3])]
        kw["name"] = ",".join(pops)
        rows = []
        for _ in range(rng.choice([1, 2])):
            w = [rng.randint(1, 9) for _ in pops]
            w = [x * 100 // sum(w) for x in w]
            w[0] += 100 - sum(w)                                   # rows sum to exactly 1 (the front end checks)
            rows.append("\t".join("%.2f" % (x / 100.0) for x in w))
        with open(os.path.join(d, "abundance.txt"), "w") as f:
            f.write("\n".join(rows) + "\n")
        kw["abundance"] = os.path.join(d, "abundance.txt")
    if mode in ("tumor", "wes+var", "snp"):
        lines = []
        for pop in pops:
            for c, L in zip(names, lens):
                pos = 20001
                while pos + 80000 < L and rng.random() < 0.6:
                    ln = rng.choice([30000, 60000])
                    cn = rng.choice([1, 3, 4])
                    lines.append("c\t%s\t%s\t%d\t%d\t%d\t%d" % (pop, c, pos, pos + ln - 1, cn, rng.randint((cn + 1) // 2, cn)))
                    pos += ln + rng.choice([1, 30000])
                used = set()
                for _ in range(rng.randint(0, 6)):
                    p = rng.randint(2000, L - 2000)
                    if any(abs(p - u) < 200 for u in used):
                        continue
                    used.add(p)
                    z = rng.choice(["homo", "het"])
                    k = rng.random()
                    if k < 0.35:
                        lines.append("i\t%s\t%s\t%d\t%s\t%s" % (pop, c, p, "".join(rng.choice("acgt") for _ in range(rng.randint(1, 9))), z))
                    elif k < 0.7:
                        lines.append("d\t%s\t%s\t%d\t%d\t%s" % (pop, c, p, rng.randint(1, 12), z))
                    elif refs[c][p - 1] in "ACGT":
                        lines.append("s\t%s\t%s\t%d\t%s\t%s\t%s" % (pop, c, p, refs[c][p - 1],
                                                                   rng.choice([b for b in "ACGT" if b != refs[c][p - 1]]), z))
        rng.shuffle(lines)
        if lines:
            with open(os.path.join(d, "variations.txt"), "w") as f:
                f.write("\n".join(lines) + "\n")
            kw["variation"] = os.path.join(d, "variations.txt")
    if mode == "snp" or rng.random() < 0.3:
        comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
        sl = []
        for c, L in zip(names, lens):
            for i in range(rng.randint(1, 30)):
                p = rng.randint(100, L - 100)
                r = refs[c][p - 1]
                if r not in "ACGT":
                    continue
                alt = rng.choice([b for b in "ACGT" if b != r])
                strand = rng.choice("+-")
                a, b = (r, alt) if strand == "+" else (comp[r], comp[alt])
                sl.append((c, p, "rs%d\t%s\t%d\t%s\t%s\t%s" % (seed * 1000 + i, c, p, "%s/%s" % ((a, b) if rng.random() < 0.5 else (b, a)),
                                                               strand, a)))
        sl.sort()
        with open(os.path.join(d, "snp.txt"), "w") as f:
            f.write("\n".join(x[2] for x in sl) + "\n")
        kw["snp"] = os.path.join(d, "snp.txt")
    if mode in ("wes", "wes+var"):
        bl = []
        for c, L in zip(names, lens):
            p = rng.randint(100, 3000)
            while p < L - 200:
                e = min(p + rng.choice([60, 150, 400, 1200, 2600]), L - 60)
                if e - p > 30:
                    bl.append("%s\t%d\t%d" % (c, p, e))
                p = e + rng.choice([1, 30, 90, 2000, 20000])
        with open(os.path.join(d, "targets.bed"), "w") as f:
            f.write("\n".join(bl) + "\n")
        kw["target"] = os.path.join(d, "targets.bed")
        kw["coverage"] = 10
    return dict(dir=d, kw=kw, seed=seed, name="job%d" % seed), mode


def build_edge_variation_scenario(seed, workdir):
    """Seeded variation files that sit on the edges of the reference's segment logic: overlapping CNVs (the 1-bp overlapping
    segments of Genome.cpp:663-674), variants at position 1 / the last base / beyond the end, dense indel + SNV clusters,
    deletions and insertions across CNV boundaries, variants inside N runs, chromosomes the FASTA does not have.  Returns
    (scenario, kind)."""
    import random
    rng = random.Random(seed)
    wd = os.path.join(workdir, "edge%d" % seed)
    os.makedirs(wd, exist_ok=True)
    data = testdata.materialize(os.path.join(workdir, "data"))
    L = rng.choice([30000, 90000, 250000])
    synth.make_genome(os.path.join(wd, "ref.fa"), [L, 8000], seed=seed, names=["chr20", "chr21"], n_runs=rng.randint(0, 2), lower_runs=1, run_len=200)
    ref = open(os.path.join(wd, "ref.fa")).read().split(">")[1].split("\n", 1)[1].replace("\n", "").upper()
    lines = []
    kind = rng.choice(["overlap_cnv", "edge_pos", "dense", "beyond_end", "unknown_chr", "n_region", "del_cross_cnv", "mixed"])
    def snv(p, z=None):
        if 1 <= p <= L and ref[p-1] in "ACGT": lines.append("s\ttest\tchr20\t%d\t%s\t%s\t%s" % (p, ref[p-1], rng.choice([b for b in "ACGT" if b != ref[p-1]]), z or rng.choice(["homo", "het"])))
    if kind == "overlap_cnv":
        a = rng.randint(2000, L // 3); b = a + rng.randint(3000, L // 3)
        lines.append("c\ttest\tchr20\t%d\t%d\t3\t2" % (a, b)); lines.append("c\ttest\tchr20\t%d\t%d\t1\t1" % (b, min(L, b + rng.randint(3000, 9000))))
    elif kind == "edge_pos":
        lines.append("i\ttest\tchr20\t1\tacgt\thomo"); lines.append("d\ttest\tchr20\t%d\t3\thet" % (L - 5)); snv(1); snv(L)
        lines.append("i\ttest\tchr20\t%d\tgg\thet" % L)
        lines.append("c\ttest\tchr20\t1\t%d\t3\t2" % rng.randint(2000, L // 2))
    elif kind == "dense":
        p = rng.randint(1000, L - 2000)
        for i in range(rng.randint(3, 8)):
            k = rng.random(); q = p + i * rng.randint(1, 12)
            if k < 0.4: lines.append("i\ttest\tchr20\t%d\t%s\t%s" % (q, "".join(rng.choice("acgt") for _ in range(rng.randint(1, 6))), rng.choice(["homo", "het"])))
            elif k < 0.7: lines.append("d\ttest\tchr20\t%d\t%d\t%s" % (q, rng.randint(1, 9), rng.choice(["homo", "het"])))
            else: snv(q)
    elif kind == "beyond_end":
        lines.append("d\ttest\tchr20\t%d\t50\thomo" % (L - 10)); lines.append("c\ttest\tchr20\t%d\t%d\t3\t2" % (L - 5000, L + 5000)); lines.append("i\ttest\tchr20\t%d\tacg\thomo" % (L + 100))
    elif kind == "unknown_chr":
        lines.append("s\ttest\tchr5\t100\tA\tC\thomo"); lines.append("c\ttest\tchrX\t1000\t5000\t3\t2"); snv(rng.randint(100, L - 100))
    elif kind == "n_region":
        idx = ref.find("N")
        if idx >= 0:
            lines.append("i\ttest\tchr20\t%d\tacgt\thomo" % (idx + 5)); lines.append("d\ttest\tchr20\t%d\t20\thet" % (idx + 50)); lines.append("c\ttest\tchr20\t%d\t%d\t3\t3" % (max(1, idx - 3000), min(L, idx + 3000)))
    elif kind == "del_cross_cnv":
        a = rng.randint(3000, L // 2); b = a + 5000
        lines.append("c\ttest\tchr20\t%d\t%d\t3\t2" % (a, b)); lines.append("d\ttest\tchr20\t%d\t12\thomo" % (a - 5)); lines.append("d\ttest\tchr20\t%d\t12\thet" % (b - 5)); lines.append("i\ttest\tchr20\t%d\taa\thomo" % a); lines.append("i\ttest\tchr20\t%d\tcc\thet" % b)
    else:
        for _ in range(rng.randint(5, 25)):
            p = rng.randint(1, L); k = rng.random()
            if k < 0.3: lines.append("i\ttest\tchr20\t%d\t%s\t%s" % (p, "".join(rng.choice("acgtn") for _ in range(rng.randint(1, 20))), rng.choice(["homo", "het"])))
            elif k < 0.6: lines.append("d\ttest\tchr20\t%d\t%d\t%s" % (p, rng.randint(1, 40), rng.choice(["homo", "het"])))
            elif k < 0.9: snv(p)
            else:
                e = min(L, p + rng.randint(1000, 20000)); lines.append("c\ttest\tchr20\t%d\t%d\t%d\t%d" % (p, e, rng.choice([1,3,4]), 2))
    rng.shuffle(lines)
    open(os.path.join(wd, "variations.txt"), "w").write("\n".join(lines) + "\n")
    kw = dict(ref=os.path.join(wd, "ref.fa"), profile=os.path.join(data, testdata.PROFILES[rng.choice(["XTen", "GAIIx"])]), layout=rng.choice(["PE", "SE"]), coverage=1,
              insertSize=rng.choice([200, 300]), threads=1, verbose=0, name="test", variation=os.path.join(wd, "variations.txt"))
    if rng.random() < 0.3: kw["ploidy"] = 3
    return dict(dir=wd, kw=kw, seed=seed, name="edge%d" % seed), kind


EDGE_INPUT_KINDS = ["bed_overlap", "bed_beyond", "bed_tiny", "bed_unknown_chr", "bed_unsorted", "bed_start_gt_end", "snp_mismatch_ref", "snp_dup", "snp_unknown_chr", "snp_beyond", "snp_bad_strand",
                   "abund_not_one", "abund_cols", "abund_zero", "bed_plus_cnv", "snp_in_cnv_indel"]


def build_edge_input_scenario(seed, workdir):
    """Edge cases of the other input files (kind = EDGE_INPUT_KINDS[seed % 16]): capture targets that overlap, run past the
    chromosome end, are tiny / unsorted / inverted / on unknown chromosomes; SNP lines whose reference allele does not match,
    duplicates, unknown chromosomes, positions beyond the end, an unknown strand symbol; abundance rows that do not sum to one,
    have the wrong number of columns or zero entries; targets and SNPs inside CNVs and indels.  Returns (scenario, kind)."""
    import random
    rng = random.Random(seed)
    kind = EDGE_INPUT_KINDS[seed % len(EDGE_INPUT_KINDS)]
    wd = os.path.join(workdir, "input%d" % seed)
    os.makedirs(wd, exist_ok=True)
    data = testdata.materialize(os.path.join(workdir, "data"))
    L = rng.choice([40000, 120000])
    synth.make_genome(os.path.join(wd, "ref.fa"), [L, 9000], seed=seed, names=["chr20", "chr21"], n_runs=rng.randint(0, 1), lower_runs=1, run_len=200)
    ref = open(os.path.join(wd, "ref.fa")).read().split(">")[1].split("\n", 1)[1].replace("\n", "").upper()
    kw = dict(ref=os.path.join(wd, "ref.fa"), profile=os.path.join(data, testdata.PROFILES[rng.choice(["XTen", "HiSeq2500"])]), layout=rng.choice(["PE", "SE"]), coverage=5,
              insertSize=rng.choice([200, 300]), threads=1, verbose=0, name="test")
    bed, snp, var, ab = [], [], [], None
    def snpline(c, p, strand="+", r=None, alt=None, first=None, i=0):
        r = r or (ref[p-1] if c == "chr20" and 1 <= p <= L else "A"); alt = alt or rng.choice([b for b in "ACGT" if b != r])
        return "rs%d\t%s\t%d\t%s/%s\t%s\t%s" % (seed * 100 + i, c, p, r, alt, strand, first or r)
    if kind == "bed_overlap":
        p = 2000
        for i in range(8): bed.append(("chr20", p, p + rng.randint(200, 1500))); p += rng.randint(50, 600)
    elif kind == "bed_beyond": bed += [("chr20", L - 300, L + 500), ("chr20", L + 1000, L + 2000), ("chr20", 5000, 5600)]
    elif kind == "bed_tiny": bed += [("chr20", 3000, 3001), ("chr20", 4000, 4010), ("chr20", 1, 40), ("chr20", 10, 30), ("chr20", 8000, 8900)]
    elif kind == "bed_unknown_chr": bed += [("chr7", 100, 900), ("chr20", 3000, 3900), ("chr21", 100, 700)]
    elif kind == "bed_unsorted": bed += [("chr20", 9000, 9800), ("chr20", 2000, 2900), ("chr21", 500, 900), ("chr20", 5000, 5200)]
    elif kind == "bed_start_gt_end": bed += [("chr20", 5000, 4000), ("chr20", 7000, 7900)]
    elif kind == "snp_mismatch_ref": snp += [snpline("chr20", p, r=rng.choice("ACGT"), i=i) for i, p in enumerate(sorted(rng.sample(range(100, L - 100), 12)))]
    elif kind == "snp_dup":
        p = rng.randint(100, L - 100); snp += [snpline("chr20", p, i=0), snpline("chr20", p, i=1), snpline("chr20", p + 1, i=2)]
    elif kind == "snp_unknown_chr": snp += [snpline("chr9", 500, i=0), snpline("chr20", 700, i=1)]
    elif kind == "snp_beyond": snp += [snpline("chr20", L + 50, i=0), snpline("chr20", L, i=1), snpline("chr20", 1, i=2)]
    elif kind == "snp_bad_strand": snp += [snpline("chr20", 900, strand="?", i=0), snpline("chr20", 1900, strand="-", i=1)]
    elif kind.startswith("abund"):
        kw["name"] = "c1,c2"
        ab = {"abund_not_one": "0.5\t0.4\n", "abund_cols": "0.5\t0.3\t0.2\n", "abund_zero": "1.0\t0.0\n0.0\t1.0\n"}[kind]
        var += ["c\tc1\tchr20\t5001\t15000\t3\t2", "s\tc2\tchr20\t%d\t%s\t%s\thomo" % (3000, ref[2999] if ref[2999] in "ACGT" else "A", "C" if ref[2999] != "C" else "G")]
    elif kind == "bed_plus_cnv":
        bed += [("chr20", 4000, 5200), ("chr20", 9500, 10400), ("chr20", 14800, 15300)]
        var += ["c\ttest\tchr20\t5001\t10000\t3\t2", "c\ttest\tchr20\t10001\t15000\t1\t1", "d\ttest\tchr20\t4990\t30\thomo", "i\ttest\tchr20\t9999\tacgtacgt\thet"]
    elif kind == "snp_in_cnv_indel":
        var += ["c\ttest\tchr20\t5001\t10000\t4\t3", "d\ttest\tchr20\t7000\t10\thet", "i\ttest\tchr20\t7100\tacg\thomo"]
        snp += [snpline("chr20", p, i=i) for i, p in enumerate([5001, 6999, 7000, 7005, 7010, 7100, 7101, 10000, 10001])]
    if bed:
        open(os.path.join(wd, "t.bed"), "w").write("".join("%s\t%d\t%d\n" % b for b in bed)); kw["target"] = os.path.join(wd, "t.bed")
    if snp:
        open(os.path.join(wd, "snp.txt"), "w").write("\n".join(snp) + "\n"); kw["snp"] = os.path.join(wd, "snp.txt")
    if var:
        open(os.path.join(wd, "v.txt"), "w").write("\n".join(var) + "\n"); kw["variation"] = os.path.join(wd, "v.txt")
    if ab:
        open(os.path.join(wd, "ab.txt"), "w").write(ab); kw["abundance"] = os.path.join(wd, "ab.txt")
    return dict(dir=wd, kw=kw, seed=seed, name="input%d" % seed), kind


def run_reference_philox(scn, tag="ref"):
    """Runs the instrumented reference; returns (list of plan paths, sorted list of fastq paths)."""
    d = scn["dir"]
    out = os.path.join(d, "out_" + tag)
    cfg = os.path.join(d, "cfg_%s.txt" % tag)
    synth.write_config(cfg, output=out, **scn["kw"])
    for f in glob.glob(os.path.join(d, "plan_%s.*.plan" % tag)):
        os.remove(f)
    env = dict(os.environ, SIMUSCOP_SEED=str(scn["seed"]), SIMUSCOP_DUMP_PLAN=os.path.join(d, "plan_" + tag))
    r = subprocess.run([REF_PHILOX, cfg], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    plans = sorted(glob.glob(os.path.join(d, "plan_%s.*.plan" % tag)), key=lambda p: int(p.split(".")[-2]))
    return plans, out


def sample_files(out, plan, index, scn):
    """FASTQ files the reference wrote for plan/sample `index` (Genome.cpp:857-866, 899-929)."""
    if scn["kw"].get("abundance"):
        with open(scn["kw"]["abundance"]) as f:
            rows = [l.split("\t") for l in f.read().splitlines() if l]
        names = scn["kw"]["name"].split(",")
        fn = "+".join("%s_%.3f" % (n, float(p)) for n, p in zip(names, rows[index]))
    else:
        fn = scn["kw"]["name"].split(",")[0]
    if plan.paired:
        return os.path.join(out, fn + "_1.fq"), os.path.join(out, fn + "_2.fq")
    return os.path.join(out, fn + ".fq"), None


def read_file(p):
    if p is None:
        return b""
    with open(p, "rb") as f:
        return f.read()


def first_diff(a, b):
    n = min(len(a), len(b))
    for i in range(n):
        if a[i] != b[i]:
            return i
    return n if len(a) != len(b) else -1


# ---------------------------------------------------------------------------------------------
# The reference's three shipped test configurations (configFiles/config_test_{wgs,wes,tumor}.txt), key for key and value for
# value, run from a directory laid out like the reference's checkout (./testData/..., ./results).  testData/ref.fa.gz is a
# missing blob upstream: a seeded synthetic chr20 of hg19's length (63 025 520 bp, covers every coordinate of
# variations*.txt / snp.txt / exon_regions.bed) with N runs and lower-case stretches stands in for it.
# ---------------------------------------------------------------------------------------------
SHIPPED = {
    "wgs": dict(ref="./testData/ref.fa.gz", profile="./testData/Illumina_GenomeAnalyzerIIx.profile",
                variation="./testData/variations.txt", snp="./testData/snp.txt", name="test", output="./results",
                layout="PE", threads=4, verbose=1, coverage=10, insertSize=250),
    "wes": dict(ref="./testData/ref.fa.gz", profile="./testData/Illumina_HiSeq2500.profile",
                variation="./testData/variations.txt", snp="./testData/snp.txt", target="./testData/exon_regions.bed",
                name="test", output="./results", layout="PE", threads=2, verbose=1, coverage=100, insertSize=200),
    "tumor": dict(ref="./testData/ref.fa.gz", profile="./testData/Illumina_GenomeAnalyzerIIx.profile",
                  variation="./testData/variations_tumor.txt", snp="./testData/snp.txt", name="clone1, clone2, clone3, normal",
                  abundance="./testData/abundance_tumor.txt", output="./results", layout="SE", threads=4, verbose=1,
                  coverage=10, insertSize=200),
}
SHIPPED_SEED = 7
SHIPPED_CHR20 = 63025520


def build_shipped_tree(workdir):
    """<workdir>/shipped/{testData,configFiles}: the data files of data/, the synthetic ref.fa.gz and the three configs."""
    import gzip
    import shutil
    root = os.path.join(workdir, "shipped")
    td = os.path.join(root, "testData")
    os.makedirs(os.path.join(root, "configFiles"), exist_ok=True)
    testdata.materialize(td)
    gz = os.path.join(td, "ref.fa.gz")
    if not os.path.exists(gz):
        fa = os.path.join(td, "ref_synth.fa")
        synth.make_genome(fa, [SHIPPED_CHR20], seed=20, names=["chr20"], n_runs=6, lower_runs=6, run_len=20000)
        with open(fa, "rb") as fi, gzip.GzipFile(gz + ".tmp", "wb", compresslevel=1, mtime=0) as fo:
            shutil.copyfileobj(fi, fo, 1 << 24)
        os.replace(gz + ".tmp", gz)
        os.remove(fa)
    for k, kw in SHIPPED.items():
        synth.write_config(os.path.join(root, "configFiles", "config_test_%s.txt" % k), **kw)
    return root


def run_shipped(binary, root, which, tag, env_extra=None):
    """Runs `binary configFiles/config_test_<which>.txt` from a private copy of the tree (the programs gunzip the reference
    next to itself and write ./results); returns {file name: path} of the FASTQ files."""
    import shutil
    run = os.path.join(root, "run_%s_%s" % (which, tag))
    shutil.rmtree(run, ignore_errors=True)
    os.makedirs(os.path.join(run, "testData"))
    for fn in os.listdir(os.path.join(root, "testData")):
        if fn in ("ref.fa", "ref.fa.fai"):
            continue
        src = os.path.join(root, "testData", fn)
        dst = os.path.join(run, "testData", fn)
        if fn == "ref.fa.gz":
            shutil.copyfile(src, dst)          # gunzipped in place by the program
        else:
            os.symlink(src, dst)
    os.makedirs(os.path.join(run, "configFiles"))
    shutil.copyfile(os.path.join(root, "configFiles", "config_test_%s.txt" % which), os.path.join(run, "configFiles", "config.txt"))
    env = dict(os.environ, SIMUSCOP_SEED=str(SHIPPED_SEED), **(env_extra or {}))
    r = subprocess.run([binary, "configFiles/config.txt"], cwd=run, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    res = os.path.join(run, "results")
    if not os.path.isdir(res):
        return {}
    return {fn: os.path.join(res, fn) for fn in sorted(os.listdir(res))}


class FastaStoreWindow:
    """Windows of the haplotype store of a variant-free job, taken from the FASTA file itself (not from the device): without
    variations every contig (population, chromosome, haplotype index) is the upper-cased chromosome, and the store is the
    contigs back to back in plan order -- contig k covers [contig_end[k-1], contig_end[k]).  Used by the bench-scale parity
    tests, whose 6-Gbase store is too large to hold as ASCII."""

    def __init__(self, plan, fasta_path):
        import numpy as np
        self.np = np
        self.fd = os.open(fasta_path, os.O_RDONLY)          # pread: safe to share with forked oracle workers
        fai = {}
        with open(fasta_path + ".fai") as f:
            for line in f:
                p = line.rstrip("\n").split("\t")
                name = p[0].split()[0]
                for pre in ("chrom", "chr"):
                    if name.startswith(pre):
                        name = name[len(pre):]
                        break
                fai[name] = (int(p[1]), int(p[2]), int(p[3]), int(p[4]))
        ends, chroms = [], []
        segs, bins, names = plan.segs, plan.bins, plan.names
        for s in range(len(segs)):
            nb = int(segs["n_bins"][s])
            if nb == 0:
                continue
            b = bins[int(segs["first_bin"][s]):int(segs["first_bin"][s]) + nb]
            b = b[b["read_count"] > 0]
            nm = names[int(segs["name_offset"][s]):int(segs["name_offset"][s]) + int(segs["name_len"][s])].decode()
            chrom = nm.split("#")[1]
            for ce in np.unique(b["contig_end"]):
                ends.append(int(ce)); chroms.append(chrom)
        order = sorted(set(zip(ends, chroms)))
        self.ends = [e for e, _ in order]
        self.chroms = [c for _, c in order]
        self.starts = [0] + self.ends[:-1]
        for st, en, c in zip(self.starts, self.ends, self.chroms):
            assert en - st == fai[c][0], "contig of %s is not the plain chromosome (%d vs %d bases)" % (c, en - st, fai[c][0])
        self.fai = fai

    def _chrom_slice(self, chrom, a, b):
        np = self.np
        length, off, lb, lw = self.fai[chrom]
        fa0 = off + (a // lb) * lw + a % lb
        fa1 = off + ((b - 1) // lb) * lw + (b - 1) % lb + 1
        raw = np.frombuffer(os.pread(self.fd, fa1 - fa0, fa0), np.uint8)
        seq = raw[raw != 10]
        assert len(seq) == b - a
        return np.where((seq >= 97) & (seq <= 122), seq - 32, seq).astype(np.uint8)

    def window(self, lo, hi):
        """ASCII store bases [lo, hi) (clipped to the store)."""
        import bisect
        np = self.np
        lo = max(0, lo); hi = min(hi, self.ends[-1])
        parts = []
        k = bisect.bisect_right(self.ends, lo)
        while lo < hi:
            st, en = self.starts[k], self.ends[k]
            b = min(hi, en)
            parts.append(self._chrom_slice(self.chroms[k], lo - st, b - st))
            lo = b; k += 1
        return np.concatenate(parts) if parts else np.zeros(0, np.uint8)


def pair_window(plan, pair_lo, pair_hi, slack=4096):
    """Store interval [lo, hi) that planned pairs [pair_lo, pair_hi) can touch: their bins' start ranges plus `slack` bases
    (longest fragment), and the bin index range."""
    import numpy as np
    rc = np.maximum(plan.bins["read_count"].astype(np.int64), 0)
    per = (rc + 1) // 2 if plan.paired else rc
    base = np.concatenate(([0], np.cumsum(per)))
    b0 = int(np.searchsorted(base, pair_lo, side="right") - 1)
    b1 = int(np.searchsorted(base, pair_hi, side="left"))
    b = plan.bins[b0:b1]
    b = b[per[b0:b1] > 0]
    lo = int((b["hap_base"] + b["spos"]).min())
    hi = int(np.minimum(b["hap_base"] + b["epos"] + slack, b["contig_end"]).max())
    return lo, hi, b0, b1
