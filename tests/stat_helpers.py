"""Statistics of a FASTQ run against its reference genome, for the statistical-parity tests: everything
Profile::predict samples from a table is counted in that table's own coordinates, so two runs (ours / the unmodified
reference, independent random streams) can be compared cell by cell with chi-square homogeneity tests.

Collected per run (numpy, vectorised over the reads):
  sub[mate][row][bin][call]     substitution counts per (k-mer context row, position bin) -- Profile::getSubBaseIndx1/2,
                                lib/profile/Profile.cpp:1527-1554; rows in Profile::initKmers order (:70-124); reads without
                                indels only (their template is known exactly from the read name / the mate's k-mer anchor)
  qual[ref*4+call][bin][q]      quality symbol counts per (reference base, called base, position bin) -- getBaseQuality, :1576-1580
  ins_len / del_len             net length change of reads longer / shorter than RL (insertion / deletion length
                                distributions, :1519-1525) and the number of reads of every class (insertion / deletion rates)
  isize                         fragment lengths (yieldInsertSize, :1486-1493)
  starts                        read-1 start positions (coverage: Segment::yieldReads draws them uniformly inside GC-weighted bins)
"""
import numpy as np
from scipy import stats

COMP = bytes.maketrans(b"ACGTN", b"TGCAN")


def read_fastq(path):
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    return lines[0::4], lines[1::4], lines[3::4]


def _codes(arr, bases):
    """ASCII -> index in `bases` (the profile's base order), 4 for anything else"""
    lut = np.full(256, 4, np.uint8)
    for i, c in enumerate(bases):
        lut[c] = i
    return lut[arr]


def kmer_rows(ctx_codes):
    """ctx_codes: (n, RL, 3) codes of (j-2, j-1, j) with -1 for the 'X' pad in front of the read; returns the row of
    Profile::initKmers' enumeration (XXb -> b, Xab -> 4+4a+b, abc -> 20+16a+4b+c) or -1 when a base is not ACGT."""
    p2, p1, c = ctx_codes[..., 0].astype(np.int64), ctx_codes[..., 1].astype(np.int64), ctx_codes[..., 2].astype(np.int64)
    row = np.where(p2 >= 0, 20 + 16 * p2 + 4 * p1 + c, np.where(p1 >= 0, 4 + 4 * p1 + c, c))
    bad = (c > 3) | ((p1 >= 0) & (p1 > 3)) | ((p2 >= 0) & (p2 > 3))
    return np.where(bad, -1, row)


def collect(fq1, fq2, genome, RL, B, bases=b"ACTG", chrom_len=None, anchor=24):
    """genome: uint8 upper-case ASCII of the (single) chromosome.  Returns a dict of count arrays (module docstring)."""
    n1, s1, q1 = read_fastq(fq1)
    n2, s2, q2 = read_fastq(fq2)
    assert len(s1) == len(s2) > 1000
    npairs = len(s1)
    L = len(genome) if chrom_len is None else chrom_len
    pos = np.array([int(x.split(b"#")[2]) for x in n1], np.int64)
    len1 = np.array([len(x) for x in s1], np.int64)
    len2 = np.array([len(x) for x in s2], np.int64)
    out = {"pairs": npairs, "starts": pos}
    for tag, ln in (("1", len1), ("2", len2)):
        d = ln - RL
        out["ins_len" + tag] = np.bincount(d[d > 0], minlength=64)[:64]
        out["del_len" + tag] = np.bincount(-d[d < 0], minlength=64)[:64]
        out["classes" + tag] = np.array([(d == 0).sum(), (d > 0).sum(), (d < 0).sum()], np.int64)
    gcodes = _codes(genome, bases)
    gb = genome.tobytes()
    bins = (np.arange(RL) * B) // RL
    sub = np.zeros((2, 84, B, 4), np.int64)
    qual = np.zeros((16, B, 94), np.int64)

    def tally(mate, tmpl, reads, quals):
        # tmpl: (n, RL) codes of the template in read orientation; reads / quals: (n, RL) ASCII
        n = tmpl.shape[0]
        if n == 0:
            return
        ctx = np.full((n, RL, 3), -1, np.int64)
        ctx[:, :, 2] = tmpl
        ctx[:, 1:, 1] = tmpl[:, :-1]
        ctx[:, 2:, 0] = tmpl[:, :-2]
        row = kmer_rows(ctx)
        call = _codes(reads, bases).astype(np.int64)
        ok = (row >= 0) & (call < 4)
        b = np.broadcast_to(bins, (n, RL))
        np.add.at(sub[mate], (row[ok], b[ok], call[ok]), 1)
        okq = (tmpl < 4) & (call < 4)
        qq = quals.astype(np.int64) - 33
        np.add.at(qual, ((tmpl.astype(np.int64) * 4 + call)[okq], b[okq], qq[okq]), 1)

    # read 1 without indels: template = genome[pos : pos + RL]
    m1 = (len1 == RL) & (pos + RL <= L)
    idx = np.flatnonzero(m1)
    if len(idx):
        t = gcodes[pos[idx, None] + np.arange(RL)[None, :]]
        r = np.frombuffer(b"".join(s1[i] for i in idx), np.uint8).reshape(-1, RL)
        q = np.frombuffer(b"".join(q1[i] for i in idx), np.uint8).reshape(-1, RL)
        tally(0, t, r, q)
    # read 2 without indels: the fragment end is found from the first `anchor` cycles of the read (low error rate) by exact
    # k-mer lookup on the forward strand; template in read orientation = reverse complement of genome[end-RL : end]
    kmer = {}
    for i in range(0, len(gb) - anchor + 1):
        kmer.setdefault(gb[i:i + anchor], i)
    comp = np.array([bases.index(COMP[c]) if c in b"ACGT" else 4 for c in bases] + [4], np.uint8)
    isz, t2, r2, qq2 = [], [], [], []
    for i in np.flatnonzero(len2 == RL):
        fw = s2[i].translate(COMP)[::-1]                    # read 2 back on the forward strand: last RL bases of the fragment
        hit = kmer.get(fw[-anchor:])
        if hit is None:
            continue
        end = hit + anchor
        if end - RL < 0 or end - pos[i] < RL or end - pos[i] > 5000:
            continue
        isz.append(end - pos[i])
        t2.append(end); r2.append(s2[i]); qq2.append(q2[i])
    out["isize"] = np.array(isz, np.int64)
    if t2:
        ends = np.array(t2, np.int64)
        t = comp[gcodes[ends[:, None] - 1 - np.arange(RL)[None, :]]]
        r = np.frombuffer(b"".join(r2), np.uint8).reshape(-1, RL)
        q = np.frombuffer(b"".join(qq2), np.uint8).reshape(-1, RL)
        tally(1, t, r, q)
    out["sub"], out["qual"] = sub, qual
    return out


def chi2_homogeneity(a, b, min_count=10):
    """Two-sample chi-square homogeneity test of two count vectors over the same cells; cells whose pooled count is below
    min_count are merged into one.  Returns (p, cells)."""
    a, b = np.asarray(a, float).ravel(), np.asarray(b, float).ravel()
    keep = (a + b) >= min_count
    a2 = np.append(a[keep], a[~keep].sum())
    b2 = np.append(b[keep], b[~keep].sum())
    m = (a2 + b2) > 0
    if m.sum() < 2 or a2.sum() == 0 or b2.sum() == 0:
        return 1.0, int(m.sum())
    return float(stats.chi2_contingency(np.vstack([a2[m], b2[m]]))[1]), int(m.sum())


def compare(a, b, alpha=1e-3, report=None):
    """Every table-level statistic of run a against run b; the family-wise level alpha is split over the tests
    (Bonferroni), as SURVEY.md 8c states (p > 0.001 after Bonferroni over the tests).  Returns the list of
    (name, p, threshold) that fail."""
    tests = []
    B = a["sub"].shape[2]
    # substitutions: per (mate, position bin) the full (context row x call) matrix
    for mate in range(2):
        for bn in range(B):
            tests.append(("substitution mate %d bin %d (84 context rows x 4 calls)" % (mate + 1, bn),
                          chi2_homogeneity(a["sub"][mate, :, bn, :], b["sub"][mate, :, bn, :])[0]))
    # qualities: per (ref, call) pair over (position bin x symbol); off-diagonal pairs (substituted bases) pooled over bins
    for rc in range(16):
        if rc // 4 == rc % 4:
            for bn in range(B):
                tests.append(("quality ref=call=%d bin %d" % (rc // 4, bn), chi2_homogeneity(a["qual"][rc, bn], b["qual"][rc, bn])[0]))
        else:
            tests.append(("quality ref %d call %d (all bins)" % (rc // 4, rc % 4),
                          chi2_homogeneity(a["qual"][rc].sum(0), b["qual"][rc].sum(0))[0]))
    for tag in ("1", "2"):
        tests.append(("read classes (no length change / longer / shorter), mate " + tag, chi2_homogeneity(a["classes" + tag], b["classes" + tag], 1)[0]))
        tests.append(("insertion length histogram, mate " + tag, chi2_homogeneity(a["ins_len" + tag], b["ins_len" + tag])[0]))
        tests.append(("deletion length histogram, mate " + tag, chi2_homogeneity(a["del_len" + tag], b["del_len" + tag])[0]))
    if len(a["isize"]) > 100 and len(b["isize"]) > 100:
        lo = int(min(a["isize"].min(), b["isize"].min())); hi = int(max(a["isize"].max(), b["isize"].max()))
        tests.append(("insert size histogram", chi2_homogeneity(np.bincount(a["isize"] - lo, minlength=hi - lo + 1),
                                                              np.bincount(b["isize"] - lo, minlength=hi - lo + 1))[0]))
        tests.append(("insert size KS", float(stats.ks_2samp(a["isize"], b["isize"])[1])))
    thr = alpha / len(tests)
    if report is not None:
        report.extend(tests)
    return [(n, p, thr) for n, p in tests if not (p > thr)]
