"""Host model of the fast kernel's indel splice (simuscop_b200/csrc/gen_fast.cu: scan_events -> splice_read -> emit_packed).

The kernel never materialises the post-indel source sequence of Profile::predict (Profile.cpp:1636-1658).  It keeps the
events as a sorted list of output-coordinate segments and assembles the read 16 bases per lane from slices of the 2-bit
packed window (and of the packed inserted bases) with funnel shifts, in store orientation.  This file restates that
index arithmetic word for word in Python and checks it, on random event lists and both read orientations, against the
direct construction the reference performs (and that oracle/ssc_oracle.c restates).  The byte-for-byte check of the CUDA
code itself is tests/test_gpu_parity.py; this one keeps the arithmetic honest where no GPU is available."""
import random

import pytest

COMP = [2, 3, 0, 1]          # complement in the profile's base order ACTG: A<->T, C<->G


def reference_source(template, events):
    """Profile.cpp:1636-1658: deletion drops bases j..j+k-1, insertion emits base j then the inserted bases."""
    out = []
    ev = {e[1]: e for e in events}
    j = 0
    while j < len(template):
        e = ev.get(j)
        if e and e[0] == "D":
            j += e[2]
            continue
        out.append(template[j])
        if e and e[0] == "I":
            out.extend(e[3])
        j += 1
    return out


def draw_events(rng, RL, max_events):
    """Events the way the scan produces them: ascending positions, candidates inside a deleted stretch skipped,
    deletions clamped to the read end (Profile.cpp:1613), inserted bases never the last code (Profile.cpp:1564)."""
    events, j = [], 0
    for _ in range(rng.randint(0, max_events)):
        j += rng.randint(0, max(1, RL // 3))
        if j >= RL:
            break
        if rng.random() < 0.5:
            n = rng.randint(1, 9)
            events.append(("I", j, n, [rng.randint(0, 2) for _ in range(n)]))
            j += 1
        else:
            n = min(rng.randint(1, 12), RL - j)
            events.append(("D", j, n, None))
            j += n
    return events


def scan_events(events):
    """Event records of the kernel: (first output position behind the event's template base, inserted bases, offset of
    the inserted bases in insb, output position - template position of everything behind the event)."""
    recs, insb, cum = [], [], 0
    for kind, j, n, bases in events:
        if kind == "I":
            recs.append((j + cum + 1, n, len(insb), cum + n))
            insb.extend(bases)
            cum += n
        else:
            recs.append((j + cum, 0, 0, cum - n))
            cum -= n
    return recs, insb, cum


def pack(codes, pad_words=1, total_words=None):
    """2 bits per base, base i in bits 2*(i & 15) of word i >> 4; `pad_words` zero words in front (index -1 is readable)."""
    n = (len(codes) + 15) // 16 + 1 if total_words is None else total_words
    words = [0] * (n + pad_words)
    for i, c in enumerate(codes):
        words[pad_words + (i >> 4)] |= (c & 3) << ((i & 15) * 2)
    return words


def funnel_r(lo, hi, sh):
    return (((hi << 32) | lo) >> sh) & 0xFFFFFFFF


def gather16(src, pad, q, lo, hi):
    wi = q >> 4                                   # arithmetic shift: -1 for q in -16..-1
    f = funnel_r(src[pad + wi], src[pad + wi + 1], (q & 15) * 2)
    return f & (0xFFFFFFFF >> (32 - 2 * hi)) & ((0xFFFFFFFF << (2 * lo)) & 0xFFFFFFFF)


def splice(window, dOff, RL, recs, insb, m, rev):
    """splice_read: the 16 words of the spliced read, store orientation, base y at packed index dOff + y."""
    ins_total = len(insb)
    insp_codes = [COMP[insb[ins_total - 1 - i]] if rev else insb[i] for i in range(ins_total)]
    insp = pack(insp_codes, total_words=9)
    out = []
    for lane in range(16):
        y0 = 16 * lane - dOff
        word, prev_o, prev_shift = 0, 0, 0
        for k in range(len(recs) + 1):
            start, n_ins, ins_off, shift_after = recs[k] if k < len(recs) else (m, 0, 0, 0)
            n, t1 = start - prev_o, prev_o - prev_shift
            dst = m - start if rev else prev_o
            src = RL - (t1 + n) if rev else t1
            lo, hi = max(dst, y0) - y0, min(dst + n, y0 + 16) - y0
            if lo < hi:
                word |= gather16(window, 1, dOff + src + y0 - dst, lo, hi)
            if n_ins > 0:
                dst = m - (start + n_ins) if rev else start
                src = ins_total - (ins_off + n_ins) if rev else ins_off
                lo, hi = max(dst, y0) - y0, min(dst + n_ins, y0 + 16) - y0
                if lo < hi:
                    word |= gather16(insp, 1, src + y0 - dst, lo, hi)
            prev_o, prev_shift = start + n_ins, shift_after
        out.append(word)
    return out


def unpack(words, first, n):
    return [(words[(first + i) >> 4] >> (((first + i) & 15) * 2)) & 3 for i in range(n)]


@pytest.mark.parametrize("RL", [33, 74, 125, 151, 160])
@pytest.mark.parametrize("rev", [False, True])
def test_spliced_read_equals_the_reference_source_sequence(RL, rev):
    rng = random.Random(1000 * RL + rev)
    checked = 0
    for trial in range(400):
        events = draw_events(rng, RL, 4 if trial % 4 else 1)
        recs, insb, cum = scan_events(events)
        m = RL + cum
        if m < 1 or m > 160 or len(insb) > 128:      # longer reads take the position-by-position path (emit_mapped)
            continue
        dOff = 32 + rng.randint(0, 15)                # gen_fast.cu: index of the read's first store base in the window
        # the read in output orientation, and the store around it (the window holds store orientation)
        template = [rng.randint(0, 3) for _ in range(RL)]
        store = [COMP[c] for c in reversed(template)] if rev else list(template)
        flank = [rng.randint(0, 3) for _ in range(dOff)]
        tail = [rng.randint(0, 3) for _ in range(256 - dOff - RL)]
        window = pack(flank + store + tail, total_words=17)
        want = reference_source(template, events)
        assert len(want) == m
        out = splice(window, dOff, RL, recs, insb, m, rev)
        got_store = unpack(out, dOff, m)
        got = [COMP[c] for c in reversed(got_store)] if rev else got_store
        assert got == want, (RL, rev, events)
        # nothing is written in front of the read: the two context bases of cycles 0 and 1 are the LUT's 'X' pads
        assert unpack(out, 0, dOff) == [0] * dOff
        checked += 1
    assert checked > 150


def test_event_segments_map_output_positions_like_the_reference():
    """emit_mapped's per-position rule: the last segment that starts at or before output position o decides."""
    rng = random.Random(7)
    for _ in range(300):
        RL = rng.choice([75, 120, 151])
        events = draw_events(rng, RL, 5)
        recs, insb, cum = scan_events(events)
        template = list(range(100, 100 + RL))        # distinct values: a template position each
        marked = [(k, j, n, [1000 + 50 * i + b for b in range(n)] if k == "I" else None) for i, (k, j, n, _) in enumerate(events)]
        want = reference_source(template, marked)
        recs_m, _, _ = scan_events(marked)
        insm = [v for (k, j, n, b) in marked if k == "I" for v in b]
        got = []
        for o in range(RL + cum):
            shift, ins_idx = 0, -1
            for start, n_ins, ins_off, shift_after in recs_m:
                d = o - start
                if d >= 0:
                    shift = shift_after
                    ins_idx = ins_off + d if d < n_ins else -1
            got.append(insm[ins_idx] if ins_idx >= 0 else template[o - shift])
        assert got == want
