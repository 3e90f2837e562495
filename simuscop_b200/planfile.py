"""Reader for SSCPLAN1 plan files (DESIGN.md "Plan file").

A plan file is the flat description of everything the hot path consumes for one output
sample: the FP64 profile CDFs, the haplotype strings, and the bins with their read counts.
It is written both by the instrumented reference (oracle/ref_shim/hooks.cpp) and by our
host front end (SIMUSCOP_DUMP_PLAN), so plans can be compared byte for byte and fed to
the oracle and to the CUDA library alike.
"""
import struct

import numpy as np

from . import abi


class Plan:
    """Flattened plan: numpy arrays laid out as include/simuscop.h expects."""

    def __init__(self):
        self.hdr = {}
        self.tables = {}
        self.genome = None        # uint8 ASCII haplotype store (contig layout)
        self.bins = None          # structured array abi.BIN_DTYPE
        self.segs = None          # structured array abi.SEG_DTYPE
        self.names = b""
        self.units = []           # (popu, chr)
        self.seg_meta = []        # dicts: unit, segIndx, CN, start, end, segsize, readCount

    @property
    def paired(self):
        return bool(self.hdr["paired"])

    def planned_pairs(self):
        rc = self.bins["read_count"].astype(np.int64)
        rc = np.maximum(rc, 0)
        return int(((rc + 1) // 2).sum() if self.paired else rc.sum())

    def profile_struct(self):
        """ctypes ProfileTables pointing into this plan's arrays (keep the Plan alive)."""
        h, t = self.hdr, self.tables
        p = abi.ProfileTables()
        for k in ("n_bases", "kmer", "bins", "n_qual", "min_qual", "read_length", "paired", "use_cdf2",
                  "fixed_insert_size", "min_insert_size", "n_isize", "n_ins", "n_del", "n_kmer_rows"):
            setattr(p, k, int(h[k]))
        p.insert_rate = h["insert_rate"]
        p.del_rate = h["del_rate"]
        p.bases = h["bases"]
        for k in ("isize_cdf", "ins_cdf", "del_cdf", "subs_cdf1", "subs_cdf2", "quality_cdf"):
            a = t.get(k)
            setattr(p, k, a.ctypes.data if a is not None and a.size else None)
        return p


_HDR_KEYS = ["n_bases", "kmer", "bins", "n_qual", "min_qual", "read_length", "paired", "use_cdf2",
             "fixed_insert_size", "min_insert_size", "n_isize", "n_ins", "n_del", "n_kmer_rows", "ploidy", "_pad"]


def _parse_profile(buf, plan):
    ints = struct.unpack_from("<16i", buf, 0)
    h = dict(zip(_HDR_KEYS, ints))
    h["insert_rate"], h["del_rate"] = struct.unpack_from("<2d", buf, 64)
    h["bases"] = bytes(buf[80:88]).rstrip(b"\0")
    off = 88
    N, B, Q, R = h["n_bases"], h["bins"], h["n_qual"], h["n_kmer_rows"]

    def take(n):
        nonlocal off
        a = np.frombuffer(buf, dtype="<f8", count=n, offset=off).copy()
        off += 8 * n
        return a
    t = {}
    t["isize_cdf"] = take(h["n_isize"])
    t["ins_cdf"] = take(h["n_ins"])
    t["del_cdf"] = take(h["n_del"])
    t["subs_cdf1"] = take(R * B * N)
    t["subs_cdf2"] = take(R * B * N) if h["use_cdf2"] else None
    t["quality_cdf"] = take(N * N * B * Q)
    assert off == len(buf), (off, len(buf))
    plan.hdr, plan.tables = h, t


def read_plan(path):
    with open(path, "rb") as f:
        data = f.read()
    assert data[:8] == b"SSCPLAN1", "not a plan file"
    mv = memoryview(data)
    off = 8
    plan = Plan()
    units = []        # per unit: list of segment dicts
    cur = None
    while off < len(data):
        tag, n = struct.unpack_from("<iq", data, off)
        off += 12
        body = mv[off:off + n]
        off += n
        if tag == 1:
            _parse_profile(body, plan)
        elif tag == 2:
            pl, cl = struct.unpack_from("<2i", body, 0)
            popu = bytes(body[8:8 + pl]).decode()
            chrom = bytes(body[8 + pl:8 + pl + cl]).decode()
            cur = {"popu": popu, "chr": chrom, "segs": []}
            units.append(cur)
        elif tag == 3:
            segIndx, CN, start, end, segsize, readCount, nb, ploidy = struct.unpack_from("<2i4q2i", body, 0)
            o = 48
            hapLen = np.frombuffer(body, dtype="<i8", count=ploidy, offset=o).copy()
            o += 8 * ploidy
            haps = []
            for h in range(ploidy):
                L = int(hapLen[h])
                haps.append(np.frombuffer(body, dtype=np.uint8, count=L, offset=o) if L else None)
                o += L
            spos = np.frombuffer(body, dtype="<i8", count=nb, offset=o).copy(); o += 8 * nb
            epos = np.frombuffer(body, dtype="<i8", count=nb, offset=o).copy(); o += 8 * nb
            hap = np.frombuffer(body, dtype="<i4", count=nb, offset=o).copy(); o += 4 * nb
            rc = np.frombuffer(body, dtype="<i4", count=nb, offset=o).copy(); o += 4 * nb
            assert o == n
            cur["segs"].append(dict(segIndx=segIndx, CN=CN, start=start, end=end, segsize=segsize,
                                    readCount=readCount, hapLen=hapLen, haps=haps, spos=spos, epos=epos,
                                    hap=hap, rc=rc))
        elif tag == 5:        # flat dump ("<path>.flat"): the ssc_bin / ssc_segment / name arrays themselves, no haplotypes
            plan.bins = np.frombuffer(body, dtype=abi.BIN_DTYPE).copy()
        elif tag == 6:
            plan.segs = np.frombuffer(body, dtype=abi.SEG_DTYPE).copy()
        elif tag == 7:
            plan.names = bytes(body)
        elif tag == 9:
            break
        else:
            raise ValueError("unknown plan record tag %d" % tag)
    if plan.bins is None:
        _flatten(plan, units)
    return plan


def _flatten(plan, units):
    ploidy = plan.hdr["ploidy"]
    pieces = []
    total = 0
    names = bytearray()
    bins_l, segs_l = [], []
    nbins = 0
    for ui, u in enumerate(units):
        plan.units.append((u["popu"], u["chr"]))
        name = ("@%s#%s#" % (u["popu"], u["chr"])).encode()
        name_off = len(names)
        names += name
        segs = u["segs"]
        # contig layout: for each haplotype index, the segments' strings in order
        hap_base = {}
        contig_end = {}
        for h in range(ploidy):
            for si, s in enumerate(segs):
                if s["haps"][h] is not None:
                    hap_base[(si, h)] = total
                    pieces.append(s["haps"][h])
                    total += len(s["haps"][h])
            contig_end[h] = total
        for si, s in enumerate(segs):
            nb = len(s["spos"])
            seg_id = len(segs_l)
            segs_l.append((nbins, nb, name_off, len(name)))
            plan.seg_meta.append(dict(unit=ui, segIndx=s["segIndx"], CN=s["CN"], start=s["start"], end=s["end"],
                                      segsize=s["segsize"], readCount=s["readCount"]))
            for i in range(nb):
                h = int(s["hap"][i])
                rc = int(s["rc"][i])
                hb = hap_base.get((si, h), 0)
                if (si, h) not in hap_base:
                    rc = 0
                bins_l.append((hb, contig_end.get(h, 0), int(s["spos"][i]), int(s["epos"][i]),
                               int(s["segsize"]) & 0xFFFFFFFF, rc, seg_id, 0))
            nbins += nb
    plan.genome = np.concatenate(pieces) if pieces else np.zeros(0, np.uint8)
    plan.bins = np.array(bins_l, dtype=abi.BIN_DTYPE) if bins_l else np.zeros(0, dtype=abi.BIN_DTYPE)
    plan.segs = np.array(segs_l, dtype=abi.SEG_DTYPE) if segs_l else np.zeros(0, dtype=abi.SEG_DTYPE)
    plan.names = bytes(names)
