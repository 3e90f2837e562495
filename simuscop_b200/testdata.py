"""Materialise the gzip-compressed input fixtures of data/ into a scratch directory."""
import gzip
import os
import shutil

from .paths import DATA

PROFILES = {
    "GAIIx": "Illumina_GenomeAnalyzerIIx.profile",
    "HiSeq2000": "Illumina_HiSeq2000.profile",
    "HiSeq2500": "Illumina_HiSeq2500.profile",
    "XTen": "Illumina_HiSeqXTen.profile",
}


def materialize(dst, names=None):
    """gunzip data/<name>.gz -> dst/<name>; returns dst."""
    os.makedirs(dst, exist_ok=True)
    for fn in sorted(os.listdir(DATA)):
        if not fn.endswith(".gz"):
            continue
        base = fn[:-3]
        if names is not None and base not in names:
            continue
        out = os.path.join(dst, base)
        if not os.path.exists(out):
            with gzip.open(os.path.join(DATA, fn), "rb") as fi, open(out + ".tmp", "wb") as fo:
                shutil.copyfileobj(fi, fo)
            os.replace(out + ".tmp", out)
    return dst
