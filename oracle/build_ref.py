#!/usr/bin/env python3
"""Build the reference binaries used as the parity oracle and the CPU baseline.

TEST INFRASTRUCTURE (oracle side) -- not part of the product.

Outputs (only into oracle/_ref/, which is git-ignored but travels to the GPU box):

  oracle/_ref/simuReads_ref     the UNMODIFIED reference `simuReads`, compiled with g++
                                straight from the sources where they lie under
                                /root/reference (flags of its Release build:
                                -std=c++11 -O3 -DNDEBUG -pthread).  Used as the CPU
                                baseline (bench.py --impl reference) and for the
                                statistical parity tests.
  oracle/_ref/simuReads_philox  the Philox-INSTRUMENTED reference: same sources, but
                                (1) lib/threadpool/ThreadPool.cpp is replaced by
                                    oracle/ref_shim/philox_pool.cpp (synchronous work
                                    items, addressed Philox words instead of mt19937);
                                (2) one-line hook calls are inserted at the draw sites
                                    of Segment.cpp / Profile.cpp and at the three
                                    wall-clock seed sites.
                                The patched copies live in a scratch directory under
                                /tmp for the duration of the build and are deleted;
                                no reference source enters this repository.

The reference's own build system (cmake) is not run.
Usage: python oracle/build_ref.py [--ref /root/reference] [--force]
"""
import argparse
import glob
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SHIM = os.path.join(HERE, "ref_shim")

CXXFLAGS = ["-std=c++11", "-O3", "-DNDEBUG", "-pthread", "-w"]

# (file, regex, replacement, minimum number of substitutions)
PATCHES = [
    # ---- Segment.cpp: Segment::yieldReads (lib/segment/Segment.cpp:741-766)
    ("lib/segment/Segment.cpp", r"^(\t\tint n = seg->getReadCount\(i\);)$", r"\1 ssc_hook_bin(n);", 1),
    ("lib/segment/Segment.cpp", r"^(\t\twhile\(n > 0\) \{)$", r"\1 ssc_hook_attempt();", 1),
    ("lib/segment/Segment.cpp", r"^(\t\t\tfragCount\+\+;)$", r"\1 ssc_hook_frag_ok();", 1),
    ("lib/segment/Segment.cpp", r"^(\t\t\t\t)(k = threadPool->randomInteger\(0, 2\);)$",
     r"\1ssc_hook_slot(SSC_SLOT_STRAND); \2", 1),
    # ---- Profile.cpp: sampler entries (lib/profile/Profile.cpp:1486-1584) and predict loops
    ("lib/profile/Profile.cpp", r"^(int Profile::yieldInsertSize\(\) \{)$", r"\1 ssc_hook_slot(SSC_SLOT_ISIZE);", 1),
    ("lib/profile/Profile.cpp", r"^(int Profile::getInsertLen\(\) \{)$", r"\1 ssc_hook_slot(SSC_SLOT_INSLEN);", 1),
    ("lib/profile/Profile.cpp", r"^(int Profile::getDelLen\(\) \{)$", r"\1 ssc_hook_slot(SSC_SLOT_DELLEN);", 1),
    ("lib/profile/Profile.cpp", r"^(int Profile::getSubBaseIndx1\(char \*kmerSeq, int binIndx\) \{)$",
     r"\1 ssc_hook_slot(SSC_SLOT_SUB);", 1),
    ("lib/profile/Profile.cpp", r"^(int Profile::getSubBaseIndx2\(char \*kmerSeq, int binIndx\) \{)$",
     r"\1 ssc_hook_slot(SSC_SLOT_SUB);", 1),
    ("lib/profile/Profile.cpp", r"^(int Profile::getBaseQuality\(int basePairIndx, int binIndx\) \{)$",
     r"\1 ssc_hook_slot(SSC_SLOT_QUAL);", 1),
    ("lib/profile/Profile.cpp", r"^(int Profile::getRandBaseQuality\(\) \{)$", r"\1 ssc_hook_slot(SSC_SLOT_QUAL);", 1),
    ("lib/profile/Profile.cpp", r"^(char\* Profile::predict\(char\* refSeq, int isRead1\) \{)$",
     r"\1 ssc_hook_mate(isRead1);", 1),
    ("lib/profile/Profile.cpp", r"^(\t\t)(k = getIndelSeq\(indelBaseIndxs\[j\]\);)$", r"\1ssc_hook_refpos(j); \2", 1),
    ("lib/profile/Profile.cpp", r"^(\t\t)(refIndx = getIndexOfBase\(sourceSeq\[j\]\);)$", r"\1ssc_hook_outpos(j); \2", 1),
    # ---- the three wall-clock seed sites
    ("lib/profile/Profile.cpp",
     r"^(\t\tunsigned seed = )chrono::system_clock::now\(\)\.time_since_epoch\(\)\.count\(\);$",
     r"\1ssc_hook_gc_seed(l);", 1),
    ("lib/genome/Genome.cpp", r"^(\t)srand\(time\(0\)\);$", r"\1srand(ssc_hook_plan_seed());", 1),
    # ---- plan dump + sample boundary (Genome::yieldReads, lib/genome/Genome.cpp:876-888, 944-957)
    ("lib/genome/Genome.cpp",
     r"^(\t+)(for\(k = 0; k < chrSegs\.size\(\); k\+\+\) \{\n\t+//Segment::yieldReads\(&chrSegs\[k\]\);)$",
     r"\1ssc_hook_dump_chr(chrSegs, curPopu, chr);\n\1\2", 2),
    ("lib/genome/Genome.cpp", r"^(\t+)(delete swp;)$", r"\1ssc_hook_sample_end(); \2", 2),
]


def run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def include_flags(ref):
    return ["-I" + d for d in sorted(glob.glob(os.path.join(ref, "lib", "*"))) if os.path.isdir(d)]


def lib_sources(ref):
    return sorted(glob.glob(os.path.join(ref, "lib", "*", "*.cpp")))


def compile_link(srcs, incs, out):
    """Compile every source to an object in a scratch dir (in parallel), then link."""
    from concurrent.futures import ThreadPoolExecutor
    objdir = tempfile.mkdtemp(prefix="ssc_refobj_")
    try:
        objs = [os.path.join(objdir, "%02d_%s.o" % (i, os.path.basename(s))) for i, s in enumerate(srcs)]
        with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
            list(ex.map(lambda so: run(["g++"] + CXXFLAGS + incs + ["-c", so[0], "-o", so[1]]), zip(srcs, objs)))
        run(["g++", "-pthread"] + objs + ["-o", out])
    finally:
        shutil.rmtree(objdir, ignore_errors=True)


def build_unmodified(ref, out):
    srcs = [os.path.join(ref, "src", "simuReads.cpp")] + lib_sources(ref)
    compile_link(srcs, include_flags(ref), out)


def build_instrumented(ref, out):
    tmp = tempfile.mkdtemp(prefix="ssc_refbuild_")
    try:
        patched = {}
        for rel, pat, rep, nmin in PATCHES:
            if rel not in patched:
                with open(os.path.join(ref, rel), "r", encoding="latin-1") as f:
                    patched[rel] = f.read()
            new, n = re.subn(pat, rep, patched[rel], flags=re.M)
            if n < nmin:
                raise SystemExit("patch anchor not found (%d < %d) in %s: %s" % (n, nmin, rel, pat))
            patched[rel] = new
        srcs = [os.path.join(ref, "src", "simuReads.cpp")]
        for s in lib_sources(ref):
            rel = os.path.relpath(s, ref)
            if rel == "lib/threadpool/ThreadPool.cpp":
                continue
            if rel in patched:
                dst = os.path.join(tmp, os.path.basename(s))
                with open(dst, "w", encoding="latin-1") as f:
                    f.write('#include "ssc_hooks.h"\n' + patched[rel])
                srcs.append(dst)
            else:
                srcs.append(s)
        srcs += [os.path.join(SHIM, "philox_pool.cpp"), os.path.join(SHIM, "hooks.cpp")]
        compile_link(srcs, include_flags(ref) + ["-I" + SHIM], out)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    if not os.path.isdir(os.path.join(a.ref, "lib")):
        print("reference tree %s not present: keeping prebuilt oracle/_ref" % a.ref)
        return 0
    os.makedirs(OUT, exist_ok=True)
    ref_bin = os.path.join(OUT, "simuReads_ref")
    phx_bin = os.path.join(OUT, "simuReads_philox")
    deps = glob.glob(os.path.join(SHIM, "*")) + [os.path.abspath(__file__), os.path.join(HERE, "ssc_oracle_philox.h")]
    newest = max(os.path.getmtime(p) for p in deps)
    if a.force or not os.path.exists(ref_bin):
        build_unmodified(a.ref, ref_bin)
    if a.force or not os.path.exists(phx_bin) or os.path.getmtime(phx_bin) < newest:
        build_instrumented(a.ref, phx_bin)
    print("oracle/_ref ready:", os.listdir(OUT))
    return 0


if __name__ == "__main__":
    sys.exit(main())
