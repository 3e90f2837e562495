#!/usr/bin/env python3
"""Measures the instruction-issue ceiling of the generation kernel on this GPU (ssc_issue_floor, csrc/floor.cu):
mode 0 = Philox4x32-10 only, mode 1 = Philox + fast per-base path + byte stores.  One JSON line per read length."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simuscop_b200 import cuda_binding  # noqa: E402


def main():
    pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
    g = cuda_binding.Generator(0)
    for rl in (151, 125, 75, 74):
        row = {"read_length": rl, "pairs_per_launch": pairs}
        for mode, name in ((0, "philox_only"), (1, "philox_plus_fast_path")):
            ms = min(g.issue_floor(mode, rl, pairs, 5) for _ in range(3))
            row[name] = {"ms_per_launch": ms, "gbases_per_s": 2.0 * rl * pairs / ms / 1e6,
                         "lane_cycles_per_s": 2.0 * ((rl + 31) // 32) * 32 * pairs / ms / 1e6}
        print(json.dumps(row), flush=True)
    g.close()


if __name__ == "__main__":
    main()
