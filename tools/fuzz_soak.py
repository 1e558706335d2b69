#!/usr/bin/env python
"""One-off soak: the random-FST per-frame parity test of tests/test_gpu_parity.py over many
more seeds, plus SimpleDecoder mode.  python tools/fuzz_soak.py [first_seed] [n_seeds]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "kaldi-decoder_b200", "python")):
    sys.path.insert(0, p)
import test_gpu_parity as T
import test_gpu_simple as S

first = int(sys.argv[1]) if len(sys.argv) > 1 else 6
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
bad = 0
for seed in range(first, first + n):
    try:
        T.test_fuzz_random_fsts_every_frame(seed)
    except AssertionError as e:
        bad += 1
        print("FASTER seed", seed, "FAILED", e)
for seed in range(first, first + max(3, n // 3)):
    try:
        S.test_simple_search_random_fsts.__wrapped__(seed % 3) if hasattr(S.test_simple_search_random_fsts, "__wrapped__") else S.test_simple_search_random_fsts(seed % 3)
    except AssertionError as e:
        bad += 1
        print("SIMPLE seed", seed, "FAILED", e)
print("soak done:", n, "seeds,", bad, "failures")
