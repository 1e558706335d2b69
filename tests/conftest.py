"""pytest configuration: the `gpu` marker and import paths.

`-m "not gpu"` runs on a CPU-only box (oracle vs reference/golden vectors, host
logic, C-ABI symbol checks); `-m gpu` runs the CUDA path through the C ABI and
compares it with the oracle.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "kaldi-decoder_b200", "python"), os.path.dirname(__file__)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
