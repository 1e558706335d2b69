"""Shared helpers of the test-suite (graphs, option sets, comparisons)."""
import functools

import numpy as np

from kaldi_decoder_b200 import synth
from oracle import kd_oracle, kd_ref

OPTION_SETS = [
    dict(beam=20.0, max_active=7000, min_active=20),
    dict(beam=8.0, max_active=50, min_active=5),
    dict(beam=16.0, max_active=2**31 - 1, min_active=0),
    dict(beam=12.0, max_active=200, min_active=20),
    dict(beam=20.0, max_active=30, min_active=29),
]


@functools.lru_cache(maxsize=None)
def small_graph(name: str):
    if name == "H":
        return synth.make_h(50)
    if name == "HL":
        return synth.make_hl(2000, 50, seed=1)
    if name == "HLG":
        return synth.make_hlg(2000, (100, 200), 50, seed=2)
    raise KeyError(name)


def sorted_tokens(states, costs):
    o = np.argsort(states, kind="stable")
    return np.asarray(states)[o], np.asarray(costs)[o]


def strip(x):
    x = np.asarray(x)
    return x[x != 0]


def same_labels(a, b) -> bool:
    return np.array_equal(strip(a.ilabels), strip(b.ilabels)) and \
        np.array_equal(strip(a.olabels), strip(b.olabels))


def rel_close(a: float, b: float, tol: float = 1e-4) -> bool:
    return abs(a - b) <= tol * max(1.0, abs(a), abs(b))
