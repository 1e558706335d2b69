#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref).

Run in the build container (where /root/reference exists and `make -C oracle ref`
has produced oracle/_ref/libkd_ref.so):

    python tests/golden/make_golden.py

Each fixture holds a small graph, float32 log-probs, decoder options and what the
reference's FasterDecoder produced for them: the best path after RemoveEpsLocal
(use_final_probs True and False), reached_final, and for every frame the token
list in the reference's own order (states, fp64 costs).  The reference ships no
decoder test or golden vector of its own, so these are the pin.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "kaldi-decoder_b200", "python"))

from kaldi_decoder_b200 import synth  # noqa: E402
from oracle import kd_ref  # noqa: E402

CASES = [
    # name, graph factory, T, n_utts, peak, options
    ("h20_bind", lambda: synth.make_h(20), 40, 3, 4.0,
     dict(beam=8.0, max_active=12, min_active=5, beam_delta=0.5, hash_ratio=2.0)),
    ("h20_default", lambda: synth.make_h(20), 40, 2, 6.0,
     dict(beam=16.0, max_active=2**31 - 1, min_active=20, beam_delta=0.5, hash_ratio=2.0)),
    ("hl300", lambda: synth.make_hl(300, 30, seed=11), 60, 3, 8.0,
     dict(beam=16.0, max_active=2**31 - 1, min_active=20, beam_delta=0.5, hash_ratio=2.0)),
    ("hlg300_peaky", lambda: synth.make_hlg(300, (20, 40), 30, seed=12), 80, 4, 10.0,
     dict(beam=20.0, max_active=7000, min_active=20, beam_delta=0.5, hash_ratio=2.0)),
    ("hlg300_bind", lambda: synth.make_hlg(300, (20, 40), 30, seed=12), 80, 3, 4.0,
     dict(beam=12.0, max_active=30, min_active=10, beam_delta=0.5, hash_ratio=2.0)),
    ("hlg300_nobeam", lambda: synth.make_hlg(300, (20, 40), 30, seed=13), 50, 2, 6.0,
     dict(beam=10.0, max_active=2**31 - 1, min_active=0, beam_delta=0.5, hash_ratio=2.0)),
]


def main():
    for name, make_graph, T, n_utts, peak, opts in CASES:
        g = make_graph()
        rg = kd_ref.RefGraph(g)
        out = dict(num_states=g.num_states, start=g.start, row_off=g.row_off, ilabel=g.ilabel,
                   olabel=g.olabel, weight=g.weight, nextstate=g.nextstate, final=g.final,
                   vocab=int(g.lm["vocab"]), n_utts=n_utts, T=T,
                   opts=np.array([opts["beam"], opts["max_active"], opts["min_active"],
                                  opts["beam_delta"], opts["hash_ratio"]], dtype=np.float64))
        for u in range(n_utts):
            lp = synth.make_logprobs(g, T, seed=1000 + 17 * u, peak=peak)
            out[f"logp{u}"] = lp
            dec = kd_ref.RefDecoder(rg, kd_ref.Options(**opts))
            dec.init_decoding()
            tok_n = []
            tok_states, tok_costs = [], []
            for f in range(T + 1):
                st, co = dec.tokens()
                tok_n.append(len(st))
                tok_states.append(st)
                tok_costs.append(co)
                if f < T:
                    dec.advance_decoding(lp, 0, 1)
            out[f"tok_n{u}"] = np.asarray(tok_n, np.int64)
            out[f"tok_states{u}"] = np.concatenate(tok_states)
            out[f"tok_costs{u}"] = np.concatenate(tok_costs)
            out[f"reached_final{u}"] = np.int32(dec.reached_final())
            for tag, ufp in (("t", True), ("f", False)):
                bp = dec.get_best_path(ufp)
                out[f"ok_{tag}{u}"] = np.int32(bp.ok)
                out[f"il_{tag}{u}"] = bp.ilabels
                out[f"ol_{tag}{u}"] = bp.olabels
                out[f"gw_{tag}{u}"] = bp.graph
                out[f"aw_{tag}{u}"] = bp.acoustic
                out[f"fin_{tag}{u}"] = bp.final
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, g.stats()["states"], "states", g.num_arcs, "arcs ->", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
