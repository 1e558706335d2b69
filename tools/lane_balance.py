#!/usr/bin/env python
"""Per-lane cycle totals of one bench step (how uneven are the lanes?)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "kaldi-decoder_b200", "python"))
import numpy as np, torch
import bench
from kaldi_decoder_b200 import capi, synth

lanes, T = 1024, 1000
g = synth.make_config_graph("C3")
dg = capi.DeviceGraph.from_graph(g)
dec = capi.LaneDecoder(dg, capi.make_options(**bench.OPTS), max_lanes=lanes, hash_capacity=1 << 18)
logp = bench.make_device_logprobs(g, lanes, T, 3, 12.0, torch.device("cuda", 0))
ids = list(range(lanes)); ptrs = [logp[u].data_ptr() for u in ids]
for _ in range(2):
    dec.init(ids); dec.advance_ptrs(ids, ptrs, [T] * lanes, 500, None, -1, capi.KD_MEM_DEVICE)
tot = []; arcs = []; toks = []
for u in ids:
    s = dec.stats(u)
    tot.append(sum(s[k] for k in s if k.startswith("cycles"))); arcs.append(s["emit_arcs"]); toks.append(s["tokens_in"])
tot = np.array(tot) / 1.965e6; arcs = np.array(arcs); toks = np.array(toks)
print("kernel_ms", dec.last_advance_info()[0])
print("per-lane busy ms: mean %.1f  p50 %.1f  p90 %.1f  p99 %.1f  max %.1f" % (tot.mean(), np.median(tot), np.percentile(tot, 90), np.percentile(tot, 99), tot.max()))
print("corr(busy, arcs) %.2f  corr(busy, tokens) %.2f" % (np.corrcoef(tot, arcs)[0, 1], np.corrcoef(tot, toks)[0, 1]))
print("tokens_in: mean %.0f max %.0f ; arcs mean %.2e max %.2e" % (toks.mean(), toks.max(), arcs.mean(), arcs.max()))
