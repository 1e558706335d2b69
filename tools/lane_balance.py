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
PH = ("cycles_cutoff", "cycles_expand", "cycles_closure", "cycles_commit")
rows = []
for u in ids:
    s = dec.stats(u)
    rows.append([s[k] for k in PH] + [s["cycles_scan"], s["tokens_in"], s["emit_arcs"], s["arcs_evaluated"], s["candidates"]])
a = np.array(rows, dtype=np.float64)
tot = a[:, :4].sum(axis=1) / 1.965e6
toks, arcs = a[:, 5], a[:, 6]
print("kernel_ms", dec.last_advance_info()[0])
print("per-lane busy ms: mean %.1f  p50 %.1f  p90 %.1f  p99 %.1f  max %.1f" % (tot.mean(), np.median(tot), np.percentile(tot, 90), np.percentile(tot, 99), tot.max()))
print("corr(busy, arcs) %.2f  corr(busy, tokens) %.2f" % (np.corrcoef(tot, arcs)[0, 1], np.corrcoef(tot, toks)[0, 1]))
print("tokens_in: mean %.0f max %.0f ; arcs mean %.2e max %.2e" % (toks.mean(), toks.max(), arcs.mean(), arcs.max()))
# fixed vs token-proportional cost per phase: cycles/frame = a + b * tokens/frame (least squares over lanes)
x = toks / T
names = list(PH) + ["cycles_scan"]
order = np.argsort(tot)
heavy = order[-51:]
for j, nme in enumerate(names):
    y = a[:, j] / T
    b, c = np.polyfit(x, y, 1)
    print("%-15s mean %8.0f  = %8.0f + %6.1f * tokens   | heaviest 5%%: %8.0f (tokens %5.0f)" %
          (nme, y.mean(), c, b, y[heavy].mean(), x[heavy].mean()))
for j, nme in ((7, "arcs_evaluated"), (8, "candidates")):
    y = a[:, j] / T
    print("%-15s mean %8.0f | heaviest 5%%: %8.0f" % (nme, y.mean(), y[heavy].mean()))
rec = (a[:, 1] - a[:, 4]) / T
b, c = np.polyfit(x, rec, 1)
print("recombine       mean %8.0f  = %8.0f + %6.1f * tokens   | heaviest 5%%: %8.0f" % (rec.mean(), c, b, rec[heavy].mean()))
print("slowest lanes: (busy ms, tokens/frame, recomb cyc/frame, cand/frame)")
for u in order[-6:]:
    print("  lane %4d  %.1f ms  %6.0f tok  %8.0f  %6.0f" % (u, tot[u], x[u], rec[u], a[u, 8] / T))
print("fastest lanes:")
for u in order[:3]:
    print("  lane %4d  %.1f ms  %6.0f tok  %8.0f  %6.0f" % (u, tot[u], x[u], rec[u], a[u, 8] / T))
