#!/usr/bin/env python
"""Where the host-visible time of one bench step goes (init / advance / best-path prepare /
fetch / Python assembly), device-resident input."""
import os, sys, time
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
la = dec._lanes(ids)
lib = capi.lib()
for it in range(4):
    t = [time.perf_counter()]
    dec.init(ids); t.append(time.perf_counter())
    dec.advance_ptrs(ids, ptrs, [T] * lanes, 500, None, -1, capi.KD_MEM_DEVICE); t.append(time.perf_counter())
    n = la.size
    ok = np.zeros(n, np.int32); rf = np.zeros(n, np.int32); cnt = np.zeros(n, np.int64)
    capi._check(lib.kd_decoder_best_path_prepare(dec.h, n, la.ctypes.data, 1, ok.ctypes.data, rf.ctypes.data, cnt.ctypes.data)); t.append(time.perf_counter())
    off = np.zeros(n, np.int64); off[1:] = np.cumsum(cnt)[:-1]; total = int(cnt.sum())
    il = np.empty(total, np.int32); ol = np.empty(total, np.int32); gw = np.empty(total, np.float32); aw = np.empty(total, np.float32)
    f2 = np.zeros((n, 2), np.float32); t.append(time.perf_counter())
    capi._check(lib.kd_decoder_best_path_fetch(dec.h, n, la.ctypes.data, off.ctypes.data, total, il.ctypes.data, ol.ctypes.data, gw.ctypes.data, aw.ctypes.data, f2.ctypes.data)); t.append(time.perf_counter())
    out = []
    for i in range(n):
        a, b = int(off[i]), int(off[i] + cnt[i])
        out.append(capi.RawPath(ok[i], rf[i], il[a:b].copy(), ol[a:b].copy(), gw[a:b].copy(), aw[a:b].copy(), f2[i].copy()))
    t.append(time.perf_counter())
    names = ["init", "advance", "prepare", "alloc", "fetch", "assemble"]
    print(" ".join(f"{nm}={1e3*(t[i+1]-t[i]):.2f}ms" for i, nm in enumerate(names)), "kernel_ms=%.2f" % dec.last_advance_info()[0], "total arcs", total)
