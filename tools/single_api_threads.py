#!/usr/bin/env python
"""Throughput of the UNCHANGED single-utterance API (`FasterDecoder.decode(DecodableCtc(m))` +
`get_best_path()`, the calls an icefall script makes) when it is driven from 1..N Python threads,
one FasterDecoder per thread on a shared DeviceGraph (C3 bench workload, host numpy input).

    python tools/single_api_threads.py [--threads 1,4,16,64] [--utts 8] [--frames 1000]
"""
import argparse, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "kaldi-decoder_b200", "python"))
import numpy as np
import bench
import kaldi_decoder as kd
from kaldi_decoder_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--threads", default="1,4,16,64")
ap.add_argument("--utts", type=int, default=8, help="utterances each thread decodes (timed)")
ap.add_argument("--frames", type=int, default=1000)
ap.add_argument("--config", default="C3")
ap.add_argument("--fresh", action="store_true",
                help="construct FasterDecoder(fst, opts) for every utterance, as icefall's "
                     "jit_pretrained_decode_with_HLG.py does (A/B: KD_B200_GRAPH_CACHE=0 "
                     "KD_B200_IDLE_DECODERS=0 switch the sharing off)")
args = ap.parse_args()

g = synth.make_config_graph(args.config)
fst = kd.StdConstFst.from_arrays(g.num_states, g.start, g.row_off, g.ilabel, g.olabel, g.weight,
                                 g.nextstate, g.final)
o = bench.OPTS
opts = kd.FasterDecoderOptions(beam=o["beam"], max_active=o["max_active"], min_active=o["min_active"],
                               beam_delta=o["beam_delta"], hash_ratio=o["hash_ratio"])
n_mats = 32
mats = synth.make_batch(g, n_mats, args.frames, seed=3, peak=12.0)
dc = kd.DeviceConfig()
dc.hash_capacity = 1 << 18
dc.arena_records = 6_000_000
print("cpus", len(os.sched_getaffinity(0)), "config", args.config, "frames", args.frames)
if args.fresh:
    print("graph cache", os.environ.get("KD_B200_GRAPH_CACHE", "default"),
          "idle decoders", os.environ.get("KD_B200_IDLE_DECODERS", "default"))
    for r in range(args.utts + 1):
        t0 = time.perf_counter()
        dec = kd.FasterDecoder(fst, opts)
        t1 = time.perf_counter()
        dec.decode(kd.DecodableCtc(mats[r % n_mats]))
        rf = dec.reached_final()
        ok, best = dec.get_best_path()
        n_words = len(kd.get_linear_symbol_sequence(best)[2])
        t2 = time.perf_counter()
        del dec
        t3 = time.perf_counter()
        print("utterance %d: construct %7.1f ms  decode+path %6.1f ms  destroy %6.1f ms  (%d words, uploads %d)" %
              (r, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, n_words, kd.graph_uploads()))
    sys.exit(0)
graph = kd.DeviceGraph(fst)
for n in [int(x) for x in args.threads.split(",")]:
    decs = [kd.FasterDecoder(graph, opts, dc) for _ in range(n)]
    errors, words = [], [0] * n
    start = threading.Barrier(n + 1)

    def work(k):
        try:
            dec = decs[k]
            dec.decode(kd.DecodableCtc(mats[k % n_mats]))  # warm-up (buffers, streams)
            dec.get_best_path()
            start.wait()
            for r in range(args.utts):
                dec.decode(kd.DecodableCtc(mats[(k + r) % n_mats]))
                ok, best = dec.get_best_path()
                words[k] += len(kd.get_linear_symbol_sequence(best)[2])
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))
            try:
                start.abort()
            except Exception:  # noqa: BLE001
                pass

    th = [threading.Thread(target=work, args=(k,)) for k in range(n)]
    for t in th:
        t.start()
    try:
        start.wait()
    except threading.BrokenBarrierError:
        pass
    t0 = time.perf_counter()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    if errors:
        print("threads %3d: ERROR %s" % (n, errors[0]))
        break
    print("threads %3d: %8.0f frames/s  (%.1f ms per utterance per thread, %d words)" %
          (n, n * args.utts * args.frames / dt, dt / args.utts * 1e3, sum(words)))
    del decs
