"""ctypes binding of oracle/libkd_oracle.so -- TEST INFRASTRUCTURE ONLY.

``libkd_oracle.so`` is this repo's own CPU restatement of the reference search
(oracle/kd_oracle.cc): mode 0 reproduces the reference's processing order bit
for bit, mode 1 is the order-independent ("canonical") statement the CUDA
kernels implement.  Only tests/, ``__graft_entry__.smoke`` and bench.py's
CPU-baseline leg may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

from .kd_ref import BestPath, Options  # plain value types, shared

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkd_oracle.so")

REFERENCE_ORDER = 0
CANONICAL = 1
SIMPLE = 2  # order-independent SimpleDecoder (beam from Options.beam; other options unused)

STAT_NAMES = ("frames", "tokens_in", "tokens_expanded", "emit_arcs", "eps_arcs", "admitted",
              "extras", "emit_ties", "eps_ties", "tokens_out", "max_tokens", "binding_max",
              "binding_min", "first_binding_frame", "all_arcs_scanned")

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "kd_oracle.cc")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.kdo_last_error.restype = C.c_char_p
        L.kdo_graph_create.restype = C.c_void_p
        L.kdo_graph_create.argtypes = [C.c_int32, C.c_int32] + [C.c_void_p] * 6
        L.kdo_graph_destroy.argtypes = [C.c_void_p]
        L.kdo_decoder_create.restype = C.c_void_p
        L.kdo_decoder_create.argtypes = [C.c_void_p, C.c_float, C.c_int32, C.c_int32, C.c_float,
                                         C.c_float, C.c_int]
        L.kdo_decoder_destroy.argtypes = [C.c_void_p]
        L.kdo_decoder_set_options.argtypes = [C.c_void_p, C.c_float, C.c_int32, C.c_int32,
                                              C.c_float, C.c_float]
        L.kdo_decoder_init.argtypes = [C.c_void_p]
        L.kdo_decoder_advance.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_int32]
        L.kdo_decoder_num_frames_decoded.argtypes = [C.c_void_p]
        L.kdo_decoder_num_frames_decoded.restype = C.c_int32
        L.kdo_decoder_reached_final.argtypes = [C.c_void_p]
        L.kdo_decoder_dump_tokens.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.kdo_decoder_dump_tokens.restype = C.c_int64
        L.kdo_decoder_best_path.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64] + [C.c_void_p] * 5
        L.kdo_decoder_best_path.restype = C.c_int64
        L.kdo_decoder_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.kdo_decode_batch.restype = C.c_double
        L.kdo_decode_batch.argtypes = [
            C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
            C.c_float, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int32,
            C.c_int64] + [C.c_void_p] * 9
        _lib = L
    return _lib


def _err() -> str:
    return lib().kdo_last_error().decode("utf-8", "replace")


class OracleGraph:
    def __init__(self, g):
        self.g = g
        arrs = [np.ascontiguousarray(g.row_off, dtype=np.int64),
                np.ascontiguousarray(g.ilabel, dtype=np.int32),
                np.ascontiguousarray(g.olabel, dtype=np.int32),
                np.ascontiguousarray(g.weight, dtype=np.float32),
                np.ascontiguousarray(g.nextstate, dtype=np.int32),
                np.ascontiguousarray(g.final, dtype=np.float32)]
        self.h = lib().kdo_graph_create(g.num_states, g.start, *[a.ctypes.data for a in arrs])

    def __del__(self):
        if getattr(self, "h", None):
            lib().kdo_graph_destroy(self.h)
            self.h = None


class OracleDecoder:
    def __init__(self, graph: OracleGraph, opts: Options, mode: int = REFERENCE_ORDER):
        self.graph = graph
        self.h = lib().kdo_decoder_create(graph.h, *opts.args(), int(mode))
        if not self.h:
            raise RuntimeError(_err())

    def __del__(self):
        if getattr(self, "h", None):
            lib().kdo_decoder_destroy(self.h)
            self.h = None

    def set_options(self, opts: Options):
        lib().kdo_decoder_set_options(self.h, *opts.args())

    def init_decoding(self):
        if lib().kdo_decoder_init(self.h) != 0:
            raise RuntimeError(_err())

    def advance_decoding(self, logp: np.ndarray, offset: int = 0, max_num_frames: int = -1):
        logp = np.ascontiguousarray(logp, dtype=np.float32)
        if lib().kdo_decoder_advance(self.h, logp.ctypes.data, logp.shape[0], logp.shape[1],
                                     offset, max_num_frames) != 0:
            raise RuntimeError(_err())

    def decode(self, logp: np.ndarray):
        self.init_decoding()
        self.advance_decoding(logp)

    def num_frames_decoded(self) -> int:
        return lib().kdo_decoder_num_frames_decoded(self.h)

    def reached_final(self) -> bool:
        return bool(lib().kdo_decoder_reached_final(self.h))

    def tokens(self):
        n = lib().kdo_decoder_dump_tokens(self.h, 0, None, None)
        st = np.empty(n, dtype=np.int32)
        co = np.empty(n, dtype=np.float64)
        lib().kdo_decoder_dump_tokens(self.h, n, st.ctypes.data, co.ctypes.data)
        return st, co

    def stats(self) -> dict:
        v = np.zeros(len(STAT_NAMES), np.int64)
        lib().kdo_decoder_stats(self.h, v.ctypes.data)
        return dict(zip(STAT_NAMES, (int(x) for x in v)))

    def get_best_path(self, use_final_probs: bool = True, raw: bool = False) -> BestPath:
        cap = 4 * max(1, self.num_frames_decoded()) + 64
        while True:
            il = np.empty(cap, np.int32)
            ol = np.empty(cap, np.int32)
            gw = np.empty(cap, np.float32)
            aw = np.empty(cap, np.float32)
            f2 = np.zeros(2, np.float32)
            n = lib().kdo_decoder_best_path(self.h, int(use_final_probs), int(raw), cap,
                                            il.ctypes.data, ol.ctypes.data, gw.ctypes.data,
                                            aw.ctypes.data, f2.ctypes.data)
            if n < 0:
                e = np.empty(0, np.int32)
                return BestPath(False, e, e, np.empty(0, np.float32), np.empty(0, np.float32), f2)
            if n <= cap:
                return BestPath(True, il[:n].copy(), ol[:n].copy(), gw[:n].copy(), aw[:n].copy(), f2)
            cap = int(n)


def decode_batch(graph: OracleGraph, logp: np.ndarray, opts: Options, num_threads: int,
                 mode: int = REFERENCE_ORDER, rows: Optional[np.ndarray] = None,
                 use_final_probs: bool = True, want_paths: bool = True):
    """Returns (seconds, paths or None, reached_final, summed stats dict, per-utt stats)."""
    logp = np.ascontiguousarray(logp, dtype=np.float32)
    n, T, V = logp.shape
    rows = np.full(n, T, np.int32) if rows is None else np.ascontiguousarray(rows, np.int32)
    rf = np.zeros(n, np.int32)
    ssum = np.zeros(len(STAT_NAMES), np.int64)
    per = np.zeros((n, len(STAT_NAMES)), np.int64)
    if want_paths:
        cap = 4 * T + 64
        il = np.empty((n, cap), np.int32)
        ol = np.empty((n, cap), np.int32)
        gw = np.empty((n, cap), np.float32)
        aw = np.empty((n, cap), np.float32)
        f2 = np.zeros((n, 2), np.float32)
        cnt = np.zeros(n, np.int64)
        ptrs = [a.ctypes.data for a in (il, ol, gw, aw, f2, cnt)]
    else:
        cap = 0
        ptrs = [None] * 6
    secs = lib().kdo_decode_batch(graph.h, logp.ctypes.data, n, T, rows.ctypes.data, V,
                                  *opts.args(), int(use_final_probs), int(mode),
                                  int(num_threads), cap, *ptrs, rf.ctypes.data,
                                  ssum.ctypes.data, per.ctypes.data)
    if secs < 0:
        raise RuntimeError(_err())
    paths = None
    if want_paths:
        paths = []
        for u in range(n):
            k = int(cnt[u])
            if k < 0:
                e = np.empty(0, np.int32)
                paths.append(BestPath(False, e, e, np.empty(0, np.float32),
                                      np.empty(0, np.float32), f2[u]))
            else:
                paths.append(BestPath(True, il[u, :k].copy(), ol[u, :k].copy(),
                                      gw[u, :k].copy(), aw[u, :k].copy(), f2[u].copy()))
    return secs, paths, rf, dict(zip(STAT_NAMES, (int(x) for x in ssum))), per
