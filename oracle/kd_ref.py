"""ctypes binding of oracle/_ref/libkd_ref.so -- TEST INFRASTRUCTURE ONLY.

The library is the *unmodified* reference ``faster-decoder.cc`` compiled where
it lies under /root/reference by oracle/Makefile (``make ref``) plus the thin
C-ABI harness ``oracle/ref_harness.cc``.  Only tests/, ``__graft_entry__.smoke``
and bench.py's CPU-baseline / ``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libkd_ref.so")

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} not built; run `make -C oracle ref`")
        L = C.CDLL(LIB_PATH)
        L.kdref_last_error.restype = C.c_char_p
        L.kdref_graph_create.restype = C.c_void_p
        L.kdref_graph_create.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.kdref_graph_destroy.argtypes = [C.c_void_p]
        L.kdref_decoder_create.restype = C.c_void_p
        L.kdref_decoder_create.argtypes = [C.c_void_p, C.c_float, C.c_int32, C.c_int32,
                                           C.c_float, C.c_float]
        L.kdref_decoder_destroy.argtypes = [C.c_void_p]
        L.kdref_decoder_set_options.argtypes = [C.c_void_p, C.c_float, C.c_int32, C.c_int32,
                                                C.c_float, C.c_float]
        L.kdref_decoder_init.argtypes = [C.c_void_p]
        L.kdref_decoder_advance.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                            C.c_int32, C.c_int32]
        L.kdref_decoder_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
        L.kdref_decoder_num_frames_decoded.argtypes = [C.c_void_p]
        L.kdref_decoder_num_frames_decoded.restype = C.c_int32
        L.kdref_decoder_reached_final.argtypes = [C.c_void_p]
        L.kdref_decoder_dump_tokens.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.kdref_decoder_dump_tokens.restype = C.c_int64
        L.kdref_decoder_best_path.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.kdref_decoder_best_path.restype = C.c_int64
        L.kdref_decode_batch.restype = C.c_double
        L.kdref_decode_batch.argtypes = [
            C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
            C.c_float, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int, C.c_int32,
            C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_void_p, C.c_void_p]
        if hasattr(L, "kdref_simple_create"):
            L.kdref_simple_create.restype = C.c_void_p
            L.kdref_simple_create.argtypes = [C.c_void_p, C.c_float]
            L.kdref_simple_destroy.argtypes = [C.c_void_p]
            L.kdref_simple_init.argtypes = [C.c_void_p]
            L.kdref_simple_advance.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                               C.c_int32, C.c_int32]
            L.kdref_simple_num_frames_decoded.argtypes = [C.c_void_p]
            L.kdref_simple_num_frames_decoded.restype = C.c_int32
            L.kdref_simple_reached_final.argtypes = [C.c_void_p]
            L.kdref_simple_final_relative_cost.argtypes = [C.c_void_p]
            L.kdref_simple_final_relative_cost.restype = C.c_float
            L.kdref_simple_best_path.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
            L.kdref_simple_best_path.restype = C.c_int64
        _lib = L
    return _lib


def _err() -> str:
    return lib().kdref_last_error().decode("utf-8", "replace")


class Options:
    """Mirror of FasterDecoderOptions (faster-decoder.h:24-63)."""

    def __init__(self, beam=16.0, max_active=2**31 - 1, min_active=20, beam_delta=0.5,
                 hash_ratio=2.0):
        self.beam, self.max_active, self.min_active = float(beam), int(max_active), int(min_active)
        self.beam_delta, self.hash_ratio = float(beam_delta), float(hash_ratio)

    def args(self):
        return (self.beam, self.max_active, self.min_active, self.beam_delta, self.hash_ratio)


class RefGraph:
    def __init__(self, g):
        self.g = g
        self._keep = [np.ascontiguousarray(g.row_off, dtype=np.int64),
                      np.ascontiguousarray(g.ilabel, dtype=np.int32),
                      np.ascontiguousarray(g.olabel, dtype=np.int32),
                      np.ascontiguousarray(g.weight, dtype=np.float32),
                      np.ascontiguousarray(g.nextstate, dtype=np.int32),
                      np.ascontiguousarray(g.final, dtype=np.float32)]
        self.h = lib().kdref_graph_create(g.num_states, g.start,
                                          *[a.ctypes.data for a in self._keep])
        if not self.h:
            raise RuntimeError(_err())

    def __del__(self):
        if getattr(self, "h", None):
            lib().kdref_graph_destroy(self.h)
            self.h = None


class BestPath:
    __slots__ = ("ok", "ilabels", "olabels", "graph", "acoustic", "final")

    def __init__(self, ok, il, ol, gw, aw, final):
        self.ok, self.ilabels, self.olabels = ok, il, ol
        self.graph, self.acoustic, self.final = gw, aw, final

    @property
    def isyms(self):
        return self.ilabels[self.ilabels != 0]

    @property
    def osyms(self):
        return self.olabels[self.olabels != 0]

    @property
    def total_cost(self) -> float:
        """graph + acoustic cost of the path, incl. the final weight (float64 sum)."""
        return float(self.graph.astype(np.float64).sum() + self.acoustic.astype(np.float64).sum()
                     + float(self.final[0]) + float(self.final[1]))


class RefDecoder:
    """The reference FasterDecoder behind the harness's C-ABI."""

    def __init__(self, graph: RefGraph, opts: Options):
        self.graph = graph
        self.h = lib().kdref_decoder_create(graph.h, *opts.args())
        if not self.h:
            raise RuntimeError(_err())
        self._keep = None

    def __del__(self):
        if getattr(self, "h", None):
            lib().kdref_decoder_destroy(self.h)
            self.h = None

    def set_options(self, opts: Options):
        lib().kdref_decoder_set_options(self.h, *opts.args())

    def init_decoding(self):
        if lib().kdref_decoder_init(self.h) != 0:
            raise RuntimeError(_err())

    def advance_decoding(self, logp: np.ndarray, offset: int = 0, max_num_frames: int = -1):
        logp = np.ascontiguousarray(logp, dtype=np.float32)
        self._keep = logp
        if lib().kdref_decoder_advance(self.h, logp.ctypes.data, logp.shape[0], logp.shape[1],
                                       offset, max_num_frames) != 0:
            raise RuntimeError(_err())

    def decode(self, logp: np.ndarray):
        logp = np.ascontiguousarray(logp, dtype=np.float32)
        if lib().kdref_decoder_decode(self.h, logp.ctypes.data, logp.shape[0], logp.shape[1]) != 0:
            raise RuntimeError(_err())

    def num_frames_decoded(self) -> int:
        return lib().kdref_decoder_num_frames_decoded(self.h)

    def reached_final(self) -> bool:
        return bool(lib().kdref_decoder_reached_final(self.h))

    def tokens(self) -> Tuple[np.ndarray, np.ndarray]:
        """(states, costs) of the live tokens, in the reference's HashList order."""
        n = lib().kdref_decoder_dump_tokens(self.h, 0, None, None)
        st = np.empty(n, dtype=np.int32)
        co = np.empty(n, dtype=np.float64)
        lib().kdref_decoder_dump_tokens(self.h, n, st.ctypes.data, co.ctypes.data)
        return st, co

    def get_best_path(self, use_final_probs: bool = True, cap: Optional[int] = None) -> BestPath:
        cap = cap or (4 * max(1, self.num_frames_decoded()) + 64)
        while True:
            il = np.empty(cap, np.int32)
            ol = np.empty(cap, np.int32)
            gw = np.empty(cap, np.float32)
            aw = np.empty(cap, np.float32)
            f2 = np.zeros(2, np.float32)
            n = lib().kdref_decoder_best_path(self.h, int(use_final_probs), cap, il.ctypes.data,
                                              ol.ctypes.data, gw.ctypes.data, aw.ctypes.data,
                                              f2.ctypes.data)
            if n == -2:
                raise RuntimeError(_err())
            if n == -1:
                e = np.empty(0, np.int32)
                return BestPath(False, e, e, np.empty(0, np.float32), np.empty(0, np.float32), f2)
            if n <= cap:
                return BestPath(True, il[:n].copy(), ol[:n].copy(), gw[:n].copy(), aw[:n].copy(), f2)
            cap = int(n)


class RefSimpleDecoder:
    """The reference SimpleDecoder (simple-decoder.cc, unmodified) behind the harness's C-ABI."""

    def __init__(self, graph: RefGraph, beam: float):
        self.graph = graph
        self.h = lib().kdref_simple_create(graph.h, float(beam))
        if not self.h:
            raise RuntimeError(_err())
        self._keep = None

    def __del__(self):
        if getattr(self, "h", None):
            lib().kdref_simple_destroy(self.h)
            self.h = None

    def init_decoding(self):
        if lib().kdref_simple_init(self.h) != 0:
            raise RuntimeError(_err())

    def advance_decoding(self, logp: np.ndarray, offset: int = 0, max_num_frames: int = -1):
        logp = np.ascontiguousarray(logp, dtype=np.float32)
        self._keep = logp
        if lib().kdref_simple_advance(self.h, logp.ctypes.data, logp.shape[0], logp.shape[1],
                                      offset, max_num_frames) != 0:
            raise RuntimeError(_err())

    def num_frames_decoded(self) -> int:
        return lib().kdref_simple_num_frames_decoded(self.h)

    def reached_final(self) -> bool:
        return bool(lib().kdref_simple_reached_final(self.h))

    def final_relative_cost(self) -> float:
        return float(lib().kdref_simple_final_relative_cost(self.h))

    def get_best_path(self, use_final_probs: bool = True) -> BestPath:
        cap = 4 * max(1, self.num_frames_decoded()) + 64
        while True:
            il = np.empty(cap, np.int32)
            ol = np.empty(cap, np.int32)
            gw = np.empty(cap, np.float32)
            aw = np.empty(cap, np.float32)
            f2 = np.zeros(2, np.float32)
            n = lib().kdref_simple_best_path(self.h, int(use_final_probs), cap, il.ctypes.data,
                                             ol.ctypes.data, gw.ctypes.data, aw.ctypes.data,
                                             f2.ctypes.data)
            if n == -2:
                raise RuntimeError(_err())
            if n == -1:
                e = np.empty(0, np.int32)
                return BestPath(False, e, e, np.empty(0, np.float32), np.empty(0, np.float32), f2)
            if n <= cap:
                return BestPath(True, il[:n].copy(), ol[:n].copy(), gw[:n].copy(), aw[:n].copy(), f2)
            cap = int(n)


def decode_batch(graph: RefGraph, logp: np.ndarray, opts: Options, num_threads: int,
                 rows: Optional[np.ndarray] = None, use_final_probs: bool = True,
                 want_paths: bool = True):
    """Decodes logp [n_utts, T, V] with one reference decoder per thread.
    Returns (seconds, list[BestPath] or None, reached_final int32[n_utts])."""
    logp = np.ascontiguousarray(logp, dtype=np.float32)
    n, T, V = logp.shape
    rows = np.full(n, T, np.int32) if rows is None else np.ascontiguousarray(rows, np.int32)
    rf = np.zeros(n, np.int32)
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
    secs = lib().kdref_decode_batch(graph.h, logp.ctypes.data, n, T, rows.ctypes.data, V,
                                    *opts.args(), int(use_final_probs), int(num_threads),
                                    cap, *ptrs, rf.ctypes.data)
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
                assert k <= cap
                paths.append(BestPath(True, il[u, :k].copy(), ol[u, :k].copy(),
                                      gw[u, :k].copy(), aw[u, :k].copy(), f2[u].copy()))
    return secs, paths, rf
