"""ctypes binding of the C ABI in include/kd_capi.h (libkd_b200.so).

This is the same boundary the C++ classes use; it exists so Python callers and
the parity tests can drive lanes in batches without going through pybind11.
There is no fallback: if the CUDA library is missing or no GPU is present the
calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.normpath(os.path.join(_HERE, "..", "..", "lib"))
# KD_B200_LIB selects another build of the same ABI (A/B experiments)
LIB_PATH = os.environ.get("KD_B200_LIB") or os.path.join(LIB_DIR, "libkd_b200.so")

KD_OK = 0
KD_SEARCH_FASTER, KD_SEARCH_SIMPLE = 0, 1
KD_MEM_HOST = 0
KD_MEM_DEVICE = 1
KD_ADVANCE_INIT = 1
KD_ADVANCE_FINALIZE = 2
INT32_MAX = 2**31 - 1


class KdOptions(C.Structure):
    _fields_ = [("beam", C.c_float), ("max_active", C.c_int32), ("min_active", C.c_int32),
                ("beam_delta", C.c_float), ("hash_ratio", C.c_float)]


class KdConfig(C.Structure):
    _fields_ = [("max_lanes", C.c_int32), ("hash_capacity", C.c_int32),
                ("arena_records", C.c_int64), ("threads_per_lane", C.c_int32),
                ("chunk_frames", C.c_int32), ("search", C.c_int32)]


class KdStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("frames", "tokens_in", "tokens_expanded", "emit_arcs",
                                         "eps_arcs", "tokens_out", "max_tokens", "eps_sweeps",
                                         "cycles_cutoff", "cycles_expand", "cycles_closure",
                                         "cycles_commit", "slots_claimed", "candidates",
                                         "arcs_evaluated", "cycles_scan", "arena_compactions",
                                         "cycles_input_wait", "table_retries")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


EXPORTED = (
    "kd_last_error", "kd_device_count", "kd_graph_create", "kd_graph_destroy", "kd_graph_info",
    "kd_decoder_create", "kd_decoder_destroy", "kd_decoder_set_options", "kd_decoder_reset", "kd_decoder_init",
    "kd_decoder_advance", "kd_decoder_num_frames_decoded", "kd_decoder_reached_final",
    "kd_decoder_best_path_prepare", "kd_decoder_best_path_fetch", "kd_decoder_best_path_view",
    "kd_decoder_best_path",
    "kd_decoder_dump_tokens", "kd_decoder_stats", "kd_decoder_last_advance_info",
    "kd_decoder_info", "kd_decoder_final_relative_cost", "kd_decoder_advance_async",
    "kd_decoder_wait", "kd_decoder_result_view", "kd_decoder_span_begin", "kd_decoder_span_end",
)

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a).  There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        L.kd_last_error.restype = C.c_char_p
        L.kd_device_count.argtypes = [C.POINTER(C.c_int)]
        L.kd_graph_create.argtypes = [C.c_int, i32, i32, vp, vp, vp, vp, vp, vp, C.POINTER(vp)]
        L.kd_graph_destroy.argtypes = [vp]
        L.kd_graph_info.argtypes = [vp, vp]
        L.kd_decoder_create.argtypes = [vp, C.POINTER(KdOptions), C.POINTER(KdConfig), C.POINTER(vp)]
        L.kd_decoder_destroy.argtypes = [vp]
        L.kd_decoder_set_options.argtypes = [vp, C.POINTER(KdOptions)]
        L.kd_decoder_reset.argtypes = [vp]
        L.kd_decoder_init.argtypes = [vp, i32, vp]
        L.kd_decoder_advance.argtypes = [vp, i32, vp, vp, vp, i32, vp, i32, C.c_int]
        L.kd_decoder_num_frames_decoded.argtypes = [vp, i32, C.POINTER(i32)]
        L.kd_decoder_reached_final.argtypes = [vp, i32, C.POINTER(i32)]
        L.kd_decoder_best_path_prepare.argtypes = [vp, i32, vp, C.c_int, vp, vp, vp]
        L.kd_decoder_best_path_fetch.argtypes = [vp, i32, vp, vp, i64, vp, vp, vp, vp, vp]
        L.kd_decoder_best_path_view.argtypes = [vp, i32, vp, vp, i64, C.POINTER(vp), C.POINTER(vp),
                                                C.POINTER(vp), C.POINTER(vp), vp]
        L.kd_decoder_best_path.argtypes = [vp, i32, C.c_int, i64, vp, vp, vp, vp,
                                           C.POINTER(i64), vp, C.POINTER(i32), C.POINTER(i32)]
        L.kd_decoder_dump_tokens.argtypes = [vp, i32, i64, vp, vp, C.POINTER(i64)]
        L.kd_decoder_stats.argtypes = [vp, i32, C.POINTER(KdStats)]
        L.kd_decoder_last_advance_info.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(i32)]
        L.kd_decoder_info.argtypes = [vp, vp]
        L.kd_decoder_final_relative_cost.argtypes = [vp, i32, C.POINTER(C.c_float)]
        L.kd_decoder_advance_async.argtypes = [vp, i32, vp, vp, vp, i32, vp, i32, C.c_int, C.c_int,
                                               vp, C.POINTER(i64)]
        L.kd_decoder_wait.argtypes = [vp, i64]
        L.kd_decoder_span_begin.argtypes = [vp]
        L.kd_decoder_span_end.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(i32)]
        L.kd_decoder_result_view.argtypes = [vp, i64, C.c_int, C.POINTER(i32), C.POINTER(vp),
                                             C.POINTER(vp), C.POINTER(vp), vp, vp, vp, vp]
        _lib = L
    return _lib


class KdError(RuntimeError):
    pass


def _check(rc: int):
    if rc != KD_OK:
        raise KdError(lib().kd_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    n = C.c_int(0)
    lib().kd_device_count(C.byref(n))
    return n.value


def make_options(beam=16.0, max_active=INT32_MAX, min_active=20, beam_delta=0.5,
                 hash_ratio=2.0) -> KdOptions:
    return KdOptions(float(beam), int(max_active), int(min_active), float(beam_delta),
                     float(hash_ratio))


class DeviceGraph:
    """kd_graph: the decoding graph resident on one GPU."""

    def __init__(self, num_states: int, start: int, row_off, ilabel, olabel, weight, nextstate,
                 final, device: int = 0):
        arrs = [np.ascontiguousarray(row_off, dtype=np.int64),
                np.ascontiguousarray(ilabel, dtype=np.int32),
                np.ascontiguousarray(olabel, dtype=np.int32),
                np.ascontiguousarray(weight, dtype=np.float32),
                np.ascontiguousarray(nextstate, dtype=np.int32),
                np.ascontiguousarray(final, dtype=np.float32)]
        h = C.c_void_p()
        _check(lib().kd_graph_create(int(device), int(num_states), int(start),
                                     *[a.ctypes.data for a in arrs], C.byref(h)))
        self.h = h
        self.device = int(device)

    @classmethod
    def from_graph(cls, g, device: int = 0) -> "DeviceGraph":
        return cls(g.num_states, g.start, g.row_off, g.ilabel, g.olabel, g.weight, g.nextstate,
                   g.final, device=device)

    def info(self) -> dict:
        v = np.zeros(5, np.int64)
        _check(lib().kd_graph_info(self.h, v.ctypes.data))
        return dict(zip(("num_states", "num_arcs", "num_eps_arcs", "max_ilabel", "device"),
                        (int(x) for x in v)))

    def close(self):
        if getattr(self, "h", None):
            lib().kd_graph_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RawPath:
    """One arc per token of the best path, before RemoveEpsLocal."""
    __slots__ = ("ok", "reached_final", "ilabels", "olabels", "graph", "acoustic", "final")

    def __init__(self, ok, reached_final, il, ol, gw, aw, final):
        self.ok, self.reached_final = bool(ok), bool(reached_final)
        self.ilabels, self.olabels, self.graph, self.acoustic, self.final = il, ol, gw, aw, final

    @property
    def isyms(self):
        return self.ilabels[self.ilabels != 0]

    @property
    def osyms(self):
        return self.olabels[self.olabels != 0]

    @property
    def total_cost(self) -> float:
        return float(self.graph.astype(np.float64).sum() + self.acoustic.astype(np.float64).sum()
                     + float(self.final[0]) + float(self.final[1]))


class PathBatch:
    """The best paths of a batch of lanes.  All arcs live in one int32 word buffer; lane i's
    four arrays (ilabel, olabel, graph, acoustic) are `lens[i]` words each, starting at
    `word_offsets[i, 0..3]`.  Item i is a RawPath (views, no per-lane copies)."""

    def __init__(self, ok, reached_final, lens, words, word_offsets, final, d2h_bytes=0):
        self.ok, self.reached_final, self.lens = ok, reached_final, lens
        self.words, self.word_offsets, self.final = words, word_offsets, final
        self._fwords = words.view(np.float32)
        self.d2h_bytes = int(d2h_bytes)  # bytes the device->host copy of these paths moved

    def __len__(self):
        return int(self.ok.shape[0])

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        n = int(self.lens[i])
        o = self.word_offsets[i]
        return RawPath(self.ok[i], self.reached_final[i], self.words[int(o[0]):int(o[0]) + n],
                       self.words[int(o[1]):int(o[1]) + n], self._fwords[int(o[2]):int(o[2]) + n],
                       self._fwords[int(o[3]):int(o[3]) + n], self.final[i])

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class LaneDecoder:
    """kd_decoder: `max_lanes` independent utterance lanes on one GPU."""

    def __init__(self, graph: DeviceGraph, opts: KdOptions, max_lanes: int = 1,
                 hash_capacity: int = 0, arena_records: int = 0, threads_per_lane: int = 0,
                 chunk_frames: int = 0, search: int = 0):
        self.graph = graph
        cfg = KdConfig(int(max_lanes), int(hash_capacity), int(arena_records),
                       int(threads_per_lane), int(chunk_frames), int(search))
        h = C.c_void_p()
        _check(lib().kd_decoder_create(graph.h, C.byref(opts), C.byref(cfg), C.byref(h)))
        self.h = h
        self.max_lanes = int(max_lanes)
        self._keep = None

    def close(self):
        if getattr(self, "h", None):
            lib().kd_decoder_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _lanes(lanes) -> np.ndarray:
        return np.ascontiguousarray(lanes, dtype=np.int32).reshape(-1)

    def set_options(self, opts: KdOptions):
        _check(lib().kd_decoder_set_options(self.h, C.byref(opts)))

    def reset(self):
        """Every lane uninitialised again, as after creation; buffers and streams are kept."""
        _check(lib().kd_decoder_reset(self.h))

    def init(self, lanes):
        la = self._lanes(lanes)
        _check(lib().kd_decoder_init(self.h, la.size, la.ctypes.data))

    def advance_ptrs(self, lanes, ptrs: Sequence[int], rows, cols: int, offsets=None,
                     max_num_frames: int = -1, mem_kind: int = KD_MEM_HOST):
        la = self._lanes(lanes)
        pa = (C.c_void_p * la.size)(*[int(p) for p in ptrs])
        ra = np.ascontiguousarray(rows, dtype=np.int32).reshape(-1)
        oa = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.int32).reshape(-1)
        _check(lib().kd_decoder_advance(self.h, la.size, la.ctypes.data, pa, ra.ctypes.data,
                                        int(cols), None if oa is None else oa.ctypes.data,
                                        int(max_num_frames), int(mem_kind)))

    def advance(self, lanes, mats: Sequence[np.ndarray], offsets=None, max_num_frames: int = -1):
        """Host float32 matrices, one [rows, cols] per lane."""
        mats = [np.ascontiguousarray(m, dtype=np.float32) for m in mats]
        self._keep = mats
        cols = mats[0].shape[1]
        assert all(m.ndim == 2 and m.shape[1] == cols for m in mats)
        self.advance_ptrs(lanes, [m.ctypes.data for m in mats], [m.shape[0] for m in mats], cols,
                          offsets, max_num_frames, KD_MEM_HOST)

    def num_frames_decoded(self, lane: int) -> int:
        v = C.c_int32(0)
        _check(lib().kd_decoder_num_frames_decoded(self.h, int(lane), C.byref(v)))
        return v.value

    def reached_final(self, lane: int) -> bool:
        v = C.c_int32(0)
        _check(lib().kd_decoder_reached_final(self.h, int(lane), C.byref(v)))
        return bool(v.value)

    def best_paths(self, lanes, use_final_probs: bool = True, copy: bool = True) -> "PathBatch":
        """Best path of every lane in `lanes` (a sequence of RawPath, built on demand).
        copy=False leaves the arcs in the decoder's pinned host buffer: the result is
        only valid until the next best_paths() call on this decoder."""
        la = self._lanes(lanes)
        n = la.size
        ok = np.zeros(n, np.int32)
        rf = np.zeros(n, np.int32)
        cnt = np.zeros(n, np.int64)
        _check(lib().kd_decoder_best_path_prepare(self.h, n, la.ctypes.data, int(use_final_probs),
                                                  ok.ctypes.data, rf.ctypes.data, cnt.ctypes.data))
        off = np.zeros(n + 1, np.int64)
        np.cumsum(cnt, out=off[1:])
        total = int(off[n])
        f2 = np.zeros((n, 2), np.float32)
        ptr = [C.c_void_p() for _ in range(4)]
        _check(lib().kd_decoder_best_path_view(self.h, n, la.ctypes.data, off.ctypes.data, total,
                                               C.byref(ptr[0]), C.byref(ptr[1]), C.byref(ptr[2]),
                                               C.byref(ptr[3]), f2.ctypes.data))
        if total == 0 or not ptr[0].value:
            words = np.empty(0, np.int32)
        else:
            # the four arrays are `total` words apart in one pinned buffer
            words = np.ctypeslib.as_array(C.cast(ptr[0], C.POINTER(C.c_int32)), shape=(4 * total,))
            if copy:
                words = words.copy()
        woff = off[:n, None] + (np.arange(4, dtype=np.int64) * total)[None, :]
        return PathBatch(ok, rf, cnt, words, woff, f2, d2h_bytes=16 * total)

    # ---- asynchronous calls (kd_decoder_advance_async / _wait / _result_view)

    def advance_async(self, lanes, ptrs: Sequence[int], rows, cols: int, offsets=None,
                      max_num_frames: int = -1, mem_kind: int = KD_MEM_HOST, init: bool = False,
                      finalize: bool = False, producer_stream: int = 0) -> int:
        """Enqueues AdvanceDecoding (optionally with InitDecoding before and the best-path
        selection after it, all in one kernel launch) and returns a ticket."""
        la = self._lanes(lanes)
        pa = (C.c_void_p * la.size)(*[int(p) for p in ptrs])
        ra = np.ascontiguousarray(rows, dtype=np.int32).reshape(-1)
        oa = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.int32).reshape(-1)
        flags = (KD_ADVANCE_INIT if init else 0) | (KD_ADVANCE_FINALIZE if finalize else 0)
        t = C.c_int64(-1)
        _check(lib().kd_decoder_advance_async(
            self.h, la.size, la.ctypes.data, pa, ra.ctypes.data, int(cols),
            None if oa is None else oa.ctypes.data, int(max_num_frames), int(mem_kind), flags,
            C.c_void_p(int(producer_stream) or None), C.byref(t)))
        return t.value

    def wait(self, ticket: int = -1):
        _check(lib().kd_decoder_wait(self.h, int(ticket)))

    def results(self, ticket: int, use_final_probs: bool = True, copy: bool = True) -> "PathBatch":
        """Best paths of a finalize=True call.  copy=False returns views of the decoder's
        pinned result buffer (valid until the third-next advance_async call)."""
        n = C.c_int32(0)
        lp, wp, op = C.c_void_p(), C.c_void_p(), C.c_void_p()
        cap = self.max_lanes
        ok = np.zeros(cap, np.int32)
        rf = np.zeros(cap, np.int32)
        cnt = np.zeros(cap, np.int64)
        f2 = np.zeros((cap, 2), np.float32)
        _check(lib().kd_decoder_result_view(self.h, int(ticket), int(use_final_probs), C.byref(n),
                                            C.byref(lp), C.byref(wp), C.byref(op), cnt.ctypes.data,
                                            ok.ctypes.data, rf.ctypes.data, f2.ctypes.data))
        m = n.value
        woff = np.ctypeslib.as_array(C.cast(op, C.POINTER(C.c_int64)), shape=(m, 4))
        n_words = int((woff[:, 3] + cnt[:m]).max()) if m else 0
        words = (np.ctypeslib.as_array(C.cast(wp, C.POINTER(C.c_int32)), shape=(n_words,))
                 if n_words else np.empty(0, np.int32))
        if copy:
            words, woff = words.copy(), woff.copy()
        pb = PathBatch(ok[:m], rf[:m], cnt[:m], words, woff, f2[:m], d2h_bytes=4 * n_words)
        pb.lanes = np.ctypeslib.as_array(C.cast(lp, C.POINTER(C.c_int32)), shape=(m,)).copy()
        return pb

    def span_begin(self):
        _check(lib().kd_decoder_span_begin(self.h))

    def span_end(self) -> Tuple[float, int]:
        """(device ms from the first launch's start to the last one's end, launches)"""
        ms = C.c_float(0)
        nl = C.c_int32(0)
        _check(lib().kd_decoder_span_end(self.h, C.byref(ms), C.byref(nl)))
        return ms.value, nl.value

    def decode(self, lanes, mats: Sequence[np.ndarray], use_final_probs: bool = True) -> "PathBatch":
        """InitDecoding + AdvanceDecoding over all rows + GetBestPath of host matrices, one
        kernel launch."""
        mats = [np.ascontiguousarray(m, dtype=np.float32) for m in mats]
        self._keep = mats
        cols = mats[0].shape[1]
        t = self.advance_async(lanes, [m.ctypes.data for m in mats], [m.shape[0] for m in mats],
                               cols, None, -1, KD_MEM_HOST, init=True, finalize=True)
        self.wait(t)
        return self.results(t, use_final_probs)

    def tokens(self, lane: int) -> Tuple[np.ndarray, np.ndarray]:
        n = C.c_int64(0)
        _check(lib().kd_decoder_dump_tokens(self.h, int(lane), 0, None, None, C.byref(n)))
        st = np.empty(n.value, np.int32)
        co = np.empty(n.value, np.float64)
        if n.value:
            _check(lib().kd_decoder_dump_tokens(self.h, int(lane), n.value, st.ctypes.data,
                                                co.ctypes.data, C.byref(n)))
        # (SimpleDecoder search: the second call returns the PruneToks view, possibly shorter)
        return st[:n.value], co[:n.value]

    def final_relative_cost(self, lane: int = 0) -> float:
        v = C.c_float(0)
        _check(lib().kd_decoder_final_relative_cost(self.h, int(lane), C.byref(v)))
        return v.value

    def stats(self, lane: int = -1) -> dict:
        s = KdStats()
        _check(lib().kd_decoder_stats(self.h, int(lane), C.byref(s)))
        return s.as_dict()

    def last_advance_info(self) -> Tuple[float, int]:
        ms = C.c_float(0)
        nl = C.c_int32(0)
        _check(lib().kd_decoder_last_advance_info(self.h, C.byref(ms), C.byref(nl)))
        return ms.value, nl.value

    def info(self) -> dict:
        v = np.zeros(6, np.int64)
        _check(lib().kd_decoder_info(self.h, v.ctypes.data))
        return dict(zip(("max_lanes", "hash_capacity", "arena_records", "threads_per_lane",
                         "device_bytes", "chunk_frames"), (int(x) for x in v)))


def merge_linear(path: RawPath):
    """RemoveEpsLocal on the linear best path (kaldifst):
    greedy left-to-right merge of neighbouring arcs that do not both carry an
    ilabel nor both an olabel; a trailing (eps, eps) arc folds into the final
    weight.  Returns (ilabels, olabels, graph, acoustic, final2)."""
    il, ol, gw, aw = [], [], [], []
    f = np.array(path.final, dtype=np.float32)
    n = len(path.ilabels)
    if n == 0:
        e = np.empty(0, np.int32)
        return e, e.copy(), np.empty(0, np.float32), np.empty(0, np.float32), f
    ci, co = int(path.ilabels[0]), int(path.olabels[0])
    cg, ca = np.float32(path.graph[0]), np.float32(path.acoustic[0])
    for k in range(1, n):
        ni, no = int(path.ilabels[k]), int(path.olabels[k])
        if not (ci != 0 and ni != 0) and not (co != 0 and no != 0):
            ci = ci if ci != 0 else ni
            co = co if co != 0 else no
            cg = np.float32(cg + np.float32(path.graph[k]))
            ca = np.float32(ca + np.float32(path.acoustic[k]))
        else:
            il.append(ci); ol.append(co); gw.append(cg); aw.append(ca)
            ci, co = ni, no
            cg, ca = np.float32(path.graph[k]), np.float32(path.acoustic[k])
    if ci == 0 and co == 0:
        f = np.array([np.float32(cg + f[0]), np.float32(ca + f[1])], dtype=np.float32)
    else:
        il.append(ci); ol.append(co); gw.append(cg); aw.append(ca)
    return (np.asarray(il, np.int32), np.asarray(ol, np.int32), np.asarray(gw, np.float32),
            np.asarray(aw, np.float32), f)
