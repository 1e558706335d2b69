"""CUDA path through the C ABI and the drop-in Python API -- needs a B200."""
import numpy as np
import pytest

from common import GOLDEN_CASES, GoldenCase, rel_close, same_labels, small_graph, sorted_tokens
from kaldi_decoder_b200 import capi, synth
from oracle import kd_oracle, kd_ref

pytestmark = pytest.mark.gpu


def _decoder(g, opts, lanes, **kw):
    dg = capi.DeviceGraph.from_graph(g)
    kw.setdefault("hash_capacity", 1 << 15)
    kw.setdefault("arena_records", 1 << 20)
    return capi.LaneDecoder(dg, capi.make_options(**opts), max_lanes=lanes, **kw)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_vectors_from_the_reference(name):
    """Golden fixtures frozen from the unmodified reference: identical label sequences and
    total cost (1e-4 relative) wherever the reference's pruning is order independent; on the
    max_active-binding fixtures a differing utterance must be an exact cost tie or a binding
    utterance."""
    gc = GoldenCase(name)
    dec = _decoder(gc.graph, gc.opts, gc.n_utts)
    lanes = list(range(gc.n_utts))
    dec.init(lanes)
    dec.advance(lanes, [gc.logp(u) for u in lanes])
    og = kd_oracle.OracleGraph(gc.graph)
    for ufp in (True, False):
        paths = dec.best_paths(lanes, ufp)
        for u in lanes:
            want = gc.best(u, ufp)
            il, ol, gw, aw, fin = capi.merge_linear(paths[u])
            assert paths[u].ok == want.ok
            assert paths[u].reached_final == gc.reached_final(u)
            o = kd_oracle.OracleDecoder(og, kd_ref.Options(**gc.opts), kd_oracle.REFERENCE_ORDER)
            o.decode(gc.logp(u))
            binding = o.stats()["binding_max"] > 0
            identical = np.array_equal(il[il != 0], want.isyms) and \
                np.array_equal(ol[ol != 0], want.osyms)
            if not binding:
                assert identical, (name, u, ufp)
                assert rel_close(paths[u].total_cost, want.total_cost, 1e-4)
                # RemoveEpsLocal grouping: same arcs, same per-arc weights
                assert np.array_equal(il, want.ilabels) and np.array_equal(ol, want.olabels)
                assert np.allclose(gw, want.graph, rtol=0, atol=1e-5)
                assert np.allclose(aw, want.acoustic, rtol=0, atol=1e-3)
                assert np.allclose(fin, want.final, rtol=0, atol=1e-6)
            elif not identical:
                # order-dependent pruning (the reference's result depends on its HashList
                # order once max_active binds): the device must then equal the
                # order-independent statement of the search exactly
                c = kd_oracle.OracleDecoder(og, kd_ref.Options(**gc.opts), kd_oracle.CANONICAL)
                c.decode(gc.logp(u))
                cb = c.get_best_path(ufp, raw=True)
                assert same_labels(paths[u], cb), (name, u, ufp)
                assert rel_close(paths[u].total_cost, cb.total_cost, 1e-6)


def test_identical_to_reference_on_peaky_hlg():
    """Bigger differential test against oracle/_ref (or the oracle's reference-order mode when
    the compiled reference is absent): 32 utterances, T=300."""
    g = synth.make_hlg(5000, (200, 400), 100, seed=21)
    opts = dict(beam=20.0, max_active=7000, min_active=20)
    n, T = 32, 300
    mats = [synth.make_logprobs(g, T, seed=900 + u, peak=12) for u in range(n)]
    dec = _decoder(g, opts, n, hash_capacity=1 << 17, arena_records=1 << 21)
    lanes = list(range(n))
    dec.init(lanes)
    dec.advance(lanes, mats)
    paths = dec.best_paths(lanes, True)
    if kd_ref.available():
        rg = kd_ref.RefGraph(g)
        _, ref_paths, rf = kd_ref.decode_batch(rg, np.stack(mats), kd_ref.Options(**opts), 8)
    else:
        og = kd_oracle.OracleGraph(g)
        _, ref_paths, rf, _, _ = kd_oracle.decode_batch(og, np.stack(mats), kd_ref.Options(**opts), 8,
                                                         kd_oracle.REFERENCE_ORDER)
    diverged = 0
    for u in lanes:
        assert paths[u].reached_final == bool(rf[u])
        if not same_labels(paths[u], ref_paths[u]):
            diverged += 1
            # a divergence is only acceptable as an exact tie in total cost
            assert paths[u].total_cost == pytest.approx(ref_paths[u].total_cost, rel=1e-6)
        assert rel_close(paths[u].total_cost, ref_paths[u].total_cost, 1e-4)
    assert diverged == 0


def test_streaming_chunks_equal_one_shot():
    g = small_graph("HLG")
    opts = dict(beam=16.0, max_active=2**31 - 1, min_active=20)
    T = 120
    lp = synth.make_logprobs(g, T, seed=4, peak=6)
    dec = _decoder(g, opts, 2)
    dec.init([0, 1])
    dec.advance([0], [lp])
    for a in range(0, T, 40):
        chunk = lp[a:a + 40]
        dec.advance([1], [chunk], offsets=[a], max_num_frames=15)
        assert dec.num_frames_decoded(1) == a + 15
        dec.advance([1], [chunk], offsets=[a])
        assert dec.num_frames_decoded(1) == a + 40
        # partial result is available at any time (GetBestPath mid-utterance)
        assert dec.best_paths([1], True)[0].ok
    s0, c0 = sorted_tokens(*dec.tokens(0))
    s1, c1 = sorted_tokens(*dec.tokens(1))
    assert np.array_equal(s0, s1) and np.array_equal(c0, c1)
    p0, p1 = dec.best_paths([0, 1], True)
    assert np.array_equal(p0.ilabels, p1.ilabels) and np.array_equal(p0.acoustic, p1.acoustic)


def test_lanes_are_independent_and_reusable():
    """The same utterance gives the same result in any lane, alone or in a batch, and a lane
    can be re-initialised for a new utterance."""
    g = small_graph("HL")
    opts = dict(beam=20.0, max_active=7000, min_active=20)
    mats = [synth.make_logprobs(g, 80, seed=u, peak=8) for u in range(5)]
    dec = _decoder(g, opts, 8)
    dec.init([0, 1, 2, 3, 4])
    dec.advance([0, 1, 2, 3, 4], mats)
    batch = dec.best_paths([0, 1, 2, 3, 4])
    dec.init([7, 5])
    dec.advance([7, 5], [mats[3], mats[1]])
    again = dec.best_paths([7, 5])
    for a, b in ((again[0], batch[3]), (again[1], batch[1])):
        assert np.array_equal(a.ilabels, b.ilabels) and np.array_equal(a.olabels, b.olabels)
        assert np.array_equal(a.graph, b.graph) and np.array_equal(a.acoustic, b.acoustic)
    # ragged lengths in one call, zero-length included
    dec.init([0, 1, 2])
    dec.advance([0, 1, 2], [mats[0][:17], mats[1][:0], mats[2]])
    assert [dec.num_frames_decoded(i) for i in (0, 1, 2)] == [17, 0, 80]
    assert dec.best_paths([1])[0].ok and len(dec.best_paths([1])[0].ilabels) == 0


def test_error_behaviour_matches_reference_assertions():
    g = small_graph("H")
    dg = capi.DeviceGraph.from_graph(g)
    lp = synth.make_logprobs(g, 10, seed=1)
    # faster-decoder.cc:24-28
    for bad in (dict(hash_ratio=0.5), dict(max_active=1), dict(min_active=5, max_active=5)):
        with pytest.raises(capi.KdError, match="Check failed"):
            capi.LaneDecoder(dg, capi.make_options(**bad))
    dec = capi.LaneDecoder(dg, capi.make_options(), max_lanes=2, hash_capacity=4096,
                           arena_records=1 << 16)
    # faster-decoder.cc:128-129
    with pytest.raises(capi.KdError, match="InitDecoding"):
        dec.advance([0], [lp])
    dec.init([0])
    dec.advance([0], [lp])
    # faster-decoder.cc:137: fewer frames ready than decoded
    with pytest.raises(capi.KdError, match="num_frames_ready >= num_frames_decoded_"):
        dec.advance([0], [lp[:4]])
    # ilabel beyond the decodable's columns (the reference reads out of bounds)
    with pytest.raises(capi.KdError, match="columns"):
        dec.advance([0], [np.zeros((12, 10), np.float32)])
    with pytest.raises(capi.KdError, match="lane id"):
        dec.init([2])
    # step 2 of GetBestPath without step 1 since the last advance
    dec.best_paths([0])
    dec.advance([0], [np.vstack([lp, lp[:3]])])
    la = np.zeros(1, np.int32)
    off = np.zeros(1, np.int64)
    buf = np.zeros(64, np.int32)
    with pytest.raises(capi.KdError, match="best_path_prepare"):
        capi._check(capi.lib().kd_decoder_best_path_fetch(
            dec.h, 1, la.ctypes.data, off.ctypes.data, 64, buf.ctypes.data, buf.ctypes.data,
            buf.ctypes.data, buf.ctypes.data, None))
    assert len(dec.best_paths([0])[0].ilabels) == 13
    # a graph without start state (faster-decoder.cc:47)
    with pytest.raises(capi.KdError, match="kNoStateId"):
        capi.DeviceGraph(g.num_states, -1, g.row_off, g.ilabel, g.olabel, g.weight, g.nextstate,
                         g.final)


def test_capacity_overflow_is_a_hard_error_and_lane_recovers():
    g = small_graph("HLG")
    dg = capi.DeviceGraph.from_graph(g)
    opts = capi.make_options(beam=30.0, max_active=2**31 - 1, min_active=0)
    lp = synth.make_logprobs(g, 60, seed=2, peak=2)  # flat posteriors: thousands of tokens
    dec = capi.LaneDecoder(dg, opts, max_lanes=1, hash_capacity=256, arena_records=1 << 20)
    dec.init([0])
    with pytest.raises(capi.KdError, match="overflow"):
        dec.advance([0], [lp])
    with pytest.raises(capi.KdError, match="error state"):
        dec.advance([0], [lp])
    small = capi.LaneDecoder(dg, opts, max_lanes=1, hash_capacity=1 << 16, arena_records=2000)
    small.init([0])
    with pytest.raises(capi.KdError, match="arena"):
        small.advance([0], [lp])
    # re-init clears the error; a narrow search then fits
    dec.set_options(capi.make_options(beam=4.0, max_active=40, min_active=0))
    dec.init([0])
    dec.advance([0], [synth.make_logprobs(g, 30, seed=3, peak=14)])
    assert dec.best_paths([0])[0].ok


def test_degenerate_inputs():
    """-inf log-probs drop arcs (inf < inf is false, faster-decoder.cc:211); when nothing
    survives GetBestPath returns false (faster-decoder.cc:386-389)."""
    g = small_graph("H")
    opts = dict(beam=16.0, max_active=2**31 - 1, min_active=20)
    dec = _decoder(g, opts, 2)
    lp = synth.make_logprobs(g, 20, seed=5, peak=6)
    dead = lp.copy()
    dead[10, :] = -np.inf
    dec.init([0, 1])
    dec.advance([0, 1], [lp, dead])
    p = dec.best_paths([0, 1])
    assert p[0].ok and not p[1].ok and len(p[1].ilabels) == 0
    assert dec.tokens(1)[0].size == 0 and not dec.reached_final(1)
    og = kd_oracle.OracleGraph(g)
    o = kd_oracle.OracleDecoder(og, kd_ref.Options(**opts), kd_oracle.REFERENCE_ORDER)
    o.decode(dead)
    assert not o.get_best_path().ok


def test_drop_in_python_api_end_to_end():
    """kaldi_decoder.FasterDecoder used as the icefall scripts use it, against the reference
    result frozen in the golden fixture."""
    import kaldi_decoder as kd
    gc = GoldenCase("hlg300_peaky")
    g = gc.graph
    fst = kd.StdVectorFst.from_arrays(g.num_states, g.start, g.row_off, g.ilabel, g.olabel,
                                      g.weight, g.nextstate, g.final)
    o = gc.opts
    decoder = kd.FasterDecoder(fst, kd.FasterDecoderOptions(beam=o["beam"], max_active=o["max_active"],
                                                            min_active=o["min_active"]))
    for u in range(gc.n_utts):
        decodable = kd.DecodableCtc(gc.logp(u))
        decoder.decode(decodable)
        assert decoder.num_frames_decoded() == gc.T
        assert decoder.reached_final() == gc.reached_final(u)
        ok, best = decoder.get_best_path()
        want = gc.best(u, True)
        assert ok == want.ok
        ok2, isyms, osyms, (gcost, acost) = kd.get_linear_symbol_sequence(best)
        assert ok2 and isyms == [int(x) for x in want.isyms] and osyms == [int(x) for x in want.osyms]
        assert rel_close(gcost + acost, want.total_cost, 1e-4)
        assert best.num_states == len(want.ilabels) + 1  # same RemoveEpsLocal grouping
    # streaming with offsets + a Python-defined decodable
    lp = gc.logp(0)

    class PyDecodable(kd.DecodableInterface):
        def __init__(self, m):
            super().__init__()
            self.m = m

        def log_likelihood(self, frame, index):
            return float(self.m[frame, index - 1])

        def is_last_frame(self, frame):
            return frame == self.m.shape[0] - 1

        def num_frames_ready(self):
            return self.m.shape[0]

        def num_indices(self):
            return self.m.shape[1]

    decoder.init_decoding()
    decoder.advance_decoding(kd.DecodableCtc(lp[:30]), 10)
    assert decoder.num_frames_decoded() == 10
    decoder.advance_decoding(kd.DecodableCtc(lp[:30]))
    decoder.advance_decoding(kd.DecodableCtc(lp[30:], offset=30))
    ok, a = decoder.get_best_path()
    decoder.decode(PyDecodable(lp))
    ok, b = decoder.get_best_path()
    assert kd.get_linear_symbol_sequence(a)[1:3] == kd.get_linear_symbol_sequence(b)[1:3]
    with pytest.raises(RuntimeError):
        kd.FasterDecoder(fst, kd.FasterDecoderOptions()).advance_decoding(kd.DecodableCtc(lp))
    # batched additive API
    bd = kd.BatchFasterDecoder(fst, kd.FasterDecoderOptions(beam=o["beam"], max_active=o["max_active"]),
                               max_lanes=4)
    bd.init_decoding([0, 1, 2])
    bd.advance_decoding([0, 1, 2], [gc.logp(0), gc.logp(1), gc.logp(2)])
    oks, lats = bd.get_best_paths([0, 1, 2])
    for u in range(3):
        assert oks[u]
        assert kd.get_linear_symbol_sequence(lats[u])[2] == [int(x) for x in gc.best(u).osyms]


def test_full_size_properties_on_the_bench_graph():
    """BASELINE-size graph (HLG 3-gram, ~4.6M arcs), T=1000: determinism, lane-position
    invariance and self-consistency of the returned costs (size-independent properties)."""
    g = synth.make_config_graph("C3")
    opts = dict(beam=20.0, max_active=7000, min_active=20)
    n, T = 48, 1000
    mats = [synth.make_logprobs(g, T, seed=7000 + u, peak=12) for u in range(n // 2)]
    mats = mats + mats[::-1]  # every utterance appears in two different lanes
    dec = _decoder(g, opts, n, hash_capacity=1 << 17, arena_records=1 << 22)
    lanes = list(range(n))
    dec.init(lanes)
    dec.advance(lanes, mats)
    first = dec.best_paths(lanes)
    st1 = dec.stats()
    dec.init(lanes)
    dec.advance(lanes, mats)
    second = dec.best_paths(lanes)
    assert dec.stats()["emit_arcs"] == st1["emit_arcs"] and st1["frames"] == n * T
    for u in lanes:
        a, b, c = first[u], second[u], first[n - 1 - u]
        for x in (b, c):
            assert np.array_equal(a.ilabels, x.ilabels) and np.array_equal(a.olabels, x.olabels)
            assert np.array_equal(a.graph, x.graph) and np.array_equal(a.acoustic, x.acoustic)
        assert a.ok and (a.ilabels != 0).sum() == T  # one emitting arc per frame
        # acoustic cost of the path == -sum of the log-probs it consumed (fp32 rounding per arc)
        emit = a.ilabels != 0
        lp = mats[u][np.arange(T), a.ilabels[emit] - 1].astype(np.float64)
        assert abs(a.acoustic.astype(np.float64).sum() + lp.sum()) < 1e-2


def test_path_batch_views_and_copies():
    """best_paths(copy=False) returns views of the decoder's pinned buffer (valid until the
    next call); copy=True owns its arrays; PathBatch behaves like a sequence."""
    g = small_graph("HLG")
    opts = dict(beam=12.0, max_active=300, min_active=20)
    dec = _decoder(g, opts, 3)
    mats = [synth.make_logprobs(g, 40 + 10 * u, seed=60 + u, peak=7) for u in range(3)]
    dec.init([0, 1, 2])
    dec.advance([0, 1, 2], mats)
    owned = dec.best_paths([0, 1, 2])
    views = dec.best_paths([0, 1, 2], copy=False)
    assert len(owned) == len(views) == 3
    for a, b in zip(owned, views):
        assert a.ok and b.ok
        assert np.array_equal(a.ilabels, b.ilabels) and np.array_equal(a.graph, b.graph)
    assert [len(p.ilabels) for p in owned[1:]] == [len(owned[1].ilabels), len(owned[2].ilabels)]
    assert np.array_equal(owned[-1].olabels, owned[2].olabels)
    with pytest.raises(IndexError):
        owned[3]
    keep = [p.ilabels.copy() for p in views]
    # another call reuses the pinned buffer: the owned copy is unaffected
    dec.best_paths([2])
    for a, k in zip(owned, keep):
        assert np.array_equal(a.ilabels, k)


def test_cuda_tensors_are_decoded_in_place():
    """kaldi_decoder.advance_decoding_cuda: torch CUDA tensors, no host copy; same result as
    the numpy path."""
    import torch
    import kaldi_decoder as kd
    g = small_graph("HLG")
    fst = kd.StdVectorFst.from_arrays(g.num_states, g.start, g.row_off, g.ilabel, g.olabel,
                                      g.weight, g.nextstate, g.final)
    bd = kd.BatchFasterDecoder(fst, kd.FasterDecoderOptions(beam=12.0, max_active=300), 4)
    mats = [synth.make_logprobs(g, 50 + 7 * u, seed=90 + u, peak=7) for u in range(2)]
    bd.init_decoding([0, 1, 2, 3])
    bd.advance_decoding([0, 1], mats)
    kd.advance_decoding_cuda(bd, [2, 3], [torch.from_numpy(m).cuda() for m in mats])
    oks, lats = bd.get_best_paths([0, 1, 2, 3])
    assert all(oks)
    seqs = [kd.get_linear_symbol_sequence(l) for l in lats]
    for u in range(2):
        assert list(seqs[u][1]) == list(seqs[u + 2][1]) and list(seqs[u][2]) == list(seqs[u + 2][2])
        assert seqs[u][3] == seqs[u + 2][3]
    with pytest.raises(ValueError):
        kd.advance_decoding_cuda(bd, [0], [torch.zeros(3, 4)])


def test_const_fst_overload_and_deferred_batches_through_the_python_package():
    """The reference's ConstFst / Fst constructor overloads (python/csrc/faster-decoder.cc:
    34-42) and the additive deferred API of BatchFasterDecoder, against the golden fixture."""
    import kaldi_decoder as kd
    gc = GoldenCase("hlg300_peaky")
    g = gc.graph
    o = gc.opts
    vfst = kd.StdVectorFst.from_arrays(g.num_states, g.start, g.row_off, g.ilabel, g.olabel,
                                       g.weight, g.nextstate, g.final)
    cfst = kd.StdConstFst(vfst)
    opts = kd.FasterDecoderOptions(beam=o["beam"], max_active=o["max_active"], min_active=o["min_active"])
    for f in (cfst, vfst):
        dec = kd.FasterDecoder(f, opts)
        dec.decode(kd.DecodableCtc(gc.logp(0)))
        ok, best = dec.get_best_path()
        assert ok and dec.reached_final() == gc.reached_final(0)
        assert kd.get_linear_symbol_sequence(best)[2] == [int(x) for x in gc.best(0).osyms]
    # deferred batches: two calls in flight on disjoint lanes
    n = min(4, gc.n_utts)
    bd = kd.BatchFasterDecoder(cfst, opts, max_lanes=2 * n)
    mats = [np.ascontiguousarray(gc.logp(u), dtype=np.float32) for u in range(n)]
    ptrs = [m.ctypes.data for m in mats]
    rows = [m.shape[0] for m in mats]
    cols = mats[0].shape[1]
    t0 = bd.decode_async(list(range(n)), ptrs, rows, cols)
    t1 = bd.decode_async(list(range(n, 2 * n)), ptrs, rows, cols)
    for t, base in ((t0, 0), (t1, n)):
        lanes, oks, lats = bd.get_results(t)
        assert lanes == list(range(base, base + n))
        for u in range(n):
            assert oks[u]
            ok2, isyms, osyms, (gcost, acost) = kd.get_linear_symbol_sequence(lats[u])
            want = gc.best(u, True)
            assert ok2 and isyms == [int(x) for x in want.isyms] and osyms == [int(x) for x in want.osyms]
            assert rel_close(gcost + acost, want.total_cost, 1e-4)
            assert bd.reached_final(base + u) == gc.reached_final(u)
    bd.wait()
    # CUDA tensors, ordered behind the stream that produces them
    torch = pytest.importorskip("torch")
    dev = [torch.from_numpy(m).cuda() for m in mats]
    t = kd.decode_cuda_async(bd, list(range(n)), dev)
    lanes, oks, lats = bd.get_results(t)
    for u in range(n):
        assert kd.get_linear_symbol_sequence(lats[u])[2] == [int(x) for x in gc.best(u).osyms]
    kd.advance_decoding_cuda(bd, [0], [dev[0][:0]], offsets=[rows[0]], device=0)  # nothing new: a no-op

    class OnlyDLPack:  # a foreign CUDA array that speaks DLPack and nothing else
        def __init__(self, t):
            self.t = t

        def __dlpack__(self, stream=None):
            return self.t.__dlpack__()

        def __dlpack_device__(self):
            return self.t.__dlpack_device__()

    wrapped = [OnlyDLPack(x) for x in dev]
    torch.cuda.synchronize()
    t = kd.decode_cuda_async(bd, list(range(n)), wrapped)
    lanes, oks, lats = bd.get_results(t)
    for u in range(n):
        assert kd.get_linear_symbol_sequence(lats[u])[2] == [int(x) for x in gc.best(u).osyms]
    with pytest.raises(ValueError):
        kd.decode_cuda_async(bd, [0], [OnlyDLPack(dev[0].double())])
    with pytest.raises(ValueError):
        kd.advance_decoding_cuda(bd, [0], [dev[0]], device=1)


def test_single_utterance_decoders_run_concurrently_from_python_threads():
    """One FasterDecoder per thread on a shared DeviceGraph: the bindings release the GIL while
    the device works, every decoder owns its lane and streams; results are the reference's."""
    import threading
    import kaldi_decoder as kd
    gc = GoldenCase("hlg300_peaky")
    g = gc.graph
    o = gc.opts
    fst = kd.StdConstFst.from_arrays(g.num_states, g.start, g.row_off, g.ilabel, g.olabel,
                                     g.weight, g.nextstate, g.final)
    graph = kd.DeviceGraph(fst)
    opts = kd.FasterDecoderOptions(beam=o["beam"], max_active=o["max_active"], min_active=o["min_active"])
    n_threads = 4
    errors = []
    # (the .npz reader is not thread-safe: everything is read before the threads start)
    logps = [gc.logp(u) for u in range(gc.n_utts)]
    wants = [(gc.best(u, True), gc.reached_final(u)) for u in range(gc.n_utts)]

    def work(k):
        try:
            dec = kd.FasterDecoder(graph, opts)
            for rep in range(6):
                u = (k + rep) % gc.n_utts
                dec.decode(kd.DecodableCtc(logps[u]))
                ok, best = dec.get_best_path()
                want, rf = wants[u]
                assert ok == want.ok and dec.reached_final() == rf
                assert kd.get_linear_symbol_sequence(best)[2] == [int(x) for x in want.osyms]
        except Exception as e:  # noqa: BLE001
            errors.append((k, repr(e)))

    threads = [threading.Thread(target=work, args=(k,)) for k in range(n_threads)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_a_decoder_per_utterance_shares_the_graph_and_recycles_the_device_side():
    """The reference's scripts construct `FasterDecoder(HLG, opts)` for every utterance
    (its constructor only stores a reference, faster-decoder.cc:21-32).  Here decoders built from
    the same FST content share ONE device graph, and a decoder that goes away hands its device
    buffers to the next one -- which must behave exactly like a new decoder."""
    import kaldi_decoder as kd
    gc = GoldenCase("hlg300_peaky")
    g = gc.graph
    o = gc.opts
    vfst = kd.StdVectorFst.from_arrays(g.num_states, g.start, g.row_off, g.ilabel, g.olabel,
                                       g.weight, g.nextstate, g.final)
    opts = kd.FasterDecoderOptions(beam=o["beam"], max_active=o["max_active"], min_active=o["min_active"])
    kd.clear_graph_cache()
    before = kd.graph_uploads()
    for u in range(min(4, gc.n_utts)):
        dec = kd.FasterDecoder(vfst, opts)       # as icefall's decode(): one object per utterance
        assert dec.num_frames_decoded() == -1    # a recycled device side starts uninitialised ...
        with pytest.raises(RuntimeError):        # ... and says so (faster-decoder.cc:128-129)
            dec.advance_decoding(kd.DecodableCtc(gc.logp(u)))
        dec.decode(kd.DecodableCtc(gc.logp(u)))
        want = gc.best(u, True)
        assert dec.reached_final() == gc.reached_final(u)
        ok, best = dec.get_best_path()
        assert ok == want.ok
        assert kd.get_linear_symbol_sequence(best)[2] == [int(x) for x in want.osyms]
        del dec
    assert kd.graph_uploads() == before + 1
    # copies and conversions hold the same content: no new upload
    dec = kd.FasterDecoder(kd.StdConstFst(vfst), opts)
    assert kd.graph_uploads() == before + 1
    # invalid options on a recycled decoder raise as on a new one, and do not lose it
    del dec
    with pytest.raises(RuntimeError):
        kd.FasterDecoder(vfst, kd.FasterDecoderOptions(max_active=1))
    # a modified FST is a different graph
    edited = kd.StdVectorFst(vfst)
    s_new = edited.add_state()
    edited.add_arc(s_new, 1, 1, 0.25, s_new)
    assert edited.content_id != vfst.content_id
    dec = kd.FasterDecoder(edited, opts)
    assert kd.graph_uploads() == before + 2
    dec.decode(kd.DecodableCtc(gc.logp(0)))      # (the new state is unreachable: same answer)
    assert kd.get_linear_symbol_sequence(dec.get_best_path()[1])[2] == [int(x) for x in gc.best(0).osyms]
    kd.clear_graph_cache()
    dec2 = kd.FasterDecoder(vfst, opts)
    assert kd.graph_uploads() == before + 3


def test_cpp_caller_written_against_the_reference_headers_runs(tmp_path):
    """The C++ drop-in caller (tests/cpp/drop_in.cc: the reference's include paths, Decode /
    InitDecoding + AdvanceDecoding / GetBestPath / ReachedFinal / NumFramesDecoded) against
    the reference's results frozen in the golden fixtures."""
    import subprocess
    import kaldi_decoder as kd
    from common import build_cpp_drop_in
    exe = build_cpp_drop_in(tmp_path)
    if exe is None:
        pytest.skip("no g++")
    for name in ("hlg300_peaky", "hl300"):
        gc = GoldenCase(name)
        g = gc.graph
        fst = kd.StdVectorFst.from_arrays(g.num_states, g.start, g.row_off, g.ilabel, g.olabel,
                                          g.weight, g.nextstate, g.final)
        fst.write(str(tmp_path / "g.fst"))
        lp = np.stack([gc.logp(u) for u in range(gc.n_utts)]).astype(np.float32)
        lp.tofile(str(tmp_path / "lp.bin"))
        o = gc.opts
        r = subprocess.run([exe, str(tmp_path / "g.fst"), str(o["beam"]), str(o["max_active"]),
                            str(o["min_active"]), str(tmp_path / "lp.bin"), str(gc.n_utts),
                            str(gc.T), str(lp.shape[2])], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        lines = r.stdout.strip().splitlines()
        assert len(lines) == gc.n_utts
        for u, ln in enumerate(lines):
            f = ln.split()
            want = gc.best(u, True)
            assert int(f[0]) == u and int(f[1]) == int(want.ok), (name, ln)
            assert int(f[2]) == int(gc.reached_final(u)) and int(f[3]) == gc.T, (name, ln)
            assert [int(x) for x in f[5:]] == [int(x) for x in want.osyms], (name, u)
            assert rel_close(float(f[4]), want.total_cost, 1e-4), (name, u, f[4], want.total_cost)
