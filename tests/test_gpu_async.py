"""Asynchronous / fused calls of the C ABI (kd_decoder_advance_async, _wait, _result_view),
the host-memory advance under serialised launches, and the deterministic epsilon tie-break
-- needs a B200."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from common import rel_close, small_graph, sorted_tokens
from kaldi_decoder_b200 import capi, synth
from oracle import kd_oracle, kd_ref

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPTS = dict(beam=14.0, max_active=400, min_active=20)


def _oracle_paths(g, mats, opts):
    og = kd_oracle.OracleGraph(g)
    out = []
    for m in mats:
        o = kd_oracle.OracleDecoder(og, kd_ref.Options(**opts), kd_oracle.CANONICAL)
        o.decode(m)
        out.append((o.get_best_path(True, raw=True), sorted_tokens(*o.tokens()), o.reached_final()))
    return out


def _same_raw(p, ob):
    return (p.ok == ob.ok and np.array_equal(p.ilabels, ob.ilabels)
            and np.array_equal(p.olabels, ob.olabels) and np.array_equal(p.graph, ob.graph)
            and np.array_equal(p.acoustic, ob.acoustic))


_CHILD = textwrap.dedent("""
    import sys, numpy as np
    sys.path[:0] = [{root!r}, {root!r} + "/kaldi-decoder_b200/python", {root!r} + "/tests"]
    from common import small_graph, sorted_tokens
    from kaldi_decoder_b200 import capi, synth
    from oracle import kd_oracle, kd_ref
    opts = dict(beam=14.0, max_active=400, min_active=20)
    g = small_graph("HLG")
    uniform = {uniform}
    n = 6
    Ts = [200] * n if uniform else [200, 150, 97, 200, 31, 180]
    if uniform:
        block = np.stack([synth.make_logprobs(g, 200, seed=40 + u, peak=6) for u in range(n)])
        mats = [block[u] for u in range(n)]      # equally spaced: one 2-D copy per chunk
    else:
        mats = [synth.make_logprobs(g, Ts[u], seed=40 + u, peak=6) for u in range(n)]
    dg = capi.DeviceGraph.from_graph(g)
    dec = capi.LaneDecoder(dg, capi.make_options(**opts), max_lanes=n, hash_capacity=1 << 14,
                           arena_records=1 << 19, chunk_frames=32)
    lanes = list(range(n))
    dec.init(lanes)
    dec.advance(lanes, mats)
    launches = dec.last_advance_info()[1]
    paths = dec.best_paths(lanes, True)
    og = kd_oracle.OracleGraph(g)
    for u in lanes:
        o = kd_oracle.OracleDecoder(og, kd_ref.Options(**opts), kd_oracle.CANONICAL)
        o.decode(mats[u])
        gs, gc = sorted_tokens(*dec.tokens(u))
        os_, oc = sorted_tokens(*o.tokens())
        assert dec.num_frames_decoded(u) == Ts[u]
        assert np.array_equal(gs, os_) and np.array_equal(gc, oc), u
        ob = o.get_best_path(True, raw=True)
        assert np.array_equal(paths[u].olabels, ob.olabels), u
    print("CHILD_OK launches=%d" % launches)
""")


@pytest.mark.parametrize("uniform", [True, False])
@pytest.mark.parametrize("early_launch", [False, True])
def test_host_advance_under_blocking_launches(uniform, early_launch):
    """CUDA_LAUNCH_BLOCKING=1 (the same holds under ncu / compute-sanitizer): the copies of a
    host-memory advance are enqueued before the search kernel, so the kernel finds its rows.
    With the test knob that launches after the first copy, lanes run dry, yield, and the
    host launches again -- the result is the same."""
    env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1")
    if early_launch:
        env["KD_B200_COPIES_BEFORE_LAUNCH"] = "1"
    r = subprocess.run([sys.executable, "-c", _CHILD.format(root=ROOT, uniform=uniform)], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "CHILD_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    launches = int(r.stdout.split("launches=")[1].split()[0])
    assert launches >= (2 if early_launch else 1), r.stdout
    if not early_launch:
        assert launches == 1, r.stdout


def test_fused_decode_is_one_launch_and_equals_the_stepwise_calls():
    g = small_graph("HLG")
    n, T = 8, 120
    mats = [synth.make_logprobs(g, T - 7 * u, seed=300 + u, peak=7) for u in range(n)]
    want = _oracle_paths(g, mats, OPTS)
    dg = capi.DeviceGraph.from_graph(g)
    dec = capi.LaneDecoder(dg, capi.make_options(**OPTS), max_lanes=n, hash_capacity=1 << 14,
                           arena_records=1 << 19)
    lanes = list(range(n))
    for use_final in (True, False):
        pb = dec.decode(lanes, mats, use_final)
        assert dec.last_advance_info()[1] == 1
        assert list(pb.lanes) == lanes
        for u in lanes:
            ob, (os_, oc), rf = want[u]
            if not use_final:
                og = kd_oracle.OracleGraph(g)
                o = kd_oracle.OracleDecoder(og, kd_ref.Options(**OPTS), kd_oracle.CANONICAL)
                o.decode(mats[u])
                ob = o.get_best_path(False, raw=True)
            assert _same_raw(pb[u], ob), u
            assert np.array_equal(pb[u].final, ob.final)
            assert pb[u].reached_final == rf == dec.reached_final(u)
            gs, gc = sorted_tokens(*dec.tokens(u))
            assert np.array_equal(gs, os_) and np.array_equal(gc, oc)
            assert dec.num_frames_decoded(u) == mats[u].shape[0]
        # the step-wise API sees the same selection without launching it again
        again = dec.best_paths(lanes, use_final)
        for u in lanes:
            assert _same_raw(again[u], pb[u])


@pytest.mark.parametrize("region", ["1,64", "2,256", "0"])
def test_a_frame_whose_table_region_fills_up_is_searched_again(region, monkeypatch):
    """A frame hashes into a region of its lane's table sized from its candidate count; the
    epsilon closure can outgrow it, then the frame is rolled back and searched again with the
    whole table.  Tiny regions (KD_B200_TABLE_REGION) make that the common case: tokens after
    every frame, best paths and ReachedFinal must still be the oracle's, fused and streaming."""
    monkeypatch.setenv("KD_B200_TABLE_REGION", region)
    g = small_graph("HLG")
    n, T = 8, 90
    mats = [synth.make_logprobs(g, T - 5 * u, seed=500 + u, peak=7) for u in range(n)]
    want = _oracle_paths(g, mats, OPTS)
    dg = capi.DeviceGraph.from_graph(g)
    dec = capi.LaneDecoder(dg, capi.make_options(**OPTS), max_lanes=n, hash_capacity=1 << 14,
                           arena_records=1 << 19)
    lanes = list(range(n))
    pb = dec.decode(lanes, mats, True)
    retries = dec.stats()["table_retries"]
    if region == "1,64":
        assert retries > 0
    if region == "0":
        assert retries == 0
    for u in lanes:
        ob, (os_, oc), rf = want[u]
        assert _same_raw(pb[u], ob), u
        assert pb[u].reached_final == rf
        gs, gc = sorted_tokens(*dec.tokens(u))
        assert np.array_equal(gs, os_) and np.array_equal(gc, oc)
    # frame by frame from host memory (the row of a frame that starts over is loaded again;
    # the prefetched next row must not be lost), tokens checked after every frame
    og = kd_oracle.OracleGraph(g)
    o = kd_oracle.OracleDecoder(og, kd_ref.Options(**OPTS), kd_oracle.CANONICAL)
    o.init_decoding()
    dec.init([0])
    m = mats[0]
    for f in range(0, m.shape[0], 3):
        k = min(3, m.shape[0] - f)
        o.advance_decoding(m, 0, k)
        dec.advance([0], [m], max_num_frames=k)
        gs, gc = sorted_tokens(*dec.tokens(0))
        os_, oc = sorted_tokens(*o.tokens())
        assert np.array_equal(gs, os_) and np.array_equal(gc, oc), f
    assert _same_raw(dec.best_paths([0], True)[0], o.get_best_path(True, raw=True))


def test_two_lane_groups_in_flight_and_deferred_sync():
    """Two calls in flight on disjoint lanes; touching a lane completes the call that owns it."""
    g = small_graph("HL")
    n, T = 12, 90
    mats = [synth.make_logprobs(g, T, seed=500 + u, peak=6) for u in range(n)]
    want = _oracle_paths(g, mats, OPTS)
    dg = capi.DeviceGraph.from_graph(g)
    dec = capi.LaneDecoder(dg, capi.make_options(**OPTS), max_lanes=n, hash_capacity=1 << 14,
                           arena_records=1 << 19)
    A, B = list(range(0, 6)), list(range(6, 12))
    keep = [np.ascontiguousarray(m) for m in mats]

    def go(group):
        return dec.advance_async(group, [keep[u].ctypes.data for u in group], [T] * len(group),
                                 keep[0].shape[1], None, -1, capi.KD_MEM_HOST, init=True,
                                 finalize=True)
    for rnd in range(3):
        ta = go(A)
        tb = go(B)
        # no wait: num_frames_decoded completes the owning call
        assert dec.num_frames_decoded(7) == T
        ra = dec.results(ta)
        rb = dec.results(tb)
        for grp, res in ((A, ra), (B, rb)):
            for k, u in enumerate(grp):
                assert _same_raw(res[k], want[u][0]), (rnd, u)
    # streaming on top of an async call: advance in two halves, finalize at the end
    h = T // 2
    t1 = dec.advance_async(A, [keep[u].ctypes.data for u in A], [h] * 6, keep[0].shape[1], None, -1,
                           capi.KD_MEM_HOST, init=True, finalize=False)
    t2 = dec.advance_async(A, [keep[u][h:].ctypes.data for u in A], [T - h] * 6, keep[0].shape[1],
                           [h] * 6, -1, capi.KD_MEM_HOST, init=False, finalize=True)
    dec.wait(-1)
    res = dec.results(t2)
    for k, u in enumerate(A):
        assert _same_raw(res[k], want[u][0]), u
    with pytest.raises(capi.KdError):
        dec.results(t1)  # not a finalize call


def test_device_input_is_ordered_behind_its_producer_stream():
    """The log-probs are produced on a torch side stream right before the call."""
    torch = pytest.importorskip("torch")
    g = small_graph("HLG")
    n, T = 4, 150
    mats = [synth.make_logprobs(g, T, seed=700 + u, peak=6) for u in range(n)]
    want = _oracle_paths(g, mats, OPTS)
    dg = capi.DeviceGraph.from_graph(g)
    dec = capi.LaneDecoder(dg, capi.make_options(**OPTS), max_lanes=n, hash_capacity=1 << 14,
                           arena_records=1 << 19)
    host = torch.from_numpy(np.stack(mats)).pin_memory()
    side = torch.cuda.Stream()
    big = torch.empty((64, 1024, 1024), device="cuda")
    for rnd in range(3):
        dev = torch.empty((n, T, mats[0].shape[1]), device="cuda")
        with torch.cuda.stream(side):
            for _ in range(20):
                big.normal_()          # keeps the side stream busy for a while
            dev.copy_(host, non_blocking=True)
            dev.add_(0.0)
        t = dec.advance_async(list(range(n)), [dev[u].data_ptr() for u in range(n)], [T] * n,
                              mats[0].shape[1], None, -1, capi.KD_MEM_DEVICE, init=True,
                              finalize=True, producer_stream=side.cuda_stream)
        res = dec.results(t)
        for u in range(n):
            assert _same_raw(res[u], want[u][0]), (rnd, u)


def _tie_graph(seed):
    """Random FST whose weights are multiples of 1/4 and whose epsilon arcs form many
    parallel routes: bit-equal epsilon arrivals at a state are the rule, not the exception."""
    rng = np.random.default_rng(seed)
    S, V = 120, 12
    n_e, n_n = 900, 700
    src = np.concatenate([rng.integers(0, S, n_e), rng.integers(0, S - 1, n_n)])
    il = np.concatenate([rng.integers(1, V + 1, n_e), np.zeros(n_n, np.int64)])
    dst_e = rng.integers(0, S, n_e)
    dst_n = np.array([rng.integers(s + 1, min(S, s + 6)) for s in src[n_e:]])
    dst = np.concatenate([dst_e, dst_n])
    ol = np.where(rng.random(n_e + n_n) < 0.5, rng.integers(1, 40, n_e + n_n), 0)
    w = (rng.integers(0, 5, n_e + n_n) * 0.25).astype(np.float32)
    final = np.where(rng.random(S) < 0.3, 0.5, np.inf).astype(np.float32)
    return synth.graph_from_arcs(S, 0, src, il, ol, w, dst, final, name=f"ties-{seed}",
                                 lm={"kind": "h", "vocab": V})


@pytest.mark.parametrize("seed", range(4))
def test_epsilon_ties_are_resolved_deterministically(seed):
    """Equal-cost epsilon arrivals: the lower epsilon-arc index is the backpointer, on the
    device and in the canonical oracle -- so the raw best path (where every olabel sits)
    is identical, run after run."""
    g = _tie_graph(seed)
    opts = dict(beam=30.0, max_active=2**31 - 1, min_active=0)
    rng = np.random.default_rng(100 + seed)
    n, T, V = 6, 40, 12
    mats = []
    for u in range(n):
        # log-probs on a 1/8 grid as well: emitting ties too
        x = (rng.integers(-40, 0, size=(T, V)) * 0.125).astype(np.float32)
        mats.append(x)
    og = kd_oracle.OracleGraph(g)
    want, ties = [], 0
    for m in mats:
        o = kd_oracle.OracleDecoder(og, kd_ref.Options(**opts), kd_oracle.CANONICAL)
        o.decode(m)
        want.append(o.get_best_path(True, raw=True))
        ties += o.stats()["eps_ties"]
    assert ties > 100, ties
    dg = capi.DeviceGraph.from_graph(g)
    dec = capi.LaneDecoder(dg, capi.make_options(**opts), max_lanes=n, hash_capacity=1 << 12,
                           arena_records=1 << 18)
    lanes = list(range(n))
    for rnd in range(5):
        pb = dec.decode(lanes, mats, True)
        for u in lanes:
            assert _same_raw(pb[u], want[u]), (seed, rnd, u)


def test_lane_error_does_not_leak_its_prefetched_row_into_the_next_lane():
    """More lanes than CTA slots: a CTA whose lane stops on an overflow (with the bulk copy
    of its next row already in flight) goes on to another lane, which must start from its
    own first row."""
    g = small_graph("HL")
    opts = dict(beam=12.0, max_active=2**31 - 1, min_active=0)
    n, T = 1300, 24
    V = int(g.lm["vocab"])
    good = [synth.make_logprobs(g, T, seed=900 + u, peak=9) for u in range(8)]
    og = kd_oracle.OracleGraph(g)
    want = []
    for m in good:
        o = kd_oracle.OracleDecoder(og, kd_ref.Options(**opts), kd_oracle.CANONICAL)
        o.decode(m)
        want.append((sorted_tokens(*o.tokens()), o.stats()["max_tokens"]))
    cap = 1 << 13
    assert max(w[1] for w in want) < cap // 16
    flat = np.full((T, V), np.float32(-np.log(V)), dtype=np.float32)  # every arc survives
    flat[:, :] += (np.arange(V, dtype=np.float32) * np.float32(1e-4))[None, :]
    o = kd_oracle.OracleDecoder(og, kd_ref.Options(**opts), kd_oracle.CANONICAL)
    o.decode(flat[:6])
    assert o.stats()["max_tokens"] > cap  # the flat lanes do overflow the table (capacity / 2 tokens)
    bad = {u for u in range(n) if u < 1100 and u % 3 == 0}
    mats = [flat if u in bad else good[u % len(good)] for u in range(n)]
    dg = capi.DeviceGraph.from_graph(g)
    dec = capi.LaneDecoder(dg, capi.make_options(**opts), max_lanes=n, hash_capacity=cap,
                           arena_records=1 << 15, threads_per_lane=160)
    lanes = list(range(n))
    with pytest.raises(capi.KdError, match="overflow"):
        dec.decode(lanes, mats, True)
    for u in lanes:
        if u in bad:
            continue
        assert dec.num_frames_decoded(u) == T, u
        gs, gc = sorted_tokens(*dec.tokens(u))
        (os_, oc), _ = want[u % len(good)]
        assert np.array_equal(gs, os_) and np.array_equal(gc, oc), u
    for u in sorted(bad)[:5]:
        with pytest.raises(capi.KdError):
            dec.tokens(u) if dec.num_frames_decoded(u) == T else dec.advance([u], [flat])
