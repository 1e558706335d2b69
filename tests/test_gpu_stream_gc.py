"""Long streams: the backpointer arena is garbage-collected in place (the reference frees dead
tokens by reference counting, faster-decoder.h:145-155) -- needs a B200."""
import numpy as np
import pytest

from common import small_graph, sorted_tokens
from kaldi_decoder_b200 import capi, synth
from oracle import kd_oracle, kd_ref

pytestmark = pytest.mark.gpu


def _same_raw(p, ob):
    return (p.ok == ob.ok and np.array_equal(p.ilabels, ob.ilabels)
            and np.array_equal(p.olabels, ob.olabels) and np.array_equal(p.graph, ob.graph)
            and np.array_equal(p.acoustic, ob.acoustic))


def test_50k_frames_through_an_arena_sized_for_5k():
    g = small_graph("HLG")
    opts = dict(beam=10.0, max_active=60, min_active=10)
    T, chunk = 50_000, 2_500
    mats = [synth.make_logprobs(g, T, seed=11, peak=6), synth.make_logprobs(g, 20_000, seed=12, peak=5)]
    og = kd_oracle.OracleGraph(g)
    oracles = []
    for m in mats:
        o = kd_oracle.OracleDecoder(og, kd_ref.Options(**opts), kd_oracle.CANONICAL)
        o.decode(m)
        oracles.append(o)
    per_frame = oracles[0].stats()["tokens_out"] / T
    arena = int(per_frame * 5_000)
    assert oracles[0].stats()["tokens_out"] > 8 * arena  # without collection: a hard overflow
    dg = capi.DeviceGraph.from_graph(g)
    dec = capi.LaneDecoder(dg, capi.make_options(**opts), max_lanes=2, hash_capacity=1 << 12,
                           arena_records=arena)
    dec.init([0, 1])
    for f in range(0, T, chunk):
        lanes = [0, 1] if f < mats[1].shape[0] else [0]
        dec.advance(lanes, [mats[u][f:f + chunk] for u in lanes], offsets=[f] * len(lanes))
        if f == 10 * chunk:
            # a partial result in the middle of the stream, after several collections
            mid = kd_oracle.OracleDecoder(og, kd_ref.Options(**opts), kd_oracle.CANONICAL)
            mid.decode(mats[0][:f + chunk])
            assert _same_raw(dec.best_paths([0], True)[0], mid.get_best_path(True, raw=True))
    assert dec.num_frames_decoded(0) == T and dec.num_frames_decoded(1) == mats[1].shape[0]
    st = dec.stats(0)
    assert st["arena_compactions"] >= 8, st
    paths = dec.best_paths([0, 1], True)
    for u in (0, 1):
        gs, gc = sorted_tokens(*dec.tokens(u))
        os_, oc = sorted_tokens(*oracles[u].tokens())
        assert np.array_equal(gs, os_) and np.array_equal(gc, oc), u
        assert _same_raw(paths[u], oracles[u].get_best_path(True, raw=True)), u
        assert len(paths[u].ilabels) >= mats[u].shape[0]


def test_one_call_longer_than_the_arena_and_true_overflow():
    """Collection also happens inside a single AdvanceDecoding call; when even the live
    history does not fit, the lane still fails loudly."""
    g = small_graph("HL")
    opts = dict(beam=9.0, max_active=40, min_active=5)
    T = 12_000
    m = synth.make_logprobs(g, T, seed=21, peak=6)
    og = kd_oracle.OracleGraph(g)
    o = kd_oracle.OracleDecoder(og, kd_ref.Options(**opts), kd_oracle.CANONICAL)
    o.decode(m)
    arena = int(o.stats()["tokens_out"] / T * 1_500)
    dg = capi.DeviceGraph.from_graph(g)
    dec = capi.LaneDecoder(dg, capi.make_options(**opts), max_lanes=1, hash_capacity=1 << 12,
                           arena_records=arena)
    pb = dec.decode([0], [m], True)
    assert dec.stats(0)["arena_compactions"] >= 3
    assert _same_raw(pb[0], o.get_best_path(True, raw=True))
    tiny = capi.LaneDecoder(dg, capi.make_options(**opts), max_lanes=1, hash_capacity=1 << 12,
                            arena_records=64)
    with pytest.raises(capi.KdError, match="arena overflow"):
        tiny.decode([0], [m], True)
