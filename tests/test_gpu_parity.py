"""CUDA path (through the C ABI) vs the CPU oracle -- needs a B200."""
import numpy as np
import pytest

from common import OPTION_SETS, rel_close, same_labels, small_graph, sorted_tokens
from kaldi_decoder_b200 import capi, synth
from oracle import kd_oracle, kd_ref

pytestmark = pytest.mark.gpu


def _opts(o):
    return capi.make_options(**o), kd_ref.Options(**o)


@pytest.mark.parametrize("gname", ["H", "HL", "HLG"])
@pytest.mark.parametrize("oi", range(len(OPTION_SETS)))
def test_token_sets_match_canonical_oracle_every_frame(gname, oi):
    """After every frame the device token list equals the canonical oracle's:
    same states, bit-identical fp64 costs."""
    g = small_graph(gname)
    kopts, ropts = _opts(OPTION_SETS[oi])
    dg = capi.DeviceGraph.from_graph(g)
    og = kd_oracle.OracleGraph(g)
    n_lanes, T = 4, 100
    dec = capi.LaneDecoder(dg, kopts, max_lanes=n_lanes, hash_capacity=1 << 15,
                           arena_records=1 << 20)
    peak = [8, 4, 6, 3, 5][oi]
    mats = [synth.make_logprobs(g, T, seed=100 * oi + u, peak=peak) for u in range(n_lanes)]
    oracles = [kd_oracle.OracleDecoder(og, ropts, kd_oracle.CANONICAL) for _ in range(n_lanes)]
    lanes = list(range(n_lanes))
    dec.init(lanes)
    for o in oracles:
        o.init_decoding()
    for f in range(T + 1):
        for u in lanes:
            gs, gc = sorted_tokens(*dec.tokens(u))
            os_, oc = sorted_tokens(*oracles[u].tokens())
            assert np.array_equal(gs, os_), (gname, oi, u, f, len(gs), len(os_))
            assert np.array_equal(gc, oc), (gname, oi, u, f)
        if f == T:
            break
        dec.advance(lanes, mats, max_num_frames=1)
        for u in lanes:
            oracles[u].advance_decoding(mats[u], 0, 1)
            assert dec.num_frames_decoded(u) == f + 1
    # best path: same labels / weights as the canonical oracle up to exact ties
    paths = dec.best_paths(lanes, True)
    for u in lanes:
        ob = oracles[u].get_best_path(True, raw=True)
        assert paths[u].ok == ob.ok
        assert paths[u].reached_final == oracles[u].reached_final()
        if ob.ok:
            assert rel_close(paths[u].total_cost, ob.total_cost, 1e-6)
            if not same_labels(paths[u], ob):
                # only an exact cost tie may change the labels
                assert paths[u].total_cost == pytest.approx(ob.total_cost, rel=1e-7)
    st = dec.stats()
    osum = {k: sum(o.stats()[k] for o in oracles) for k in ("frames", "tokens_in", "tokens_out",
                                                            "tokens_expanded", "emit_arcs")}
    for k, v in osum.items():
        assert st[k] == v, (k, st[k], v)


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_random_fsts_every_frame(seed):
    """Unstructured random graphs: negative weights, non-deterministic states, epsilon
    chains, dense states (label tables) -- device token sets == canonical oracle, every frame."""
    g = synth.make_random_fst(num_states=150 + 40 * seed, num_arcs=1500 + 300 * seed, vocab=25,
                              eps_frac=0.1 + 0.03 * seed, seed=seed)
    o = dict(beam=[6.0, 10.0, 14.0][seed % 3], max_active=[2**31 - 1, 60, 400][seed % 3],
             min_active=[0, 5, 20][seed % 3])
    kopts, ropts = capi.make_options(**o), kd_ref.Options(**o)
    dg = capi.DeviceGraph.from_graph(g)
    og = kd_oracle.OracleGraph(g)
    n_lanes, T = 3, 60
    dec = capi.LaneDecoder(dg, kopts, max_lanes=n_lanes, hash_capacity=1 << 14, arena_records=1 << 18)
    rng = np.random.default_rng(seed)
    mats = []
    for u in range(n_lanes):
        x = rng.standard_normal((T, 25)).astype(np.float32) * np.float32(1.5)
        x[np.arange(T), rng.integers(0, 25, size=T)] += np.float32(5.0)
        x -= np.log(np.exp(x).sum(axis=1, keepdims=True))
        mats.append(x.astype(np.float32))
    oracles = [kd_oracle.OracleDecoder(og, ropts, kd_oracle.CANONICAL) for _ in range(n_lanes)]
    lanes = list(range(n_lanes))
    dec.init(lanes)
    for orc in oracles:
        orc.init_decoding()
    for f in range(T + 1):
        for u in lanes:
            gs, gc = sorted_tokens(*dec.tokens(u))
            os_, oc = sorted_tokens(*oracles[u].tokens())
            assert np.array_equal(gs, os_), (seed, u, f, len(gs), len(os_))
            assert np.array_equal(gc, oc), (seed, u, f)
        if f == T:
            break
        dec.advance(lanes, mats, max_num_frames=1)
        for u in lanes:
            oracles[u].advance_decoding(mats[u], 0, 1)
    paths = dec.best_paths(lanes, True)
    for u in lanes:
        ob = oracles[u].get_best_path(True, raw=True)
        assert paths[u].ok == ob.ok and paths[u].reached_final == oracles[u].reached_final()
        if ob.ok:
            assert rel_close(paths[u].total_cost, ob.total_cost, 1e-6)


def _every_frame(g, o, mats, **dec_kw):
    kopts, ropts = capi.make_options(**o), kd_ref.Options(**o)
    dg = capi.DeviceGraph.from_graph(g)
    og = kd_oracle.OracleGraph(g)
    lanes = list(range(len(mats)))
    dec = capi.LaneDecoder(dg, kopts, max_lanes=len(mats), **dec_kw)
    oracles = [kd_oracle.OracleDecoder(og, ropts, kd_oracle.CANONICAL) for _ in lanes]
    dec.init(lanes)
    for orc in oracles:
        orc.init_decoding()
    peak_tokens = 0
    for f in range(mats[0].shape[0]):
        dec.advance(lanes, mats, max_num_frames=1)
        for u in lanes:
            oracles[u].advance_decoding(mats[u], 0, 1)
            gs, gc = sorted_tokens(*dec.tokens(u))
            os_, oc = sorted_tokens(*oracles[u].tokens())
            assert np.array_equal(gs, os_), (u, f, len(gs), len(os_))
            assert np.array_equal(gc, oc), (u, f)
            peak_tokens = max(peak_tokens, len(gs))
    paths = dec.best_paths(lanes, True)
    for u in lanes:
        ob = oracles[u].get_best_path(True, raw=True)
        assert paths[u].ok == ob.ok and paths[u].reached_final == oracles[u].reached_final()
        if ob.ok:
            assert rel_close(paths[u].total_cost, ob.total_cost, 1e-6)
    return dec, peak_tokens


def test_full_candidate_buffer_leaves_holes_not_tokens():
    """A candidate buffer (hash_capacity / 4 records) far smaller than a frame's candidates:
    the overflow is recombined against the running cutoff, and what the exact cutoff then
    rejects must not show up as tokens (nor in counts that feed GetCutoff)."""
    g = small_graph("H")
    T = 40
    mats = [synth.make_logprobs(g, T, seed=900 + u, peak=2) for u in range(3)]
    for o in (dict(beam=9.0, max_active=2**31 - 1, min_active=0),
              dict(beam=12.0, max_active=25, min_active=10)):
        dec, _ = _every_frame(g, o, mats, hash_capacity=256, arena_records=1 << 16)
        st = dec.stats()
        assert st["candidates"] > 64 * st["frames"] // 2  # the buffer did overflow


def test_front_list_overflow_falls_back_to_the_block():
    """More tokens close to the best than the front list holds (2048)."""
    g = synth.make_random_fst(num_states=9000, num_arcs=120000, vocab=25, eps_frac=0.05, seed=77)
    rng = np.random.default_rng(5)
    T = 8
    mats = []
    for u in range(2):
        x = rng.standard_normal((T, 25)).astype(np.float32) * np.float32(0.05)
        x -= np.log(np.exp(x).sum(axis=1, keepdims=True))
        mats.append(x.astype(np.float32))
    o = dict(beam=40.0, max_active=2**31 - 1, min_active=0)
    _, peak_tokens = _every_frame(g, o, mats, hash_capacity=1 << 16, arena_records=1 << 19)
    assert peak_tokens > 4096


def test_wide_rows_take_the_global_memory_path():
    """More columns than fit the shared-memory row staging: the row is gathered from global
    memory (no label order); results must not change."""
    g = small_graph("HL")
    opts = dict(beam=14.0, max_active=500, min_active=20)
    T, V = 50, 50
    lp = synth.make_logprobs(g, T, seed=8, peak=6)
    dg = capi.DeviceGraph.from_graph(g)
    dec = capi.LaneDecoder(dg, capi.make_options(**opts), max_lanes=2, hash_capacity=1 << 14,
                           arena_records=1 << 18)
    # 9000 columns: the row stays in global memory; 3000: staged in shared memory but too
    # wide for the per-frame label order (no label lookups)
    for W in (9000, 3000):
        wide = np.full((T, W), -30.0, dtype=np.float32)
        wide[:, :V] = lp
        dec.init([0, 1])
        dec.advance([0], [lp])
        dec.advance([1], [wide])
        s0, c0 = sorted_tokens(*dec.tokens(0))
        s1, c1 = sorted_tokens(*dec.tokens(1))
        assert np.array_equal(s0, s1) and np.array_equal(c0, c1), W
        a, b = dec.best_paths([0, 1])
        assert np.array_equal(a.ilabels, b.ilabels) and np.array_equal(a.acoustic, b.acoustic), W
    with pytest.raises(capi.KdError, match="duplicate lane"):
        dec.init([0, 0])
