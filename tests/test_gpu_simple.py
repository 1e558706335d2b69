"""SimpleDecoder search mode on the GPU vs the unmodified reference
SimpleDecoder (oracle/_ref, simple-decoder.cc) -- needs a B200."""
import math

import numpy as np
import pytest

from common import rel_close, small_graph, sorted_tokens
from kaldi_decoder_b200 import capi, synth
from oracle import kd_oracle, kd_ref

pytestmark = pytest.mark.gpu

needs_ref = pytest.mark.skipif(not kd_ref.available(), reason="oracle/_ref not built")


def _same_final_relative_cost(a: float, b: float) -> bool:
    if math.isinf(a) or math.isinf(b):
        return math.isinf(a) and math.isinf(b)
    return abs(a - b) <= 1e-4 * max(1.0, abs(a), abs(b))


def _compare_every_frame(g, beam, mat, **dec_kw):
    dg = capi.DeviceGraph.from_graph(g)
    dec = capi.LaneDecoder(dg, capi.make_options(beam=beam), max_lanes=1,
                           search=capi.KD_SEARCH_SIMPLE, **dec_kw)
    ref = kd_ref.RefSimpleDecoder(kd_ref.RefGraph(g), beam)
    dec.init([0])
    ref.init_decoding()
    T = mat.shape[0]
    for f in range(T + 1):
        assert dec.num_frames_decoded(0) == ref.num_frames_decoded() == f
        for ufp in (True, False):
            p = dec.best_paths([0], ufp)[0]
            r = ref.get_best_path(ufp)
            assert p.ok == r.ok, (f, ufp)
            assert p.reached_final == ref.reached_final(), (f, ufp)
            if not r.ok:
                continue
            # total cost within float tolerance (the acoustic part of each arc is recovered
            # from cost differences); labels identical unless the costs tie
            assert rel_close(p.total_cost, r.total_cost, 1e-5), (f, ufp, p.total_cost, r.total_cost)
            if not (np.array_equal(p.isyms, r.isyms) and np.array_equal(p.osyms, r.osyms)):
                assert p.total_cost == pytest.approx(r.total_cost, rel=1e-6), (f, ufp)
        assert _same_final_relative_cost(dec.final_relative_cost(0), ref.final_relative_cost()), f
        if f == T:
            break
        dec.advance([0], [mat], max_num_frames=1)
        ref.advance_decoding(mat, 0, 1)
    return dec


@needs_ref
@pytest.mark.parametrize("gname,beam,peak", [("H", 6.0, 4), ("HL", 8.0, 6), ("HLG", 10.0, 5),
                                             ("HLG", 3.0, 8), ("HL", 14.0, 3)])
def test_simple_search_matches_reference_every_frame(gname, beam, peak):
    g = small_graph(gname)
    mat = synth.make_logprobs(g, 60, seed=31 + int(beam), peak=peak)
    dec = _compare_every_frame(g, beam, mat, hash_capacity=1 << 17, arena_records=1 << 19)
    st = dec.stats()
    assert st["frames"] == 60 and st["tokens_out"] > 0


@needs_ref
@pytest.mark.parametrize("seed", range(3))
def test_simple_search_random_fsts(seed):
    g = synth.make_random_fst(num_states=120 + 30 * seed, num_arcs=1400, vocab=20,
                              eps_frac=0.12, seed=40 + seed)
    rng = np.random.default_rng(seed)
    T = 40
    x = rng.standard_normal((T, 20)).astype(np.float32) * np.float32(1.5)
    x[np.arange(T), rng.integers(0, 20, size=T)] += np.float32(4.0)
    x -= np.log(np.exp(x).sum(axis=1, keepdims=True))
    _compare_every_frame(g, [5.0, 9.0, 12.0][seed], x.astype(np.float32),
                         hash_capacity=1 << 14, arena_records=1 << 18)


@needs_ref
def test_python_simple_decoder_drop_in():
    """kaldi_decoder.SimpleDecoder(fst, beam): the reference's Python surface
    (python/csrc/simple-decoder.cc:14-44)."""
    import kaldi_decoder as kd

    g = small_graph("HLG")
    mat = synth.make_logprobs(g, 80, seed=5, peak=6)
    fst = kd.StdVectorFst.from_arrays(g.num_states, g.start, g.row_off, g.ilabel, g.olabel,
                                      g.weight, g.nextstate, g.final)
    dec = kd.SimpleDecoder(fst, 9.0)
    assert dec.decode(kd.DecodableCtc(mat)) is True
    assert dec.num_frames_decoded() == 80
    ok, lat = dec.get_best_path()
    assert ok
    ok2, isyms, osyms, weight = kd.get_linear_symbol_sequence(lat)
    ref = kd_ref.RefSimpleDecoder(kd_ref.RefGraph(g), 9.0)
    ref.init_decoding()
    ref.advance_decoding(mat)
    r = ref.get_best_path()
    assert ok2 and list(isyms) == list(r.isyms) and list(osyms) == list(r.osyms)
    assert rel_close(float(weight[0]) + float(weight[1]), r.total_cost, 1e-5)
    assert dec.reached_final() == ref.reached_final()
    assert _same_final_relative_cost(dec.final_relative_cost(), ref.final_relative_cost())
    # streaming calls
    dec.init_decoding()
    dec.advance_decoding(kd.DecodableCtc(mat), 30)
    assert dec.num_frames_decoded() == 30
    dec.advance_decoding(kd.DecodableCtc(mat))
    assert dec.num_frames_decoded() == 80
    ok3, lat3 = dec.get_best_path()
    assert ok3 and list(kd.get_linear_symbol_sequence(lat3)[1]) == list(isyms)


@pytest.mark.parametrize("gname,beam,peak", [("H", 6.0, 4), ("HL", 8.0, 6), ("HLG", 10.0, 5),
                                             ("HLG", 3.0, 8)])
def test_simple_search_token_sets_equal_the_oracle_every_frame(gname, beam, peak):
    """Device token list (PruneToks view) == oracle mode 2 after every frame: same states,
    bit-identical fp64 costs; raw best paths arc for arc."""
    g = small_graph(gname)
    mat = synth.make_logprobs(g, 60, seed=11 + int(beam), peak=peak)
    dg = capi.DeviceGraph.from_graph(g)
    dec = capi.LaneDecoder(dg, capi.make_options(beam=beam), max_lanes=1,
                           search=capi.KD_SEARCH_SIMPLE, hash_capacity=1 << 16,
                           arena_records=1 << 19)
    orc = kd_oracle.OracleDecoder(kd_oracle.OracleGraph(g), kd_ref.Options(beam=beam),
                                  kd_oracle.SIMPLE)
    dec.init([0])
    orc.init_decoding()
    for f in range(mat.shape[0] + 1):
        gs, gc = sorted_tokens(*dec.tokens(0))
        os_, oc = sorted_tokens(*orc.tokens())
        assert np.array_equal(gs, os_), (f, len(gs), len(os_))
        assert np.array_equal(gc, oc), f
        if f < mat.shape[0]:
            dec.advance([0], [mat], max_num_frames=1)
            orc.advance_decoding(mat, 0, 1)
    p = dec.best_paths([0], True)[0]
    o = orc.get_best_path(True, raw=True)
    assert p.ok == o.ok
    if o.ok:
        assert rel_close(p.total_cost, o.total_cost, 1e-6)
        if np.array_equal(p.ilabels, o.ilabels):
            assert np.array_equal(p.graph, o.graph) and np.array_equal(p.acoustic, o.acoustic)


@pytest.mark.parametrize("seed", range(4))
def test_simple_search_random_fsts_token_sets(seed):
    """Random graphs with negative weights and epsilon chains: device token list == oracle
    mode 2 after every frame (states and bit-identical costs)."""
    g = synth.make_random_fst(num_states=120 + 40 * seed, num_arcs=1500 + 300 * seed, vocab=20,
                              eps_frac=0.1 + 0.03 * seed, seed=500 + seed)
    rng = np.random.default_rng(100 + seed)
    T = 40
    x = rng.standard_normal((T, 20)).astype(np.float32) * np.float32(1.5)
    x[np.arange(T), rng.integers(0, 20, size=T)] += np.float32(4.0)
    x -= np.log(np.exp(x).sum(axis=1, keepdims=True))
    mat = x.astype(np.float32)
    beam = [5.0, 8.0, 12.0, 6.5][seed]
    dg = capi.DeviceGraph.from_graph(g)
    dec = capi.LaneDecoder(dg, capi.make_options(beam=beam), max_lanes=1,
                           search=capi.KD_SEARCH_SIMPLE, hash_capacity=1 << 16,
                           arena_records=1 << 19)
    orc = kd_oracle.OracleDecoder(kd_oracle.OracleGraph(g), kd_ref.Options(beam=beam),
                                  kd_oracle.SIMPLE)
    dec.init([0])
    orc.init_decoding()
    for f in range(T + 1):
        gs, gc = sorted_tokens(*dec.tokens(0))
        os_, oc = sorted_tokens(*orc.tokens())
        assert np.array_equal(gs, os_), (seed, f, len(gs), len(os_))
        assert np.array_equal(gc, oc), (seed, f)
        if f < T:
            dec.advance([0], [mat], max_num_frames=1)
            orc.advance_decoding(mat, 0, 1)
