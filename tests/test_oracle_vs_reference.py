"""oracle/kd_oracle.cc (own restatement) against oracle/_ref (the unmodified reference
object code).  Skipped where the compiled reference is not available."""
import numpy as np
import pytest

from common import OPTION_SETS, rel_close, same_labels, small_graph
from kaldi_decoder_b200 import synth
from oracle import kd_oracle, kd_ref

pytestmark = pytest.mark.skipif(not kd_ref.available(), reason="oracle/_ref not built")


@pytest.mark.parametrize("gname", ["H", "HL", "HLG"])
@pytest.mark.parametrize("oi", range(len(OPTION_SETS)))
def test_every_frame_token_list_is_identical(gname, oi):
    g = small_graph(gname)
    rg, og = kd_ref.RefGraph(g), kd_oracle.OracleGraph(g)
    opts = kd_ref.Options(**OPTION_SETS[oi])
    peak = [8, 4, 6, 3, 5][oi]
    for seed in range(2):
        lp = synth.make_logprobs(g, 100, seed=50 * oi + seed, peak=peak)
        r = kd_ref.RefDecoder(rg, opts)
        o = kd_oracle.OracleDecoder(og, opts, kd_oracle.REFERENCE_ORDER)
        r.init_decoding()
        o.init_decoding()
        for f in range(100):
            r.advance_decoding(lp, 0, 1)
            o.advance_decoding(lp, 0, 1)
            rs, rc = r.tokens()
            os_, oc = o.tokens()
            assert np.array_equal(rs, os_) and np.array_equal(rc, oc), (gname, oi, seed, f)
        assert r.reached_final() == o.reached_final()
        for ufp in (True, False):
            a, b = r.get_best_path(ufp), o.get_best_path(ufp)
            assert a.ok == b.ok
            for x, y in ((a.ilabels, b.ilabels), (a.olabels, b.olabels), (a.graph, b.graph),
                         (a.acoustic, b.acoustic), (a.final, b.final)):
                assert np.array_equal(x, y)


def test_streaming_chunks_and_object_reuse():
    """advance_decoding in chunks with offsets == one shot; a decoder object reused for a
    second utterance keeps the reference's persistent hash size (and still matches)."""
    g = small_graph("HLG")
    rg, og = kd_ref.RefGraph(g), kd_oracle.OracleGraph(g)
    opts = kd_ref.Options(beam=12.0, max_active=200, min_active=20)
    r = kd_ref.RefDecoder(rg, opts)
    o = kd_oracle.OracleDecoder(og, opts, kd_oracle.REFERENCE_ORDER)
    for seed in (3, 4):
        lp = synth.make_logprobs(g, 90, seed=seed, peak=4)
        r.init_decoding()
        o.init_decoding()
        for a in range(0, 90, 30):
            r.advance_decoding(lp[a:a + 30], a, -1)
            o.advance_decoding(lp[a:a + 30], a, 7)
            o.advance_decoding(lp[a:a + 30], a, -1)
            assert r.num_frames_decoded() == o.num_frames_decoded() == a + 30
            rs, rc = r.tokens()
            os_, oc = o.tokens()
            assert np.array_equal(rs, os_) and np.array_equal(rc, oc)
        assert same_labels(r.get_best_path(), o.get_best_path())


def test_batch_entry_points_agree():
    g = small_graph("HL")
    rg, og = kd_ref.RefGraph(g), kd_oracle.OracleGraph(g)
    opts = kd_ref.Options(beam=20.0, max_active=7000)
    mats = synth.make_batch(g, 6, 80, seed=9, peak=8)
    _, rp, rrf = kd_ref.decode_batch(rg, mats, opts, 3)
    _, op, orf, stats, per = kd_oracle.decode_batch(og, mats, opts, 3, kd_oracle.REFERENCE_ORDER)
    assert np.array_equal(rrf, orf)
    for a, b in zip(rp, op):
        assert np.array_equal(a.ilabels, b.ilabels) and np.array_equal(a.acoustic, b.acoustic)
    assert stats["frames"] == 6 * 80 and per.shape[0] == 6


@pytest.mark.parametrize("seed", range(5))
def test_random_fsts_negative_weights_nondeterminism(seed):
    """Unstructured random graphs (negative weights, duplicate ilabels per state, epsilon
    chains): the restatement still equals the reference frame by frame."""
    g = synth.make_random_fst(num_states=120 + 30 * seed, num_arcs=1200 + 200 * seed, vocab=20,
                              eps_frac=0.1 + 0.04 * seed, seed=seed)
    rg, og = kd_ref.RefGraph(g), kd_oracle.OracleGraph(g)
    opts = kd_ref.Options(beam=[6.0, 10.0, 14.0][seed % 3], max_active=[2**31 - 1, 60, 400][seed % 3],
                          min_active=[0, 5, 20][seed % 3])
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((70, 20)).astype(np.float32) * np.float32(1.5)
    x[np.arange(70), rng.integers(0, 20, size=70)] += np.float32(4.0)
    x = (x - np.log(np.exp(x).sum(axis=1, keepdims=True))).astype(np.float32)
    r = kd_ref.RefDecoder(rg, opts)
    o = kd_oracle.OracleDecoder(og, opts, kd_oracle.REFERENCE_ORDER)
    r.init_decoding()
    o.init_decoding()
    for f in range(70):
        r.advance_decoding(x, 0, 1)
        o.advance_decoding(x, 0, 1)
        rs, rc = r.tokens()
        os_, oc = o.tokens()
        assert np.array_equal(rs, os_) and np.array_equal(rc, oc), (seed, f)
    a, b = r.get_best_path(), o.get_best_path()
    assert a.ok == b.ok and np.array_equal(a.ilabels, b.ilabels) and np.array_equal(a.graph, b.graph)
    assert np.array_equal(a.acoustic, b.acoustic) and np.array_equal(a.final, b.final)


def _frc(states, costs, final):
    """SimpleDecoder::FinalRelativeCost (simple-decoder.cc:78-101) from a token list."""
    if len(states) == 0:
        return float("inf")
    f = final[states].astype(np.float64)
    with np.errstate(invalid="ignore"):
        return float(np.float32((costs + f).min() - costs.min()))


@pytest.mark.parametrize("gname,beam,peak", [("H", 6.0, 4), ("HL", 8.0, 6), ("HLG", 10.0, 5),
                                             ("HLG", 3.0, 8), ("HL", 14.0, 3)])
def test_simple_mode_matches_reference_simple_decoder(gname, beam, peak):
    """oracle mode 2 (order-independent SimpleDecoder) vs the unmodified simple-decoder.cc,
    after every frame: ReachedFinal, FinalRelativeCost, best path."""
    import math
    g = small_graph(gname)
    mat = synth.make_logprobs(g, 50, seed=77 + int(beam), peak=peak)
    orc = kd_oracle.OracleDecoder(kd_oracle.OracleGraph(g), kd_ref.Options(beam=beam),
                                  kd_oracle.SIMPLE)
    ref = kd_ref.RefSimpleDecoder(kd_ref.RefGraph(g), beam)
    orc.init_decoding()
    ref.init_decoding()
    fin = np.asarray(g.final)
    for f in range(mat.shape[0] + 1):
        assert orc.reached_final() == ref.reached_final(), f
        st, co = orc.tokens()
        a, b = _frc(st, co, fin), ref.final_relative_cost()
        assert (math.isinf(a) and math.isinf(b)) or abs(a - b) <= 1e-4 * max(1.0, abs(b)), (f, a, b)
        for ufp in (True, False):
            p, r = orc.get_best_path(ufp), ref.get_best_path(ufp)
            assert p.ok == r.ok
            if r.ok:
                assert rel_close(p.total_cost, r.total_cost, 1e-5), (f, ufp)
                if not (np.array_equal(p.isyms, r.isyms) and np.array_equal(p.osyms, r.osyms)):
                    assert p.total_cost == pytest.approx(r.total_cost, rel=1e-6)
        if f < mat.shape[0]:
            orc.advance_decoding(mat, 0, 1)
            ref.advance_decoding(mat, 0, 1)


@pytest.mark.parametrize("seed", range(6))
def test_simple_mode_random_fsts(seed):
    """Unstructured graphs (negative weights, epsilon chains, non-determinism): oracle mode 2 vs
    the reference SimpleDecoder after every frame."""
    import math
    g = synth.make_random_fst(num_states=100 + 30 * seed, num_arcs=1200 + 200 * seed, vocab=20,
                              eps_frac=0.1 + 0.02 * seed, seed=300 + seed)
    rng = np.random.default_rng(seed)
    T = 40
    x = rng.standard_normal((T, 20)).astype(np.float32) * np.float32(1.5)
    x[np.arange(T), rng.integers(0, 20, size=T)] += np.float32(4.0)
    x -= np.log(np.exp(x).sum(axis=1, keepdims=True))
    mat = x.astype(np.float32)
    beam = [4.0, 7.0, 11.0][seed % 3]
    orc = kd_oracle.OracleDecoder(kd_oracle.OracleGraph(g), kd_ref.Options(beam=beam),
                                  kd_oracle.SIMPLE)
    ref = kd_ref.RefSimpleDecoder(kd_ref.RefGraph(g), beam)
    orc.init_decoding()
    ref.init_decoding()
    fin = np.asarray(g.final)
    for f in range(T + 1):
        assert orc.reached_final() == ref.reached_final(), f
        st, co = orc.tokens()
        a, b = _frc(st, co, fin), ref.final_relative_cost()
        assert (math.isinf(a) and math.isinf(b)) or abs(a - b) <= 1e-4 * max(1.0, abs(b)), (f, a, b)
        p, r = orc.get_best_path(True), ref.get_best_path(True)
        assert p.ok == r.ok, f
        if r.ok:
            assert rel_close(p.total_cost, r.total_cost, 1e-5), (f, p.total_cost, r.total_cost)
            if not (np.array_equal(p.isyms, r.isyms) and np.array_equal(p.osyms, r.osyms)):
                assert p.total_cost == pytest.approx(r.total_cost, rel=1e-6)
        if f < T:
            orc.advance_decoding(mat, 0, 1)
            ref.advance_decoding(mat, 0, 1)
