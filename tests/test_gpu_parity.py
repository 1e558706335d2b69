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
