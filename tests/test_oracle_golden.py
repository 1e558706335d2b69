"""The CPU oracle against the golden vectors frozen from the unmodified reference."""
import numpy as np
import pytest

from common import GOLDEN_CASES, GoldenCase, same_labels
from oracle import kd_oracle, kd_ref


def _paths_equal(a, b):
    return (a.ok == b.ok and np.array_equal(a.ilabels, b.ilabels)
            and np.array_equal(a.olabels, b.olabels) and np.array_equal(a.graph, b.graph)
            and np.array_equal(a.acoustic, b.acoustic) and np.array_equal(a.final, b.final))


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_reference_order_oracle_reproduces_golden_bit_for_bit(name):
    """mode 0 of oracle/kd_oracle.cc == the reference: same token list (order, states,
    fp64 costs) after every frame, same best path arcs and weights."""
    gc = GoldenCase(name)
    og = kd_oracle.OracleGraph(gc.graph)
    for u in range(gc.n_utts):
        dec = kd_oracle.OracleDecoder(og, kd_ref.Options(**gc.opts), kd_oracle.REFERENCE_ORDER)
        dec.init_decoding()
        lp = gc.logp(u)
        for f, (gs, gco) in enumerate(gc.tokens(u)):
            st, co = dec.tokens()
            assert np.array_equal(st, gs), (name, u, f)
            assert np.array_equal(co, gco), (name, u, f)
            if f < gc.T:
                dec.advance_decoding(lp, 0, 1)
        assert dec.reached_final() == gc.reached_final(u)
        for ufp in (True, False):
            assert _paths_equal(dec.get_best_path(ufp), gc.best(u, ufp)), (name, u, ufp)


@pytest.mark.parametrize("name", ["h20_default", "hl300", "hlg300_peaky", "hlg300_nobeam"])
def test_canonical_oracle_matches_golden_labels_when_pruning_is_order_independent(name):
    """mode 1 (what the CUDA kernels implement) gives the reference's label sequences and
    total cost on the fixtures where max_active never binds."""
    gc = GoldenCase(name)
    og = kd_oracle.OracleGraph(gc.graph)
    for u in range(gc.n_utts):
        dec = kd_oracle.OracleDecoder(og, kd_ref.Options(**gc.opts), kd_oracle.CANONICAL)
        dec.decode(gc.logp(u))
        got, want = dec.get_best_path(True), gc.best(u, True)
        assert got.ok == want.ok
        assert dec.reached_final() == gc.reached_final(u)
        assert same_labels(got, want), (name, u)
        assert abs(got.total_cost - want.total_cost) <= 1e-4 * max(1.0, abs(want.total_cost))
        assert dec.stats()["binding_max"] == 0


def test_golden_fixtures_cover_binding_and_ties():
    """The fixture set exercises max_active / min_active binding (order-dependent pruning)."""
    gc = GoldenCase("hlg300_bind")
    og = kd_oracle.OracleGraph(gc.graph)
    dec = kd_oracle.OracleDecoder(og, kd_ref.Options(**gc.opts), kd_oracle.REFERENCE_ORDER)
    dec.decode(gc.logp(0))
    st = dec.stats()
    assert st["binding_max"] > 0 and st["binding_min"] > 0 and st["extras"] > 0
