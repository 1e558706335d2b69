"""CPU-only checks: the C-ABI library loads and exports every declared symbol, the Python
package imports, FST I/O and RemoveEpsLocal restatements agree, generators are seeded."""
import ctypes
import os
import re

import numpy as np
import pytest

from common import GoldenCase, small_graph
from kaldi_decoder_b200 import capi, parallel, synth
from oracle import kd_oracle, kd_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_capi_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "kd_capi.h")).read()
    declared = set(re.findall(r"KD_API\s+[\w\s\*]+?\b(kd_\w+)\s*\(", header))
    assert len(declared) >= 18
    assert declared == set(capi.EXPORTED), declared ^ set(capi.EXPORTED)
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in sorted(declared):
        assert getattr(lib, name) is not None


def test_capi_fails_loudly_without_a_gpu():
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    g = small_graph("H")
    with pytest.raises(capi.KdError, match="no CPU fallback"):
        capi.DeviceGraph.from_graph(g)


def test_python_package_surface_matches_reference_names():
    import kaldi_decoder as kd
    for name in ("DecodableCtc", "DecodableInterface", "FasterDecoder", "FasterDecoderOptions"):
        assert hasattr(kd, name)
    o = kd.FasterDecoderOptions()
    assert (o.beam, o.max_active, o.min_active, o.beam_delta, o.hash_ratio) == \
        (16.0, 2**31 - 1, 20, 0.5, 2.0)
    # faster-decoder.h:51-62 ToString format
    assert str(kd.FasterDecoderOptions(beam=20, max_active=7000)) == \
        "FasterDecoderOptions(beam=20, max_active=7000, min_active=20, beam_delta=0.5, hash_ratio=2)"
    o.beam = 11.5
    assert o.beam == 11.5
    d = kd.DecodableCtc(np.arange(12, dtype=np.float32).reshape(3, 4), offset=2)
    # decodable-ctc.cc:22-38: one-based index, frame - offset, frames ready = offset + rows
    assert d.num_frames_ready() == 5 and d.num_indices() == 4
    assert d.log_likelihood(3, 2) == 5.0 and d.is_last_frame(4) and not d.is_last_frame(3)


def test_python_decodable_can_be_subclassed():
    import kaldi_decoder as kd

    class Two(kd.DecodableInterface):
        def __init__(self):
            super().__init__()

        def log_likelihood(self, frame, index):
            return -float(frame + index)

        def is_last_frame(self, frame):
            return frame == 1

        def num_frames_ready(self):
            return 2

        def num_indices(self):
            return 3

    t = Two()
    assert t.num_frames_ready() == 2 and t.log_likelihood(1, 2) == -3.0


def test_fst_text_and_binary_round_trip(tmp_path):
    import kaldi_decoder as kd
    g = small_graph("HL")
    f = kd.StdVectorFst.from_arrays(g.num_states, g.start, g.row_off, g.ilabel, g.olabel,
                                    g.weight, g.nextstate, g.final)
    assert f.num_states == g.num_states and f.start == g.start
    p = str(tmp_path / "g.fst")
    f.write(p)
    h = kd.StdVectorFst.read(p)
    n, start, off, il, ol, w, ns, fin = h.to_arrays()
    assert n == g.num_states and start == g.start
    for a, b in ((off, g.row_off), (il, g.ilabel), (ol, g.olabel), (w, g.weight),
                 (ns, g.nextstate), (fin, g.final)):
        assert np.array_equal(a, b)
    t = kd.StdVectorFst.from_str("0 1 3 4 0.5\n1 1 2 0\n1 2 0 7 1.25\n2 0.75\n")
    assert t.arcs(1) == [(2, 0, 0.0, 1), (0, 7, 1.25, 2)] and t.final(2) == 0.75
    assert kd.StdVectorFst.from_str(t.to_str()).to_str() == t.to_str()


def test_remove_eps_local_restatements_agree():
    """minifst's general RemoveEpsLocal == the linear merges used by the oracle and by
    capi.merge_linear, on random linear chains (PARITY UNPINNED vs kaldifst)."""
    import kaldi_decoder as kd
    rng = np.random.default_rng(0)
    for trial in range(200):
        n = int(rng.integers(0, 12))
        il = rng.choice([0, 0, 1, 2, 3], size=n).astype(np.int32)
        ol = rng.choice([0, 0, 0, 5, 6], size=n).astype(np.int32)
        gw = rng.uniform(0, 2, size=n).astype(np.float32)
        aw = rng.uniform(0, 2, size=n).astype(np.float32)
        fin = (np.float32(rng.uniform(0, 1)), np.float32(0.0))
        lat = kd._kaldi_decoder._remove_eps_local_linear(list(il), list(ol), list(gw), list(aw), fin) \
            if hasattr(kd, "_kaldi_decoder") else None
        if lat is None:
            from kaldi_decoder.lib import _kaldi_decoder as m
            lat = m._remove_eps_local_linear(list(il), list(ol), list(gw), list(aw), fin)
        arcs, s = [], lat.start
        while lat.num_arcs(s) == 1:
            a = lat.arcs(s)[0]
            arcs.append(a[:4])
            s = a[4]
        raw = capi.RawPath(True, True, il, ol, gw, aw, np.array(fin, np.float32))
        mil, mol, mgw, maw, mfin = capi.merge_linear(raw)
        assert [a[0] for a in arcs] == list(mil) and [a[1] for a in arcs] == list(mol), trial
        assert np.allclose([a[2] for a in arcs], mgw, rtol=0, atol=0)
        assert np.allclose([a[3] for a in arcs], maw, rtol=0, atol=0)
        assert lat.final(s) == (float(mfin[0]), float(mfin[1]))
        ok, isy, osy, tot = kd.get_linear_symbol_sequence(lat)
        assert ok and isy == [int(x) for x in il if x] and osy == [int(x) for x in ol if x]


def test_oracle_merge_matches_python_merge_on_golden_raw_paths():
    gc = GoldenCase("hlg300_peaky")
    og = kd_oracle.OracleGraph(gc.graph)
    dec = kd_oracle.OracleDecoder(og, kd_ref.Options(**gc.opts), kd_oracle.REFERENCE_ORDER)
    dec.decode(gc.logp(0))
    raw = dec.get_best_path(True, raw=True)
    merged = dec.get_best_path(True, raw=False)
    rp = capi.RawPath(True, True, raw.ilabels, raw.olabels, raw.graph, raw.acoustic, raw.final)
    il, ol, gw, aw, fin = capi.merge_linear(rp)
    assert np.array_equal(il, merged.ilabels) and np.array_equal(ol, merged.olabels)
    assert np.array_equal(gw, merged.graph) and np.array_equal(aw, merged.acoustic)
    assert np.array_equal(fin, merged.final)
    assert len(raw.ilabels) >= len(merged.ilabels) == gc.T  # one arc per frame survives


def test_generators_are_seeded_and_valid():
    a, b = synth.make_hlg(500, (10, 20), 40, seed=5), synth.make_hlg(500, (10, 20), 40, seed=5)
    for x, y in ((a.row_off, b.row_off), (a.ilabel, b.ilabel), (a.weight, b.weight),
                 (a.nextstate, b.nextstate), (a.final, b.final)):
        assert np.array_equal(x, y)
    a.validate()
    st = a.stats()
    assert st["eps_arcs"] > 0 and st["max_ilabel"] <= 40 and st["final_states"] == 31
    lp1, lp2 = synth.make_logprobs(a, 50, seed=3), synth.make_logprobs(a, 50, seed=3)
    assert np.array_equal(lp1, lp2) and lp1.dtype == np.float32 and lp1.shape == (50, 40)
    assert np.allclose(np.exp(lp1.astype(np.float64)).sum(axis=1), 1.0, atol=1e-4)
    h = synth.make_h(7)
    assert h.num_arcs == 49 and (h.ilabel > 0).all()


def test_shard_helpers():
    for n, w in ((8192, 8), (10, 3), (5, 8), (0, 2)):
        spans = [parallel.shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    parts = parallel.shard_by_length([10, 1, 1, 1, 9, 2, 8], 3)
    assert sorted(sum(parts, [])) == list(range(7))
    loads = [sum([10, 1, 1, 1, 9, 2, 8][i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= 2
