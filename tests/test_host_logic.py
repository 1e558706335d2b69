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


def test_ctypes_structs_mirror_the_header():
    """capi.py restates kd_options / kd_decoder_config / kd_stats for ctypes: same fields, same
    order, same widths as include/kd_capi.h (a field added on one side only would shift every
    counter after it)."""
    header = open(os.path.join(ROOT, "include", "kd_capi.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    ctype = {"int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "float": ctypes.c_float}
    for cname, pystruct in (("kd_options", capi.KdOptions), ("kd_decoder_config", capi.KdConfig),
                            ("kd_stats", capi.KdStats)):
        body = re.search(r"typedef struct\s*\w*\s*\{([^}]*)\}\s*%s\s*;" % cname, header).group(1)
        fields = re.findall(r"(int32_t|int64_t|float)\s+(\w+)\s*;", body)
        assert [(n, ctype[t]) for t, n in fields] == list(pystruct._fields_), cname
    assert "table_retries" in dict(capi.KdStats._fields_)


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


# ---------------------------------------------------------------- SURVEY section 8, rows f1 / f4

def test_lattice_faster_decoder_config_matches_the_reference_struct():
    """lattice-faster-decoder.h:23-134: defaults, field names, ToString, Check."""
    import kaldi_decoder as kd
    c = kd.LatticeFasterDecoderConfig()
    assert (c.beam, c.max_active, c.min_active, c.lattice_beam, c.prune_interval,
            c.determinize_lattice, c.beam_delta, c.hash_ratio, c.memory_pool_tokens_block_size,
            c.memory_pool_links_block_size) == (16.0, 2**31 - 1, 200, 10.0, 25, True, 0.5, 2.0, 256, 256)
    assert abs(c.prune_scale - 0.1) < 1e-7
    assert str(kd.LatticeFasterDecoderConfig(beam=12, max_active=5000, determinize_lattice=False)) == (
        "LatticeFasterDecoderConfig(beam=12, max_active=5000, min_active=200, lattice_beam=10, "
        "prune_interval=25, determinize_lattice=False, beam_delta=0.5, hash_ratio=2, "
        "prune_scale=0.1, memory_pool_tokens_block_size=256, memory_pool_links_block_size=256)")
    c.lattice_beam = 6.5
    assert c.lattice_beam == 6.5
    c.check()
    c.prune_scale = 1.5
    with pytest.raises(RuntimeError):
        c.check()


def _fst_header(fst_type, version, flags, start, num_states, num_arcs):
    import struct
    def s(x):
        return struct.pack("<i", len(x)) + x.encode()
    return (struct.pack("<i", 2125659606) + s(fst_type) + s("standard")
            + struct.pack("<iiQqqq", version, flags, 0, start, num_states, num_arcs))


def _symbol_table(name, symbols):
    import struct
    b = struct.pack("<i", 2125658996) + struct.pack("<i", len(name)) + name.encode()
    b += struct.pack("<qq", len(symbols), len(symbols))
    for k, sym in enumerate(symbols):
        b += struct.pack("<i", len(sym)) + sym.encode() + struct.pack("<q", k)
    return b


# state -> (final weight, [(ilabel, olabel, weight, nextstate)])
_TOY = {0: (float("inf"), [(1, 5, 0.5, 1), (0, 7, 0.25, 2), (3, 0, 1.5, 0)]),
        1: (float("inf"), [(2, 0, 0.0, 2)]),
        2: (0.75, [])}


def _vector_body():
    import struct
    b = b""
    for s in sorted(_TOY):
        fin, arcs = _TOY[s]
        b += struct.pack("<fq", fin, len(arcs))
        for il, ol, w, ns in arcs:
            b += struct.pack("<iifi", il, ol, w, ns)
    return b


def _const_body(prefix_len, aligned):
    import struct
    def pad(n):
        return b"\0" * ((16 - n % 16) % 16) if aligned else b""
    b = pad(prefix_len)
    pos = 0
    for s in sorted(_TOY):
        fin, arcs = _TOY[s]
        nie = sum(1 for a in arcs if a[0] == 0)
        noe = sum(1 for a in arcs if a[1] == 0)
        b += struct.pack("<fIIII", fin, pos, len(arcs), nie, noe)
        pos += len(arcs)
    b += pad(prefix_len + len(b))
    for s in sorted(_TOY):
        for il, ol, w, ns in _TOY[s][1]:
            b += struct.pack("<iifi", il, ol, w, ns)
    return b


@pytest.mark.parametrize("kind", ["vector", "vector+symbols", "const-v2", "const-v1-aligned",
                                  "const-flag-aligned", "const+symbols-aligned"])
def test_openfst_binary_reader_on_byte_level_fixtures(kind, tmp_path):
    """Files assembled byte by byte from OpenFst's published layout (SURVEY.md App. B.5;
    fst.cc FstHeader::Read, vector-fst.h / const-fst.h Read, symbol-table.cc): header, optional
    embedded symbol tables (skipped), `vector` body, `const` body with and without 16-byte
    alignment.  (No real OpenFst is available to write them; the layout is from its source.)"""
    import kaldi_decoder as kd
    n_arcs = sum(len(a) for _, a in _TOY.values())
    syms = b""
    flags = 0
    if "symbols" in kind:
        syms = _symbol_table("isyms", ["<eps>", "a", "b", "c"]) + _symbol_table("osyms", ["<eps>", "x"])
        flags |= 3
    if kind.startswith("vector"):
        data = _fst_header("vector", 2, flags, 0, len(_TOY), n_arcs) + syms + _vector_body()
    else:
        version = 1 if "v1" in kind else 2
        if "flag-aligned" in kind or "symbols-aligned" in kind:
            flags |= 4
        aligned = version == 1 or (flags & 4) != 0
        head = _fst_header("const", version, flags, 0, len(_TOY), n_arcs) + syms
        data = head + _const_body(len(head), aligned)
    path = tmp_path / (kind + ".fst")
    path.write_bytes(data)
    for cls in (kd.StdVectorFst, kd.StdConstFst):
        f = cls.read(str(path))
        assert isinstance(f, kd.StdFst)
        assert f.fst_type == ("vector" if cls is kd.StdVectorFst else "const")
        assert f.start == 0 and f.num_states == len(_TOY)
        for s, (fin, arcs) in _TOY.items():
            assert f.final(s) == np.float32(fin)
            assert f.arcs(s) == [(il, ol, float(np.float32(w)), ns) for il, ol, w, ns in arcs]
    # truncated files fail loudly
    (tmp_path / "cut.fst").write_bytes(data[:len(data) - 5])
    with pytest.raises(RuntimeError):
        kd.StdVectorFst.read(str(tmp_path / "cut.fst"))


def test_const_and_vector_fst_types_and_constructor_overloads():
    """python/csrc/faster-decoder.cc:34-42: FasterDecoder(fst, config) is overloaded on
    Fst / VectorFst / ConstFst.  Without a GPU the constructors must get as far as the
    device ("no CPU fallback"), not fail on the argument type."""
    import kaldi_decoder as kd
    g = small_graph("HL")
    v = kd.StdVectorFst.from_arrays(g.num_states, g.start, g.row_off, g.ilabel, g.olabel, g.weight,
                                    g.nextstate, g.final)
    c = kd.StdConstFst(v)
    c2 = kd.StdConstFst.from_arrays(g.num_states, g.start, g.row_off, g.ilabel, g.olabel, g.weight,
                                    g.nextstate, g.final)
    assert c.num_states == v.num_states == c2.num_states == g.num_states
    for a, b in zip(v.to_arrays(), c.to_arrays()):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    assert kd.StdVectorFst(c).to_str() == v.to_str() == c2.to_str()
    if capi.device_count() > 0:
        pytest.skip("a GPU is present (covered by the GPU tests)")
    opts = kd.FasterDecoderOptions(beam=8.0)
    for f in (v, c):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            kd.FasterDecoder(f, opts)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            kd.SimpleDecoder(f, 8.0)
    with pytest.raises(TypeError):
        kd.FasterDecoder("not an fst", opts)


def test_fst_content_id_follows_the_content_not_the_object():
    """Decoders built from FSTs with the same content_id share one device graph
    (DeviceGraph::Shared): the id must survive copies and conversions, and must not survive a
    modification -- the reference reads its `const Fst&` at decode time, so a stale device copy
    would be a silent difference."""
    import kaldi_decoder as kd
    g = small_graph("HL")
    args = (g.num_states, g.start, g.row_off, g.ilabel, g.olabel, g.weight, g.nextstate, g.final)
    v = kd.StdVectorFst.from_arrays(*args)
    same = v.content_id
    assert same != 0 and v.content_id == same                      # reading does not change it
    assert kd.StdVectorFst.from_arrays(*args).content_id != same   # equal arcs, another object: no claim
    assert kd.StdConstFst(v).content_id == same                    # conversions hold the same content
    assert kd.StdVectorFst(kd.StdConstFst(v)).content_id == same
    w = kd.StdVectorFst(v)
    for edit in (lambda f: f.add_state(), lambda f: f.set_start(1), lambda f: f.set_final(0, 1.5),
                 lambda f: f.add_arc(0, 1, 2, 0.5, 1)):
        before = w.content_id
        edit(w)
        assert w.content_id != before and w.content_id != same
    assert v.content_id == same                                    # the source of the copy is untouched
    assert w.num_states == v.num_states + 1 and w.start == 1 and w.final(0) == 1.5


def test_host_sources_compile_against_an_openfst_shaped_fst_h(tmp_path):
    """with OpenFst on the include path, csrc/minifst/fst is dropped.  The
    host sources are syntax-checked against tests/openfst_shape (OpenFst's Fst / ExpandedFst /
    MutableFst split: no NumStates() on Fst<Arc>, iterator data with a polymorphic base);
    a translation unit that does call Fst::NumStates() must be rejected by the same header."""
    import shutil
    import subprocess
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    inc = ["-I", os.path.join(ROOT, "tests", "openfst_shape"),
           "-I", os.path.join(ROOT, "kaldi-decoder_b200", "csrc", "minifst"),
           "-I", os.path.join(ROOT, "include"), "-I", ROOT]
    for src in ("faster-decoder.cc", "simple-decoder.cc", "fst-io.cc", "decodable-ctc.cc"):
        r = subprocess.run([cxx, "-std=c++17", "-fsyntax-only", *inc,
                            os.path.join(ROOT, "kaldi-decoder_b200", "csrc", src)],
                           capture_output=True, text=True)
        assert r.returncode == 0, src + "\n" + r.stderr[-3000:]
    bad = tmp_path / "bad.cc"
    bad.write_text('#include "fst/fst.h"\n'
                   "int f(const fst::Fst<fst::StdArc> &g) { return g.NumStates(); }\n")
    r = subprocess.run([cxx, "-std=c++17", "-fsyntax-only", *inc, str(bad)], capture_output=True,
                       text=True)
    assert r.returncode != 0 and "NumStates" in r.stderr


def test_dlpack_ingest_parses_the_capsule_and_rejects_host_arrays():
    """kaldi_decoder._from_dlpack: the DLManagedTensor is read through ctypes, the capsule is
    consumed exactly once (no double free when the capsule object dies afterwards)."""
    torch = pytest.importorskip("torch")
    import gc
    import kaldi_decoder as kd

    class W:
        def __init__(self, t):
            self.t = t

        def __dlpack__(self, stream=None):
            return self.t.__dlpack__()

    for _ in range(50):
        with pytest.raises(ValueError, match="CUDA"):
            kd._from_dlpack(W(torch.zeros(3, 4)))
    gc.collect()


def test_cpp_caller_written_against_the_reference_headers_builds(tmp_path):
    """tests/cpp/drop_in.cc includes "kaldi-decoder/csrc/faster-decoder.h" and uses the
    reference's C++ API (faster-decoder.h:65-107); it must compile and link against the B200
    implementation, and -- without a GPU -- fail with the library's loud error, not a crash."""
    import subprocess
    from common import build_cpp_drop_in
    exe = build_cpp_drop_in(tmp_path)
    if exe is None:
        pytest.skip("no g++")
    gc = GoldenCase("h20_default")
    g = gc.graph
    import kaldi_decoder as kd
    fst = kd.StdVectorFst.from_arrays(g.num_states, g.start, g.row_off, g.ilabel, g.olabel,
                                      g.weight, g.nextstate, g.final)
    fst.write(str(tmp_path / "g.fst"))
    lp = np.stack([gc.logp(u) for u in range(gc.n_utts)]).astype(np.float32)
    lp.tofile(str(tmp_path / "lp.bin"))
    o = gc.opts
    r = subprocess.run([exe, str(tmp_path / "g.fst"), str(o["beam"]), str(o["max_active"]),
                        str(o["min_active"]), str(tmp_path / "lp.bin"), str(gc.n_utts),
                        str(gc.T), str(lp.shape[2])], capture_output=True, text=True, timeout=300)
    if capi.device_count() == 0:
        assert r.returncode == 1 and "no CPU fallback" in r.stderr, (r.returncode, r.stderr[-500:])
    else:
        assert r.returncode == 0, r.stderr[-2000:]
