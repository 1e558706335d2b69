"""Seeded synthetic workloads for the token-passing search.

There is no network, no dataset and no checkpoint in the build environment, so
every graph and every log-prob matrix used by the tests and by ``bench.py`` is
generated here from a seed:

* ``make_h``      -- CTC topology H over V tokens (V states, V*V arcs, no
                     epsilon-input arcs).
* ``make_hl``     -- H o L-like graph: prefix trie of a synthetic lexicon with
                     a token state and a blank state per trie node, word-end
                     epsilon-input arcs back to the root.
* ``make_hlg``    -- the HL construction replicated per n-gram history of a
                     synthetic back-off LM (word-end arcs go to the successor
                     history, one epsilon back-off arc per history).
* ``make_logprobs`` -- peaky CTC posteriors ``N(0, sigma) + peak * onehot``,
                     log-softmax in float32, for a word sequence sampled from
                     the graph's own LM.

Labels follow the icefall convention the reference's DecodableCtc assumes
(decodable-ctc.cc:23-28): ilabel = token id + 1, so blank (token 0) is ilabel
1 and ilabel 0 is epsilon.  olabel 0 is epsilon, words are olabel = word + 1.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np

INF = np.float32(np.inf)


@dataclass
class Graph:
    """A decoding graph in CSR form, arcs in their original per-state order."""

    num_states: int
    start: int
    row_off: np.ndarray  # int64 [S+1]
    ilabel: np.ndarray  # int32 [E]
    olabel: np.ndarray  # int32 [E]
    weight: np.ndarray  # float32 [E]
    nextstate: np.ndarray  # int32 [E]
    final: np.ndarray  # float32 [S], +inf = not final
    lm: Optional[dict] = field(default=None, repr=False)
    name: str = ""

    @property
    def num_arcs(self) -> int:
        return int(self.ilabel.shape[0])

    def stats(self) -> Dict[str, float]:
        deg = np.diff(self.row_off)
        n_eps = int((self.ilabel == 0).sum())
        return {
            "name": self.name,
            "states": int(self.num_states),
            "arcs": self.num_arcs,
            "eps_arcs": n_eps,
            "eps_frac": n_eps / max(1, self.num_arcs),
            "final_states": int(np.isfinite(self.final).sum()),
            "max_out_degree": int(deg.max()) if deg.size else 0,
            "mean_out_degree": float(deg.mean()) if deg.size else 0.0,
            "max_ilabel": int(self.ilabel.max()) if self.num_arcs else 0,
        }

    def validate(self) -> None:
        assert self.row_off.dtype == np.int64 and self.row_off.shape == (self.num_states + 1,)
        assert self.row_off[0] == 0 and self.row_off[-1] == self.num_arcs
        assert np.all(np.diff(self.row_off) >= 0)
        for a, dt in ((self.ilabel, np.int32), (self.olabel, np.int32),
                      (self.weight, np.float32), (self.nextstate, np.int32)):
            assert a.dtype == dt and a.shape == (self.num_arcs,)
        assert self.final.dtype == np.float32 and self.final.shape == (self.num_states,)
        if self.num_arcs:
            assert self.nextstate.min() >= 0 and self.nextstate.max() < self.num_states
            assert self.ilabel.min() >= 0 and self.olabel.min() >= 0
        assert 0 <= self.start < self.num_states


def graph_from_arcs(num_states: int, start: int, src, ilabel, olabel, weight, nextstate,
                    final, order_key=None, name: str = "", lm=None) -> Graph:
    """Builds the CSR from an unordered arc list.  Arcs of one state are ordered
    by ``order_key`` (default: the order given)."""
    src = np.asarray(src, dtype=np.int64)
    n = src.shape[0]
    if order_key is None:
        order_key = np.arange(n, dtype=np.int64)
    perm = np.lexsort((np.asarray(order_key), src))
    src = src[perm]
    row_off = np.zeros(num_states + 1, dtype=np.int64)
    np.add.at(row_off, src + 1, 1)
    row_off = np.cumsum(row_off)
    g = Graph(
        num_states=int(num_states),
        start=int(start),
        row_off=row_off,
        ilabel=np.ascontiguousarray(np.asarray(ilabel, dtype=np.int32)[perm]),
        olabel=np.ascontiguousarray(np.asarray(olabel, dtype=np.int32)[perm]),
        weight=np.ascontiguousarray(np.asarray(weight, dtype=np.float32)[perm]),
        nextstate=np.ascontiguousarray(np.asarray(nextstate, dtype=np.int32)[perm]),
        final=np.ascontiguousarray(np.asarray(final, dtype=np.float32)),
        lm=lm,
        name=name,
    )
    g.validate()
    return g


# --------------------------------------------------------------------------- H

def make_h(vocab: int = 500) -> Graph:
    """Standard CTC topology: state i = "last frame emitted token i"; arc i->j
    for every pair, ilabel j+1, olabel j when a new non-blank token starts."""
    V = int(vocab)
    i = np.repeat(np.arange(V, dtype=np.int64), V)
    j = np.tile(np.arange(V, dtype=np.int64), V)
    olabel = np.where((j != i) & (j != 0), j, 0)
    return graph_from_arcs(
        V, 0, i, j + 1, olabel, np.zeros(V * V, np.float32), j,
        np.zeros(V, np.float32), name=f"H-{V}",
        lm={"kind": "h", "vocab": V})


def make_random_fst(num_states: int = 200, num_arcs: int = 2000, vocab: int = 30,
                    eps_frac: float = 0.15, seed: int = 0, neg_weight_frac: float = 0.1,
                    dense_states: int = 4) -> Graph:
    """Unstructured random FST for fuzzing: random weights (some negative), several arcs
    with the same ilabel out of one state (non-deterministic), epsilon-input arcs only
    from lower to higher state ids (no epsilon cycles), a few states with a large fan-out
    of distinct labels (label-table candidates), random final states."""
    rng = np.random.default_rng(np.random.PCG64(seed))
    S, V = int(num_states), int(vocab)
    src = rng.integers(0, S, size=num_arcs)
    dst = rng.integers(0, S, size=num_arcs)
    il = rng.integers(1, V + 1, size=num_arcs)
    is_eps = rng.random(num_arcs) < eps_frac
    # epsilon arcs go strictly upwards
    lo, hi = np.minimum(src, dst), np.maximum(src, dst)
    up = is_eps & (lo != hi)
    src = np.where(up, lo, src)
    dst = np.where(up, hi, dst)
    is_eps = up
    il = np.where(is_eps, 0, il)
    ol = np.where(rng.random(num_arcs) < 0.3, rng.integers(1, 50, size=num_arcs), 0)
    w = rng.uniform(0.0, 3.0, size=num_arcs).astype(np.float32)
    neg = rng.random(num_arcs) < neg_weight_frac
    w = np.where(neg & ~is_eps, -w * 0.3, w).astype(np.float32)
    w = np.where(rng.random(num_arcs) < 0.2, np.float32(0.0), w).astype(np.float32)
    # dense states: exactly one emitting arc per label (their random emitting arcs are dropped)
    ds = rng.choice(S, size=min(dense_states, S), replace=False)
    keep = ~(np.isin(src, ds) & (il != 0))
    src, dst, il, ol, w = src[keep], dst[keep], il[keep], ol[keep], w[keep]
    d_src = np.repeat(ds, V)
    d_il = np.tile(np.arange(1, V + 1), ds.shape[0])
    d_dst = rng.integers(0, S, size=d_src.shape[0])
    d_w = rng.uniform(0.0, 2.0, size=d_src.shape[0]).astype(np.float32)
    src = np.concatenate([src, d_src]); dst = np.concatenate([dst, d_dst])
    il = np.concatenate([il, d_il]); ol = np.concatenate([ol, np.zeros(d_src.shape[0], np.int64)])
    w = np.concatenate([w, d_w])
    final = np.where(rng.random(S) < 0.2, rng.uniform(0, 2, size=S), np.inf).astype(np.float32)
    return graph_from_arcs(S, 0, src, il, ol, w, dst, final, name=f"random-{S}-{seed}",
                           lm={"kind": "h", "vocab": V})


# ----------------------------------------------------------------- lexicon, LM

def make_lexicon(n_words: int, vocab: int, rng: np.random.Generator,
                 max_len: int = 6) -> Tuple[np.ndarray, np.ndarray]:
    """Unique spellings of 1..max_len non-blank tokens.  Returns (spell, lens):
    spell int32 [n_words, max_len] padded with 0, lens int32 [n_words]."""
    V = int(vocab)
    len_p = np.array([0.01, 0.09, 0.25, 0.30, 0.20, 0.15][:max_len], dtype=np.float64)
    len_p /= len_p.sum()
    spell = np.zeros((0, max_len), dtype=np.int64)
    while spell.shape[0] < n_words:
        m = int((n_words - spell.shape[0]) * 1.3) + 64
        lens = rng.choice(np.arange(1, max_len + 1), size=m, p=len_p)
        toks = rng.integers(1, V, size=(m, max_len), dtype=np.int64)
        toks[np.arange(max_len)[None, :] >= lens[:, None]] = 0
        spell = np.concatenate([spell, toks], axis=0)
        key = np.zeros(spell.shape[0], dtype=np.int64)
        for d in range(max_len):
            key = key * V + spell[:, d]
        _, first = np.unique(key, return_index=True)
        spell = spell[np.sort(first)]
    spell = spell[:n_words]
    lens = (spell != 0).sum(axis=1)
    return spell.astype(np.int32), lens.astype(np.int32)


def _zipf_probs(n: int, s: float = 1.0) -> np.ndarray:
    p = 1.0 / np.power(np.arange(1, n + 1, dtype=np.float64), s)
    return p / p.sum()


def _pair_lookup(keys_sorted: np.ndarray, ids_sorted: np.ndarray, q: np.ndarray) -> np.ndarray:
    """ids for the queries present in keys_sorted, -1 otherwise."""
    if keys_sorted.size == 0:
        return np.full(q.shape, -1, dtype=np.int64)
    pos = np.searchsorted(keys_sorted, q)
    pos = np.minimum(pos, keys_sorted.size - 1)
    hit = keys_sorted[pos] == q
    return np.where(hit, ids_sorted[pos], -1)


def _make_lm(n_words: int, orders, rng: np.random.Generator, succ_range=(5, 50)) -> dict:
    """Synthetic back-off LM.  ``orders`` = number of histories per order above
    unigram, e.g. (n_bigram_hist, n_trigram_hist[, n_4gram_hist]).

    History 0 is the unigram state.  A history is a tuple of up to len(orders)
    words; its entries are (word, cost, destination history); it backs off to
    the history with the oldest word dropped (or further down if that one does
    not exist).
    """
    W = int(n_words)
    uni_p = _zipf_probs(W)
    lo, hi = succ_range

    # hist_words[h] = tuple-coded context; level 0 = unigram.
    hist_ctx = [()]  # python tuples only for the *histories* (10^3..10^5), not entries
    hist_level = [0]
    level_ids = {0: np.array([0], dtype=np.int64)}
    # entries of already-built levels, used to draw the next level's contexts
    ent_root, ent_word, ent_p = [], [], []

    # unigram entries: every word
    ent_root.append(np.zeros(W, dtype=np.int64))
    ent_word.append(np.arange(W, dtype=np.int64))
    ent_p.append(uni_p * 1.0)

    prev_level_entries = (np.zeros(W, dtype=np.int64), np.arange(W, dtype=np.int64), uni_p)
    ctx_arrays = {0: np.zeros((1, 0), dtype=np.int64)}  # level -> [n_hist, level] words
    n_hist = 1
    for level, n_h in enumerate(orders, start=1):
        # a level-`level` context = (context of a level-1 lower history) + one of its successors
        pr, pw, pp = prev_level_entries
        n_h = int(min(n_h, pr.shape[0]))
        pick = rng.choice(pr.shape[0], size=n_h, replace=False, p=pp / pp.sum())
        pick.sort()
        lower_ctx = ctx_arrays[level - 1]
        base = level_ids[level - 1][0]
        ctx = np.concatenate([lower_ctx[pr[pick] - base], pw[pick][:, None]], axis=1)
        ids = n_hist + np.arange(n_h, dtype=np.int64)
        ctx_arrays[level] = ctx
        level_ids[level] = ids
        n_hist += n_h
        # successors
        k = rng.integers(lo, hi + 1, size=n_h)
        tot = int(k.sum())
        r = np.repeat(ids, k)
        # Zipf-distributed successor words; duplicates inside one history removed
        w = rng.choice(W, size=tot, p=uni_p)
        key = r * W + w
        _, first = np.unique(key, return_index=True)
        first.sort()
        r, w = r[first], w[first]
        p = uni_p[w] * rng.uniform(0.5, 2.0, size=w.shape[0])
        ent_root.append(r)
        ent_word.append(w)
        ent_p.append(p)
        prev_level_entries = (r, w, p)

    ent_root = np.concatenate(ent_root)
    ent_word = np.concatenate(ent_word)
    ent_p = np.concatenate(ent_p)
    n_levels = len(orders)

    # normalise per history; histories above unigram keep `1 - bo_mass` for explicit words
    bo_mass = np.zeros(n_hist, dtype=np.float64)
    bo_mass[1:] = rng.uniform(0.1, 0.5, size=n_hist - 1)
    sums = np.zeros(n_hist, dtype=np.float64)
    np.add.at(sums, ent_root, ent_p)
    ent_p = ent_p / sums[ent_root] * (1.0 - bo_mass[ent_root])
    ent_cost = (-np.log(ent_p)).astype(np.float32)
    bo_cost = np.zeros(n_hist, dtype=np.float32)
    bo_cost[1:] = (-np.log(bo_mass[1:])).astype(np.float32)

    # lookup tables: context (as base-W number, with length) -> history id
    def code(ctx: np.ndarray) -> np.ndarray:
        c = np.zeros(ctx.shape[0], dtype=np.int64)
        for d in range(ctx.shape[1]):
            c = c * W + ctx[:, d]
        return c

    tables = {}
    for level in range(1, n_levels + 1):
        c = code(ctx_arrays[level])
        o = np.argsort(c, kind="stable")
        tables[level] = (c[o], level_ids[level][o])

    hist_lvl = np.zeros(n_hist, dtype=np.int64)
    hist_ctx_arr = np.zeros((n_hist, max(1, n_levels)), dtype=np.int64)  # right-aligned
    for level in range(1, n_levels + 1):
        ids = level_ids[level]
        hist_lvl[ids] = level
        hist_ctx_arr[ids, n_levels - level:] = ctx_arrays[level]

    def longest_history(ctx_full: np.ndarray, ctx_len: np.ndarray) -> np.ndarray:
        """ctx_full [n, n_levels] right-aligned words, ctx_len[n] valid count.
        Returns the id of the longest existing history that is a suffix."""
        out = np.zeros(ctx_full.shape[0], dtype=np.int64)
        done = np.zeros(ctx_full.shape[0], dtype=bool)
        for level in range(n_levels, 0, -1):
            cand = (~done) & (ctx_len >= level)
            if not cand.any():
                continue
            c = code(ctx_full[cand][:, n_levels - level:])
            ks, vs = tables[level]
            hid = _pair_lookup(ks, vs, c)
            idx = np.nonzero(cand)[0]
            ok = hid >= 0
            out[idx[ok]] = hid[ok]
            done[idx[ok]] = True
        return out

    # destination of entry (h, w): longest existing suffix of ctx(h) + w
    if n_levels > 0:
        full = np.concatenate([hist_ctx_arr[ent_root][:, 1:], ent_word[:, None]], axis=1) \
            if n_levels > 1 else ent_word[:, None]
        flen = np.minimum(hist_lvl[ent_root] + 1, n_levels)
        ent_dest = longest_history(full, flen)
        # back-off target of h: longest existing suffix of ctx(h) minus its oldest word
        bfull = hist_ctx_arr.copy()
        blen = np.maximum(hist_lvl - 1, 0)
        # drop oldest word: the valid words are right-aligned, so just shorten the length
        bo_dest = longest_history(bfull, blen)
        bo_dest[0] = -1
    else:
        ent_dest = np.zeros(ent_root.shape[0], dtype=np.int64)
        bo_dest = np.full(n_hist, -1, dtype=np.int64)

    order = np.lexsort((ent_word, ent_root))
    ent_root, ent_word, ent_cost, ent_dest, ent_p = (
        ent_root[order], ent_word[order], ent_cost[order], ent_dest[order], ent_p[order])
    off = np.zeros(n_hist + 1, dtype=np.int64)
    np.add.at(off, ent_root + 1, 1)
    off = np.cumsum(off)
    return {
        "kind": "lm", "n_words": W, "n_hist": int(n_hist),
        "ent_root": ent_root, "ent_word": ent_word, "ent_cost": ent_cost,
        "ent_dest": ent_dest, "ent_p": ent_p, "ent_off": off,
        "bo_dest": bo_dest, "bo_cost": bo_cost, "bo_mass": bo_mass,
    }


def _trie_graph(lm: dict, spell: np.ndarray, lens: np.ndarray, vocab: int,
                rng: np.random.Generator, name: str) -> Graph:
    """One prefix trie per LM history (root state = history id)."""
    V = int(vocab)
    n_roots = lm["n_hist"]
    e_root, e_word = lm["ent_root"], lm["ent_word"]
    e_cost, e_dest = lm["ent_cost"], lm["ent_dest"]
    E = e_root.shape[0]
    max_len = spell.shape[1]
    e_len = lens[e_word].astype(np.int64)

    node_parent, node_tok = [], []
    n_nodes = 0
    alive = np.arange(E, dtype=np.int64)
    cur = e_root.astype(np.int64).copy()  # combined id: < n_roots root, else n_roots + node
    end_node = np.full(E, -1, dtype=np.int64)
    for d in range(max_len):
        sel = e_len[alive] > d
        alive, cur = alive[sel], cur[sel]
        if alive.size == 0:
            break
        tok = spell[e_word[alive], d].astype(np.int64)
        key = cur * V + tok
        uniq, inv = np.unique(key, return_inverse=True)
        ids = n_roots + n_nodes + np.arange(uniq.shape[0], dtype=np.int64)
        node_parent.append(uniq // V)
        node_tok.append(uniq % V)
        cur = ids[inv]
        ends = e_len[alive] == d + 1
        end_node[alive[ends]] = cur[ends]
        n_nodes += uniq.shape[0]
    node_parent = np.concatenate(node_parent)
    node_tok = np.concatenate(node_tok)
    assert (end_node >= 0).all()

    def t_state(cid):  # combined node id -> token state
        return n_roots + 2 * (cid - n_roots)

    def b_state(cid):
        return n_roots + 2 * (cid - n_roots) + 1

    nid = n_roots + np.arange(n_nodes, dtype=np.int64)
    T, B = t_state(nid), b_state(nid)
    il_tok = node_tok + 1
    par_is_root = node_parent < n_roots
    par_tok = np.where(par_is_root, -1, node_tok[np.maximum(node_parent - n_roots, 0)])

    src, il, ol, w, dst, rank = [], [], [], [], [], []

    def add(s, i, o, ww, d, r):
        s = np.asarray(s, dtype=np.int64)
        n = s.shape[0]
        src.append(s)
        il.append(np.broadcast_to(np.asarray(i, dtype=np.int64), (n,)))
        ol.append(np.broadcast_to(np.asarray(o, dtype=np.int64), (n,)))
        w.append(np.broadcast_to(np.asarray(ww, dtype=np.float32), (n,)))
        dst.append(np.asarray(d, dtype=np.int64))
        rank.append(np.full(n, r, dtype=np.int64))

    roots = np.arange(n_roots, dtype=np.int64)
    add(roots, 1, 0, 0.0, roots, 0)                      # root blank self-loop
    add(T, il_tok, 0, 0.0, T, 0)                         # token self-loop
    add(T, 1, 0, 0.0, B, 1)                              # token -> blank
    add(B, 1, 0, 0.0, B, 0)                              # blank self-loop
    m = par_is_root
    add(node_parent[m], il_tok[m], 0, 0.0, T[m], 2)      # root -> first token
    m2 = (~par_is_root) & (par_tok != node_tok)
    add(t_state(node_parent[m2]), il_tok[m2], 0, 0.0, T[m2], 2)   # token -> next token
    m3 = ~par_is_root
    add(b_state(node_parent[m3]), il_tok[m3], 0, 0.0, T[m3], 2)   # blank -> next token
    add(t_state(end_node), 0, e_word + 1, e_cost, e_dest, 3)      # word end (from token)
    add(b_state(end_node), 0, e_word + 1, e_cost, e_dest, 3)      # word end (from blank)
    has_bo = lm["bo_dest"] >= 0
    add(roots[has_bo], 0, 0, lm["bo_cost"][has_bo], lm["bo_dest"][has_bo], 4)  # back-off

    src = np.concatenate(src)
    il = np.concatenate(il)
    ol = np.concatenate(ol)
    w = np.concatenate(w)
    dst = np.concatenate(dst)
    rank = np.concatenate(rank)
    order_key = (rank * (V + 2) + il) * (lm["n_words"] + 2) + ol
    S = n_roots + 2 * n_nodes
    final = np.full(S, np.inf, dtype=np.float32)
    final[:n_roots] = rng.uniform(0.0, 2.0, size=n_roots).astype(np.float32)
    lm = dict(lm)
    lm.update({"spell": spell, "lens": lens, "vocab": V})
    return graph_from_arcs(S, 0, src, il, ol, w, dst, final, order_key=order_key,
                           name=name, lm=lm)


def make_hl(n_words: int = 200_000, vocab: int = 500, seed: int = 1) -> Graph:
    rng = np.random.default_rng(np.random.PCG64(seed))
    spell, lens = make_lexicon(n_words, vocab, rng)
    lm = _make_lm(n_words, (), rng)
    return _trie_graph(lm, spell, lens, vocab, rng, name=f"HL-{n_words}")


def make_hlg(n_words: int = 50_000, orders=(2_500, 5_000), vocab: int = 500,
             seed: int = 2, succ_range=(5, 50), name: Optional[str] = None) -> Graph:
    rng = np.random.default_rng(np.random.PCG64(seed))
    spell, lens = make_lexicon(n_words, vocab, rng)
    lm = _make_lm(n_words, tuple(orders), rng, succ_range=succ_range)
    return _trie_graph(lm, spell, lens, vocab, rng,
                       name=name or f"HLG-{len(orders) + 1}g-{n_words}")


# ------------------------------------------------------------------- log-probs

def sample_words(g: Graph, rng: np.random.Generator, n: int) -> np.ndarray:
    lm = g.lm
    off, p, dest, word = lm["ent_off"], lm["ent_p"], lm["ent_dest"], lm["ent_word"]
    out = np.empty(n, dtype=np.int64)
    h = 0
    for i in range(n):
        while h != 0 and rng.random() < lm["bo_mass"][h]:
            h = int(lm["bo_dest"][h])
        a, b = int(off[h]), int(off[h + 1])
        cdf = _cdf_cache(lm, h, a, b, p)
        j = a + int(np.searchsorted(cdf, rng.random() * cdf[-1]))
        j = min(j, b - 1)
        out[i] = word[j]
        h = int(dest[j])
    return out


def _cdf_cache(lm, h, a, b, p):
    cache = lm.setdefault("_cdf", {})
    c = cache.get(h)
    if c is None:
        c = np.cumsum(p[a:b])
        if len(cache) > 4096 and h != 0:
            return c
        cache[h] = c
    return c


def make_alignment(g: Graph, T: int, rng: np.random.Generator) -> np.ndarray:
    """A CTC frame alignment (token id per frame, 0 = blank) of length T."""
    lm = g.lm
    V = lm["vocab"]
    ali = []
    prev = 0
    if lm["kind"] == "h":
        units = [np.array([t]) for t in rng.integers(1, V, size=T)]
    else:
        words = sample_words(g, rng, max(4, T // 3))
        units = [lm["spell"][w, :lm["lens"][w]] for w in words]
    for unit in units:
        piece = []
        p = prev
        for tok in unit:
            tok = int(tok)
            nb = int(rng.integers(0, 4))
            if tok == p and nb == 0:
                nb = 1
            piece.extend([0] * nb)
            piece.append(tok)
            if rng.random() < 0.25:
                piece.append(tok)
            p = tok
        if len(ali) + len(piece) > T:
            break
        ali.extend(piece)
        prev = p
    ali.extend([0] * (T - len(ali)))
    return np.asarray(ali[:T], dtype=np.int64)


def make_logprobs(g: Graph, T: int, seed: int, peak: float = 12.0, sigma: float = 1.0,
                  vocab: Optional[int] = None) -> np.ndarray:
    """float32 [T, V] log-softmax of N(0, sigma) + peak * onehot(alignment)."""
    V = int(vocab or g.lm["vocab"])
    rng = np.random.default_rng(np.random.PCG64(seed))
    ali = make_alignment(g, T, rng)
    x = rng.standard_normal((T, V), dtype=np.float32)
    if sigma != 1.0:
        x *= np.float32(sigma)
    x[np.arange(T), ali] += np.float32(peak)
    m = x.max(axis=1, keepdims=True)
    x -= m
    lse = np.log(np.exp(x).sum(axis=1, keepdims=True, dtype=np.float32))
    x -= lse
    return x


def make_batch(g: Graph, n_utts: int, T: int, seed: int, peak: float = 12.0,
               sigma: float = 1.0, out: Optional[np.ndarray] = None) -> np.ndarray:
    """float32 [n_utts, T, V]; utterance u uses seed * 1_000_003 + u."""
    V = int(g.lm["vocab"])
    if out is None:
        out = np.empty((n_utts, T, V), dtype=np.float32)
    for u in range(n_utts):
        out[u] = make_logprobs(g, T, seed * 1_000_003 + u, peak=peak, sigma=sigma)
    return out


# ----------------------------------------------------------- BASELINE configs

def make_config_graph(config: str) -> Graph:
    """The graphs of BASELINE.json's configs (C1..C4; C5 reuses C3)."""
    if config == "C1":
        return make_h(500)
    if config == "C2":
        return make_hl(200_000, 500, seed=1)
    if config in ("C3", "C5"):
        return make_hlg(50_000, (2_500, 5_000), 500, seed=2, name="HLG-3g")
    if config == "C4":
        return make_hlg(200_000, (20_000, 80_000, 160_000), 500, seed=3, name="HLG-4g")
    raise ValueError(f"unknown config {config}")
