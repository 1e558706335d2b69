// oracle/kd_oracle.cc  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A CPU restatement of the reference's token-passing Viterbi beam search
// (FasterDecoder), written from its behaviour, in two modes:
//
//   mode 0 "reference order": reproduces the reference bit for bit, including
//          everything that depends on its processing order -- the HashList
//          iteration order (hash-list-inl.h:37-51,127-173), the running
//          next-frame cutoff of ProcessEmitting (faster-decoder.cc:172-217),
//          first-arrival-wins recombination (faster-decoder.cc:219-228 with
//          faster-decoder.h:141-143) and the LIFO worklist of
//          ProcessNonemitting (faster-decoder.cc:59-119).  It is pinned
//          against oracle/_ref (the unmodified reference object code) by
//          tests/test_oracle_vs_reference.py, token list by token list.
//
//   mode 1 "canonical": the order-independent statement of the same search that
//          the CUDA kernels implement: exact GetCutoff,
//          admit emitting arcs with new_weight < C* where
//          C* = min(new_weight) + adaptive_beam (the final value of the
//          reference's running cutoff), recombine by (cost, arc index),
//          epsilon closure to the fixed point under `<= C*`.  It equals mode 0
//          whenever the reference admits no "extra" tokens that later matter
//          and no exact cost tie decides a backpointer.
//
//   mode 2 "simple": the order-independent statement of the reference's SimpleDecoder
//          (simple-decoder.cc:29-41, 150-279; simple-decoder.h:89-101): beam-only
//          pruning; an emitting arc is admitted when (cost + w) + ac < C*, C* =
//          min of that + beam (simple-decoder.cc:168-176), but the token stores
//          cost + float(w + ac) (simple-decoder.h:96); the closure runs under
//          best stored cost + beam with `>` (simple-decoder.cc:196-215) and expands
//          every token; PruneToks (cc:251-279, `cost < best + beam`) is applied to
//          what the accessors return and to what the next frame expands.  Pinned
//          against the unmodified simple-decoder.cc in oracle/_ref
//          (tests/test_oracle_vs_reference.py).
//
// Both modes carry counters (tokens, arcs, extras, ties, binding frames) used
// for the algorithmic-bytes figure and the parity report.
//
// PARITY PIN: the reference ships no decoder test or golden vector
// (SURVEY.md §4), so this file is pinned against the compiled reference
// itself, and against tests/golden/*.npz generated from it.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
// load the library built from this file.

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace kdo {

constexpr double kInf = std::numeric_limits<double>::infinity();

struct Graph {
  int32_t num_states = 0;
  int32_t start = -1;
  std::vector<int64_t> off;
  std::vector<int32_t> il, ol, ns;
  std::vector<float> w, fin;  // fin = +inf for non-final states
};

struct Opts {
  float beam = 16.0f;
  int32_t max_active = std::numeric_limits<int32_t>::max();
  int32_t min_active = 20;
  float beam_delta = 0.5f;
  float hash_ratio = 2.0f;
};

struct Stats {
  int64_t frames = 0;
  int64_t tokens_in = 0;        // sum over frames of tokens at frame start
  int64_t tokens_expanded = 0;  // tokens with cost < weight_cutoff
  int64_t emit_arcs = 0;        // emitting arcs visited (A_emit)
  int64_t eps_arcs = 0;         // epsilon arcs visited in the closure (A_eps)
  int64_t admitted = 0;         // emitting arcs that passed the pruning test
  int64_t extras = 0;           // mode 0: admitted with new_weight >= final C*
  int64_t emit_ties = 0;        // equal-cost recombinations, emitting phase
  int64_t eps_ties = 0;         // equal-cost recombinations, epsilon phase
  int64_t tokens_out = 0;       // sum over frames of tokens at frame end
  int64_t max_tokens = 0;
  int64_t binding_max = 0;      // frames where GetCutoff returned via max_active
  int64_t binding_min = 0;      // frames where GetCutoff returned via min_active
  int64_t first_binding_frame = -1;
  int64_t all_arcs_scanned = 0;  // arcs the reference touches incl. skipped ones
};

struct PathArc {
  int32_t il, ol;
  float graph, ac;
};

// The state -> token map whose iteration order the reference's pruning and
// tie-breaking depend on (HashList, hash-list-inl.h).  Semantics restated:
// a key lives in bucket key % size; the iteration list is the concatenation
// of the buckets in order of first occupation, each bucket in insertion
// order; the bucket count only grows and survives Clear().
class OrderedStateMap {
 public:
  struct Cell {
    int32_t state;
    int32_t tok;
  };

  void SetSize(size_t n) {
    size_ = n;
    if (n > head_.size()) {
      head_.resize(n, -1);
      tail_.resize(n, -1);
    }
  }
  size_t Size() const { return size_; }

  // Returns the cell index holding `state`, inserting (state, tok) if absent.
  int32_t Insert(int32_t state, int32_t tok) {
    size_t b = static_cast<size_t>(state) % size_;
    for (int32_t c = head_[b]; c >= 0; c = next_[c]) {
      if (cells_[c].state == state) return c;
    }
    int32_t c = static_cast<int32_t>(cells_.size());
    cells_.push_back({state, tok});
    next_.push_back(-1);
    if (head_[b] < 0) {
      head_[b] = c;
      order_.push_back(b);
    } else {
      next_[tail_[b]] = c;
    }
    tail_[b] = c;
    return c;
  }

  Cell &At(int32_t c) { return cells_[c]; }
  const Cell &At(int32_t c) const { return cells_[c]; }
  size_t Count() const { return cells_.size(); }

  // Cell indices in iteration order.
  void ListOrder(std::vector<int32_t> *out) const {
    out->clear();
    out->reserve(cells_.size());
    for (size_t b : order_) {
      for (int32_t c = head_[b]; c >= 0; c = next_[c]) out->push_back(c);
    }
  }

  // Hands the (state, tok) pairs over in iteration order and empties the map.
  void Take(std::vector<Cell> *out) {
    out->clear();
    out->reserve(cells_.size());
    for (size_t b : order_) {
      for (int32_t c = head_[b]; c >= 0; c = next_[c]) out->push_back(cells_[c]);
      head_[b] = tail_[b] = -1;
    }
    order_.clear();
    cells_.clear();
    next_.clear();
  }

 private:
  size_t size_ = 0;
  std::vector<int32_t> head_, tail_;
  std::vector<size_t> order_;
  std::vector<Cell> cells_;
  std::vector<int32_t> next_;
};

class Decoder {
 public:
  Decoder(const Graph *g, const Opts &o, int mode) : g_(g), o_(o), mode_(mode) {
    if (!(o.hash_ratio >= 1.0f)) throw std::runtime_error("hash_ratio >= 1.0");
    if (!(o.max_active > 1)) throw std::runtime_error("max_active > 1");
    if (!(o.min_active >= 0 && o.min_active < o.max_active))
      throw std::runtime_error("0 <= min_active < max_active");
    map_.SetSize(1000);  // faster-decoder.cc:31
  }

  void SetOptions(const Opts &o) { o_ = o; }

  // faster-decoder.cc:42-56
  void Init() {
    ReleaseAll();
    if (g_->start < 0) throw std::runtime_error("graph has no start state");
    int32_t t = NewTok(-1, -1, 0.0);
    map_.Insert(g_->start, t);
    // simple-decoder.cc:29-41: ProcessNonemitting's own cutoff, best (0) + beam
    Closure(mode_ == 2 ? 0.0 + o_.beam : static_cast<double>(std::numeric_limits<float>::max()));
    frames_ = 0;
    stats_ = Stats();
  }

  // faster-decoder.cc:126-152
  void Advance(const float *p, int32_t rows, int32_t cols, int32_t offset,
               int32_t max_frames) {
    if (frames_ < 0) throw std::runtime_error("Init() before Advance()");
    int32_t ready = offset + rows;
    if (ready < frames_) throw std::runtime_error("frames ready < frames decoded");
    int32_t target = ready;
    if (max_frames >= 0) target = std::min(target, frames_ + max_frames);
    while (frames_ < target) {
      const float *row = p + static_cast<int64_t>(frames_ - offset) * cols;
      double c = mode_ == 0 ? EmitReferenceOrder(row)
                            : (mode_ == 1 ? EmitCanonical(row) : EmitSimple(row));
      Closure(c);
      stats_.frames++;
      int64_t n = static_cast<int64_t>(map_.Count());
      stats_.tokens_out += n;
      stats_.max_tokens = std::max(stats_.max_tokens, n);
    }
  }

  int32_t NumFrames() const { return frames_; }
  const Stats &GetStats() const { return stats_; }

  // faster-decoder.cc:347-354
  bool ReachedFinal() const {
    std::vector<int32_t> order;
    map_.ListOrder(&order);
    DropPruned(&order);
    for (int32_t c : order) {
      const auto &cell = map_.At(c);
      if (toks_[cell.tok].cost != kInf && g_->fin[cell.state] != kInfF()) return true;
    }
    return false;
  }

  // tokens in iteration order (mode 0) or by state id (mode 1)
  void Tokens(std::vector<int32_t> *states, std::vector<double> *costs) const {
    std::vector<int32_t> order;
    IterOrder(&order);
    states->clear();
    costs->clear();
    for (int32_t c : order) {
      states->push_back(map_.At(c).state);
      costs->push_back(toks_[map_.At(c).tok].cost);
    }
  }

  // faster-decoder.cc:356-424.  `raw` skips the RemoveEpsLocal-style merge and
  // returns one arc per token.
  bool BestPath(bool use_final_probs, bool raw, std::vector<PathArc> *out,
                float final2[2]) const {
    out->clear();
    final2[0] = final2[1] = 0.0f;
    std::vector<int32_t> order;
    IterOrder(&order);
    bool is_final = ReachedFinal();
    int32_t best = -1, best_state = -1;
    if (!is_final) {
      for (int32_t c : order) {
        int32_t t = map_.At(c).tok;
        if (best < 0 || toks_[best].cost > toks_[t].cost) {
          best = t;
          best_state = map_.At(c).state;
        }
      }
    } else {
      double best_cost = kInf;
      for (int32_t c : order) {
        int32_t t = map_.At(c).tok;
        double this_cost =
            toks_[t].cost + static_cast<double>(g_->fin[map_.At(c).state]);
        if (this_cost < best_cost && this_cost != kInf) {
          best_cost = this_cost;
          best = t;
          best_state = map_.At(c).state;
        }
      }
    }
    if (best < 0) return false;
    std::vector<PathArc> rev;
    for (int32_t t = best; t >= 0; t = toks_[t].prev) {
      int32_t p = toks_[t].prev;
      float tot = static_cast<float>(toks_[t].cost - (p >= 0 ? toks_[p].cost : 0.0));
      int32_t a = toks_[t].arc;
      float graph = a >= 0 ? g_->w[a] : 0.0f;
      PathArc pa;
      pa.il = a >= 0 ? g_->il[a] : 0;
      pa.ol = a >= 0 ? g_->ol[a] : 0;
      pa.graph = graph;
      pa.ac = tot - graph;
      rev.push_back(pa);
    }
    rev.pop_back();  // the start token carries no arc (faster-decoder.cc:404-406)
    std::vector<PathArc> arcs(rev.rbegin(), rev.rend());
    if (is_final && use_final_probs) {
      final2[0] = g_->fin[best_state];
      final2[1] = 0.0f;
    }
    if (raw) {
      *out = std::move(arcs);
      return true;
    }
    MergeLinear(arcs, out, final2);
    return true;
  }

  // RemoveEpsLocal restricted to a linear chain (kaldifst, not vendored by the
  // reference; PARITY UNPINNED by reference tests):
  // greedy left-to-right merge of neighbours that do not both carry an ilabel
  // nor both an olabel; a trailing (eps, eps) arc folds into the final weight.
  static void MergeLinear(const std::vector<PathArc> &in, std::vector<PathArc> *out,
                          float final2[2]) {
    out->clear();
    if (in.empty()) return;
    PathArc cur = in[0];
    for (size_t i = 1; i < in.size(); ++i) {
      const PathArc &n = in[i];
      bool both_il = cur.il != 0 && n.il != 0;
      bool both_ol = cur.ol != 0 && n.ol != 0;
      if (!both_il && !both_ol) {
        cur.il = cur.il != 0 ? cur.il : n.il;
        cur.ol = cur.ol != 0 ? cur.ol : n.ol;
        cur.graph = cur.graph + n.graph;
        cur.ac = cur.ac + n.ac;
      } else {
        out->push_back(cur);
        cur = n;
      }
    }
    if (cur.il == 0 && cur.ol == 0) {
      final2[0] = cur.graph + final2[0];
      final2[1] = cur.ac + final2[1];
    } else {
      out->push_back(cur);
    }
  }

 private:
  struct Tok {
    double cost;
    int32_t prev;
    int32_t arc;  // index into the graph's arc arrays, -1 for the start token
    int32_t refs;
  };

  static float kInfF() { return std::numeric_limits<float>::infinity(); }

  // mode 2: the cells PruneToks keeps (simple-decoder.cc:251-279); not applied before the
  // first frame (InitDecoding does not prune)
  void DropPruned(std::vector<int32_t> *order) const {
    if (mode_ != 2 || frames_ <= 0) return;
    double best = kInf;
    for (int32_t c : *order) best = std::min(best, toks_[map_.At(c).tok].cost);
    const double cutoff = best + o_.beam;
    size_t k = 0;
    for (int32_t c : *order)
      if (toks_[map_.At(c).tok].cost < cutoff) (*order)[k++] = c;
    order->resize(k);
  }

  void IterOrder(std::vector<int32_t> *order) const {
    map_.ListOrder(order);
    DropPruned(order);
    if (mode_ != 0) {
      std::sort(order->begin(), order->end(), [this](int32_t a, int32_t b) {
        return map_.At(a).state < map_.At(b).state;
      });
    }
  }

  int32_t NewTok(int32_t arc, int32_t prev, double cost) {
    int32_t t;
    if (!free_.empty()) {
      t = free_.back();
      free_.pop_back();
    } else {
      t = static_cast<int32_t>(toks_.size());
      toks_.push_back(Tok());
    }
    toks_[t] = Tok{cost, prev, arc, 1};
    if (prev >= 0) toks_[prev].refs++;
    return t;
  }

  // drop one reference; frees the chain of tokens nobody points at any more
  void Release(int32_t t) {
    while (t >= 0 && --toks_[t].refs == 0) {
      int32_t p = toks_[t].prev;
      free_.push_back(t);
      t = p;
    }
  }

  void ReleaseAll() {
    std::vector<OrderedStateMap::Cell> cells;
    map_.Take(&cells);
    toks_.clear();
    free_.clear();
  }

  struct Cutoff {
    double weight_cutoff;
    float adaptive_beam;
    int32_t best;  // position in the list, -1 if empty
  };

  // faster-decoder.cc:244-336.  Note the scratch array is float
  // (faster-decoder.h:189), so the order statistics are taken over costs
  // rounded to float.
  Cutoff GetCutoff(const std::vector<OrderedStateMap::Cell> &list) {
    Cutoff r;
    double best_cost = kInf;
    r.best = -1;
    scratch_.clear();
    for (size_t i = 0; i < list.size(); ++i) {
      double c = toks_[list[i].tok].cost;
      scratch_.push_back(static_cast<float>(c));
      if (c < best_cost) {
        best_cost = c;
        r.best = static_cast<int32_t>(i);
      }
    }
    size_t n = scratch_.size();
    if (o_.max_active == std::numeric_limits<int32_t>::max() && o_.min_active == 0) {
      r.adaptive_beam = o_.beam;
      r.weight_cutoff = best_cost + o_.beam;
      return r;
    }
    double beam_cutoff = best_cost + o_.beam;
    double min_cut = kInf, max_cut = kInf;
    size_t maxa = static_cast<size_t>(o_.max_active);
    size_t mina = static_cast<size_t>(o_.min_active);
    if (n > maxa) {
      std::nth_element(scratch_.begin(), scratch_.begin() + maxa, scratch_.end());
      max_cut = scratch_[maxa];
    }
    if (max_cut < beam_cutoff) {
      r.adaptive_beam = static_cast<float>(max_cut - best_cost + o_.beam_delta);
      r.weight_cutoff = max_cut;
      NoteBinding(&stats_.binding_max);
      return r;
    }
    if (n > mina) {
      if (mina == 0) {
        min_cut = best_cost;
      } else {
        std::nth_element(scratch_.begin(), scratch_.begin() + mina,
                         n > maxa ? scratch_.begin() + maxa : scratch_.end());
        min_cut = scratch_[mina];
      }
    }
    if (min_cut > beam_cutoff) {
      r.adaptive_beam = static_cast<float>(min_cut - best_cost + o_.beam_delta);
      r.weight_cutoff = min_cut;
      NoteBinding(&stats_.binding_min);
      return r;
    }
    r.adaptive_beam = o_.beam;
    r.weight_cutoff = beam_cutoff;
    return r;
  }

  void NoteBinding(int64_t *counter) {
    (*counter)++;
    if (stats_.first_binding_frame < 0) stats_.first_binding_frame = frames_;
  }

  // faster-decoder.cc:338-345
  void MaybeGrow(size_t n) {
    size_t want = static_cast<size_t>(static_cast<float>(n) * o_.hash_ratio);
    if (want > map_.Size()) map_.SetSize(want);
  }

  static inline double ArcCost(float w, double cost, float ac) {
    return static_cast<double>(w) + cost + static_cast<double>(ac);
  }

  // faster-decoder.cc:155-241, in the reference's own processing order.
  double EmitReferenceOrder(const float *row) {
    std::vector<OrderedStateMap::Cell> old;
    map_.Take(&old);
    Cutoff cut = GetCutoff(old);
    stats_.tokens_in += static_cast<int64_t>(old.size());
    MaybeGrow(old.size());
    const float ab = cut.adaptive_beam;
    double next_cutoff = kInf;
    if (cut.best >= 0) {
      int32_t s = old[cut.best].state;
      double c = toks_[old[cut.best].tok].cost;
      for (int64_t a = g_->off[s]; a < g_->off[s + 1]; ++a) {
        if (g_->il[a] == 0) continue;
        float ac = -row[g_->il[a] - 1];
        double nw = ArcCost(g_->w[a], c, ac);
        if (nw + ab < next_cutoff) next_cutoff = nw + ab;
      }
    }
    admitted_costs_.clear();
    for (const auto &cell : old) {
      int32_t t = cell.tok;
      double c = toks_[t].cost;
      if (c < cut.weight_cutoff) {
        stats_.tokens_expanded++;
        int32_t s = cell.state;
        stats_.all_arcs_scanned += g_->off[s + 1] - g_->off[s];
        for (int64_t a = g_->off[s]; a < g_->off[s + 1]; ++a) {
          if (g_->il[a] == 0) continue;
          stats_.emit_arcs++;
          float ac = -row[g_->il[a] - 1];
          double nw = ArcCost(g_->w[a], c, ac);
          if (nw < next_cutoff) {
            stats_.admitted++;
            admitted_costs_.push_back(nw);
            int32_t nt = NewTok(static_cast<int32_t>(a), t, nw);
            int32_t cidx = map_.Insert(g_->ns[a], nt);
            if (nw + ab < next_cutoff) next_cutoff = nw + ab;
            auto &dst = map_.At(cidx);
            if (dst.tok != nt) {
              if (toks_[dst.tok].cost > nw) {
                Release(dst.tok);
                dst.tok = nt;
              } else {
                if (toks_[dst.tok].cost == nw) stats_.emit_ties++;
                Release(nt);
              }
            }
          }
        }
      }
      Release(t);
    }
    for (double nw : admitted_costs_)
      if (!(nw < next_cutoff)) stats_.extras++;
    frames_++;
    return next_cutoff;
  }

  // Order-independent statement of the same frame (what the CUDA path does).
  double EmitCanonical(const float *row) {
    std::vector<OrderedStateMap::Cell> old;
    map_.Take(&old);
    Cutoff cut = GetCutoff(old);
    stats_.tokens_in += static_cast<int64_t>(old.size());
    MaybeGrow(old.size());
    const float ab = cut.adaptive_beam;
    double min_nw = kInf;
    for (const auto &cell : old) {
      double c = toks_[cell.tok].cost;
      if (!(c < cut.weight_cutoff)) continue;
      int32_t s = cell.state;
      for (int64_t a = g_->off[s]; a < g_->off[s + 1]; ++a) {
        if (g_->il[a] == 0) continue;
        double nw = ArcCost(g_->w[a], c, -row[g_->il[a] - 1]);
        if (nw < min_nw) min_nw = nw;
      }
    }
    const double cstar = min_nw + ab;  // == final value of the running cutoff
    for (const auto &cell : old) {
      int32_t t = cell.tok;
      double c = toks_[t].cost;
      if (c < cut.weight_cutoff) {
        stats_.tokens_expanded++;
        int32_t s = cell.state;
        for (int64_t a = g_->off[s]; a < g_->off[s + 1]; ++a) {
          if (g_->il[a] == 0) continue;
          stats_.emit_arcs++;
          double nw = ArcCost(g_->w[a], c, -row[g_->il[a] - 1]);
          if (!(nw < cstar)) continue;
          stats_.admitted++;
          int32_t nt = NewTok(static_cast<int32_t>(a), t, nw);
          int32_t cidx = map_.Insert(g_->ns[a], nt);
          auto &dst = map_.At(cidx);
          if (dst.tok != nt) {
            const Tok &cur = toks_[dst.tok];
            bool better = nw < cur.cost || (nw == cur.cost && a < cur.arc);
            if (nw == cur.cost) stats_.emit_ties++;
            if (better) {
              Release(dst.tok);
              dst.tok = nt;
            } else {
              Release(nt);
            }
          }
        }
      }
      Release(t);
    }
    frames_++;
    return cstar;
  }

  // SimpleDecoder::ProcessEmitting (simple-decoder.cc:150-192), order-independent
  // statement.  Returns ProcessNonemitting's cutoff (cc:196-204).
  double EmitSimple(const float *row) {
    std::vector<OrderedStateMap::Cell> old;
    map_.Take(&old);
    stats_.tokens_in += static_cast<int64_t>(old.size());
    MaybeGrow(old.size());
    // PruneToks of the previous frame (cc:251-279); none before the first frame
    double best = kInf;
    for (const auto &cell : old) best = std::min(best, toks_[cell.tok].cost);
    const double keep_below = frames_ > 0 ? best + o_.beam : kInf;
    auto kept = [&](double c) { return frames_ > 0 ? c < keep_below : true; };
    double min_pv = kInf;
    for (const auto &cell : old) {
      double c = toks_[cell.tok].cost;
      if (!kept(c)) continue;
      int32_t s = cell.state;
      for (int64_t a = g_->off[s]; a < g_->off[s + 1]; ++a) {
        if (g_->il[a] == 0) continue;
        double pv = c + static_cast<double>(g_->w[a]) + static_cast<double>(-row[g_->il[a] - 1]);
        if (pv < min_pv) min_pv = pv;
      }
    }
    const double cstar = min_pv + o_.beam;  // final value of the running cutoff (cc:171-176)
    double min_stored = kInf;
    for (const auto &cell : old) {
      int32_t t = cell.tok;
      double c = toks_[t].cost;
      if (kept(c)) {
        stats_.tokens_expanded++;
        int32_t s = cell.state;
        for (int64_t a = g_->off[s]; a < g_->off[s + 1]; ++a) {
          if (g_->il[a] == 0) continue;
          stats_.emit_arcs++;
          const float ac = -row[g_->il[a] - 1];
          double pv = c + static_cast<double>(g_->w[a]) + static_cast<double>(ac);
          if (!(pv < cstar)) continue;  // cc:170
          stats_.admitted++;
          const float wa = g_->w[a] + ac;  // float sum, simple-decoder.h:96
          const double stored = c + static_cast<double>(wa);
          if (stored < min_stored) min_stored = stored;
          int32_t nt = NewTok(static_cast<int32_t>(a), t, stored);
          int32_t cidx = map_.Insert(g_->ns[a], nt);
          auto &dst = map_.At(cidx);
          if (dst.tok != nt) {
            const Tok &cur = toks_[dst.tok];
            bool better = stored < cur.cost || (stored == cur.cost && a < cur.arc);
            if (stored == cur.cost) stats_.emit_ties++;
            if (better) {
              Release(dst.tok);
              dst.tok = nt;
            } else {
              Release(nt);
            }
          }
        }
      }
      Release(t);
    }
    frames_++;
    return min_stored + o_.beam;
  }

  // faster-decoder.cc:59-119.  Mode 0 keeps the LIFO worklist; modes 1/2 sweep
  // to the same fixed point (costs are order-independent).  Exact ties: mode 0 keeps the
  // incumbent; modes 1/2 keep an emitting-phase incumbent, and among epsilon arrivals the
  // one over the lowest arc index (order-independent, see ExpandEps).
  void Closure(double cutoff) {
    std::vector<int32_t> work;
    map_.ListOrder(&work);
    if (mode_ != 0) {
      // deterministic: by state id; processed front to back in sweeps
      std::sort(work.begin(), work.end(), [this](int32_t a, int32_t b) {
        return map_.At(a).state < map_.At(b).state;
      });
      std::vector<int32_t> next;
      while (!work.empty()) {
        next.clear();
        for (int32_t c : work) ExpandEps(c, cutoff, &next);
        work.swap(next);
      }
      return;
    }
    while (!work.empty()) {
      int32_t c = work.back();
      work.pop_back();
      ExpandEps(c, cutoff, &work);
    }
  }

  void ExpandEps(int32_t c, double cutoff, std::vector<int32_t> *push) {
    int32_t s = map_.At(c).state;
    int32_t t = map_.At(c).tok;
    double cost = toks_[t].cost;
    // (SimpleDecoder expands every token, simple-decoder.cc:206-211)
    if (mode_ != 2 && cost > cutoff) return;
    stats_.all_arcs_scanned += g_->off[s + 1] - g_->off[s];
    for (int64_t a = g_->off[s]; a < g_->off[s + 1]; ++a) {
      if (g_->il[a] != 0) continue;
      stats_.eps_arcs++;
      double nc = cost + static_cast<double>(g_->w[a]);
      if (nc > cutoff) continue;
      int32_t nt = NewTok(static_cast<int32_t>(a), t, nc);
      int32_t cidx = map_.Insert(g_->ns[a], nt);
      auto &dst = map_.At(cidx);
      if (dst.tok == nt) {
        push->push_back(cidx);
        continue;
      }
      if (toks_[dst.tok].cost > nc) {
        Release(dst.tok);
        dst.tok = nt;
        push->push_back(cidx);
      } else {
        if (toks_[dst.tok].cost == nc) {
          stats_.eps_ties++;
          // Canonical modes: of two epsilon arrivals with bit-equal cost the one over the
          // lower arc index is the token's backpointer (the reference, mode 0, keeps the
          // first arrival of its LIFO order, faster-decoder.cc:107-112).  An incumbent
          // from the emitting phase, or the start token, stays.  The token is rewritten
          // in place: tokens already expanded from it keep pointing at it, as the CUDA
          // path's "predecessor = slot of the source state" does.
          Tok &cur = toks_[dst.tok];
          if (mode_ != 0 && cur.arc >= 0 && g_->il[cur.arc] == 0 && a < cur.arc) {
            int32_t old_prev = cur.prev;
            cur.arc = static_cast<int32_t>(a);
            cur.prev = t;
            toks_[t].refs++;
            Release(old_prev);
          }
        }
        Release(nt);
      }
      // `t` may have been freed and reused only if nobody references it; the
      // map cell `c` still does, so `cost` stays valid for this loop.
    }
  }

  const Graph *g_;
  Opts o_;
  int mode_;
  int32_t frames_ = -1;
  OrderedStateMap map_;
  std::vector<Tok> toks_;
  std::vector<int32_t> free_;
  std::vector<float> scratch_;
  std::vector<double> admitted_costs_;
  Stats stats_;
};

}  // namespace kdo

// ------------------------------------------------------------------ C ABI

namespace {
thread_local std::string g_err;

kdo::Opts MakeOpts(float beam, int32_t max_active, int32_t min_active, float beam_delta,
                   float hash_ratio) {
  kdo::Opts o;
  o.beam = beam;
  o.max_active = max_active;
  o.min_active = min_active;
  o.beam_delta = beam_delta;
  o.hash_ratio = hash_ratio;
  return o;
}

int64_t Flatten(const std::vector<kdo::PathArc> &arcs, int64_t cap, int32_t *il, int32_t *ol,
                float *gw, float *aw) {
  int64_t n = static_cast<int64_t>(arcs.size());
  for (int64_t i = 0; i < n && i < cap; ++i) {
    il[i] = arcs[i].il;
    ol[i] = arcs[i].ol;
    gw[i] = arcs[i].graph;
    aw[i] = arcs[i].ac;
  }
  return n;
}
}  // namespace

extern "C" {

const char *kdo_last_error() { return g_err.c_str(); }

void *kdo_graph_create(int32_t num_states, int32_t start, const int64_t *row_off,
                       const int32_t *ilabel, const int32_t *olabel, const float *weight,
                       const int32_t *nextstate, const float *final_w) {
  auto *g = new kdo::Graph;
  g->num_states = num_states;
  g->start = start;
  g->off.assign(row_off, row_off + num_states + 1);
  int64_t e = g->off.back();
  g->il.assign(ilabel, ilabel + e);
  g->ol.assign(olabel, olabel + e);
  g->w.assign(weight, weight + e);
  g->ns.assign(nextstate, nextstate + e);
  g->fin.assign(final_w, final_w + num_states);
  return g;
}

void kdo_graph_destroy(void *g) { delete static_cast<kdo::Graph *>(g); }

void *kdo_decoder_create(void *graph, float beam, int32_t max_active, int32_t min_active,
                         float beam_delta, float hash_ratio, int mode) {
  try {
    return new kdo::Decoder(static_cast<kdo::Graph *>(graph),
                            MakeOpts(beam, max_active, min_active, beam_delta, hash_ratio),
                            mode);
  } catch (const std::exception &e) {
    g_err = e.what();
    return nullptr;
  }
}

void kdo_decoder_destroy(void *d) { delete static_cast<kdo::Decoder *>(d); }

int kdo_decoder_set_options(void *d, float beam, int32_t max_active, int32_t min_active,
                            float beam_delta, float hash_ratio) {
  static_cast<kdo::Decoder *>(d)->SetOptions(
      MakeOpts(beam, max_active, min_active, beam_delta, hash_ratio));
  return 0;
}

int kdo_decoder_init(void *d) {
  try {
    static_cast<kdo::Decoder *>(d)->Init();
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

int kdo_decoder_advance(void *d, const float *logp, int32_t rows, int32_t cols,
                        int32_t offset, int32_t max_num_frames) {
  try {
    static_cast<kdo::Decoder *>(d)->Advance(logp, rows, cols, offset, max_num_frames);
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

int32_t kdo_decoder_num_frames_decoded(void *d) {
  return static_cast<kdo::Decoder *>(d)->NumFrames();
}

int kdo_decoder_reached_final(void *d) {
  return static_cast<kdo::Decoder *>(d)->ReachedFinal() ? 1 : 0;
}

int64_t kdo_decoder_dump_tokens(void *d, int64_t cap, int32_t *states, double *costs) {
  std::vector<int32_t> s;
  std::vector<double> c;
  static_cast<kdo::Decoder *>(d)->Tokens(&s, &c);
  int64_t n = static_cast<int64_t>(s.size());
  for (int64_t i = 0; i < n && i < cap; ++i) {
    states[i] = s[i];
    costs[i] = c[i];
  }
  return n;
}

int64_t kdo_decoder_best_path(void *d, int use_final_probs, int raw, int64_t cap, int32_t *il,
                              int32_t *ol, float *gw, float *aw, float *final2) {
  std::vector<kdo::PathArc> arcs;
  bool ok = static_cast<kdo::Decoder *>(d)->BestPath(use_final_probs != 0, raw != 0, &arcs,
                                                    final2);
  if (!ok) return -1;
  return Flatten(arcs, cap, il, ol, gw, aw);
}

// 14 int64 counters, in the order of kdo::Stats.
void kdo_decoder_stats(void *d, int64_t *out) {
  const kdo::Stats &s = static_cast<kdo::Decoder *>(d)->GetStats();
  const int64_t v[] = {s.frames,   s.tokens_in, s.tokens_expanded, s.emit_arcs,
                       s.eps_arcs, s.admitted,  s.extras,          s.emit_ties,
                       s.eps_ties, s.tokens_out, s.max_tokens,     s.binding_max,
                       s.binding_min, s.first_binding_frame, s.all_arcs_scanned};
  std::memcpy(out, v, sizeof(v));
}

// Same contract as kdref_decode_batch (oracle/ref_harness.cc), plus `mode` and
// summed counters (15 int64, may be null).
double kdo_decode_batch(void *graph, const float *logp, int32_t n_utts, int32_t max_rows,
                        const int32_t *rows, int32_t cols, float beam, int32_t max_active,
                        int32_t min_active, float beam_delta, float hash_ratio,
                        int use_final_probs, int mode, int32_t num_threads, int64_t cap,
                        int32_t *il, int32_t *ol, float *gw, float *aw, float *final2,
                        int64_t *n_out, int32_t *reached_final, int64_t *stats_sum,
                        int64_t *per_utt_stats) {
  auto *g = static_cast<kdo::Graph *>(graph);
  if (num_threads < 1) num_threads = 1;
  std::atomic<int32_t> next{0};
  std::atomic<int> failed{0};
  std::string err;
  std::vector<std::vector<int64_t>> sums(num_threads, std::vector<int64_t>(15, 0));
  auto worker = [&](int tid) {
    try {
      kdo::Decoder dec(g, MakeOpts(beam, max_active, min_active, beam_delta, hash_ratio), mode);
      while (true) {
        int32_t u = next.fetch_add(1);
        if (u >= n_utts) break;
        dec.Init();
        dec.Advance(logp + static_cast<int64_t>(u) * max_rows * cols, rows[u], cols, 0, -1);
        int rf = dec.ReachedFinal() ? 1 : 0;
        std::vector<kdo::PathArc> arcs;
        float f2[2];
        bool ok = dec.BestPath(use_final_probs != 0, false, &arcs, f2);
        if (reached_final) reached_final[u] = rf;
        if (n_out) {
          if (!ok) {
            n_out[u] = -1;
          } else {
            n_out[u] = Flatten(arcs, cap, il + u * cap, ol + u * cap, gw + u * cap, aw + u * cap);
            final2[2 * u] = f2[0];
            final2[2 * u + 1] = f2[1];
          }
        }
        int64_t st[15];
        kdo_decoder_stats(&dec, st);
        for (int i = 0; i < 15; ++i) {
          if (i == 10) sums[tid][i] = std::max(sums[tid][i], st[i]);
          else if (i == 13) sums[tid][i] += (st[i] >= 0);  // utterances with a binding frame
          else sums[tid][i] += st[i];
        }
        if (per_utt_stats) std::memcpy(per_utt_stats + 15 * u, st, sizeof(st));
      }
    } catch (const std::exception &e) {
      if (failed.exchange(1) == 0) err = e.what();
    }
  };
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> threads;
  for (int32_t t = 1; t < num_threads; ++t) threads.emplace_back(worker, t);
  worker(0);
  for (auto &t : threads) t.join();
  auto t1 = std::chrono::steady_clock::now();
  if (failed.load()) {
    g_err = err;
    return -1.0;
  }
  if (stats_sum) {
    for (int i = 0; i < 15; ++i) {
      stats_sum[i] = 0;
      for (int t = 0; t < num_threads; ++t) {
        if (i == 10) stats_sum[i] = std::max(stats_sum[i], sums[t][i]);
        else stats_sum[i] += sums[t][i];
      }
    }
  }
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
