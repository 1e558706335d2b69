// minifst/fst/fst.h
//
// A minimal, self-contained stand-in for the slice of OpenFst that the
// kaldi-decoder hot path touches.  OpenFst / kaldifst are not available in
// this build environment (the reference fetches them at configure time,
// /root/reference/cmake/kaldifst.cmake:4-24), so the B200 build ships its own
// value types with OpenFst's names and call signatures:
//
//   fst::kNoStateId, fst::kNoLabel
//   fst::TropicalWeightTpl<float>  {Value, One, Zero, ==, !=}
//   fst::ArcTpl<W>, fst::StdArc    {ilabel, olabel, weight, nextstate}
//   fst::Fst<A>                    {Start, Final, NumArcs, InitStateIterator, InitArcIterator}
//   fst::ExpandedFst<A>            {NumStates}   (as in OpenFst, NOT on Fst<A>)
//   fst::CountStates(const Fst<A>&)
//   fst::MutableFst<A>             {DeleteStates, AddState, SetStart, AddArc,
//                                   SetFinal}
//   fst::VectorFst<A>, fst::ConstFst<A>
//   fst::ArcIterator<F>, fst::StateIterator<F>
//
// Users that do have OpenFst should compile against the real headers instead
// (put them before this directory on the include path); the decoder only uses
// the names above (reference use sites: faster-decoder.cc:45-51, 80-82,
// 179-181, 205-207, 350, 364, 408-420).
#ifndef KALDI_DECODER_B200_MINIFST_FST_FST_H_
#define KALDI_DECODER_B200_MINIFST_FST_FST_H_

#include <sys/types.h>

#include <atomic>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace fst {

constexpr int kNoStateId = -1;
constexpr int kNoLabel = -1;

// Tropical semiring over T: Plus = min, Times = +, Zero = +inf, One = 0.
template <class T>
class TropicalWeightTpl {
 public:
  using ValueType = T;

  constexpr TropicalWeightTpl() : value_(0) {}
  constexpr TropicalWeightTpl(T v) : value_(v) {}  // NOLINT (implicit, as OpenFst)

  static constexpr TropicalWeightTpl Zero() {
    return TropicalWeightTpl(std::numeric_limits<T>::infinity());
  }
  static constexpr TropicalWeightTpl One() { return TropicalWeightTpl(0); }

  constexpr const T &Value() const { return value_; }

  static const std::string &Type() {
    static const std::string type = "tropical";
    return type;
  }

 private:
  T value_;
};

template <class T>
constexpr bool operator==(const TropicalWeightTpl<T> &a,
                          const TropicalWeightTpl<T> &b) {
  return a.Value() == b.Value();
}

template <class T>
constexpr bool operator!=(const TropicalWeightTpl<T> &a,
                          const TropicalWeightTpl<T> &b) {
  return !(a == b);
}

template <class T>
constexpr TropicalWeightTpl<T> Times(const TropicalWeightTpl<T> &a,
                                     const TropicalWeightTpl<T> &b) {
  return TropicalWeightTpl<T>(a.Value() + b.Value());
}

template <class T>
constexpr TropicalWeightTpl<T> Plus(const TropicalWeightTpl<T> &a,
                                    const TropicalWeightTpl<T> &b) {
  return a.Value() < b.Value() ? a : b;
}

using TropicalWeight = TropicalWeightTpl<float>;

template <class W>
struct ArcTpl {
  using Weight = W;
  using Label = int;
  using StateId = int;

  Label ilabel;
  Label olabel;
  Weight weight;
  StateId nextstate;

  ArcTpl() : ilabel(0), olabel(0), weight(), nextstate(kNoStateId) {}
  ArcTpl(Label il, Label ol, Weight w, StateId ns)
      : ilabel(il), olabel(ol), weight(std::move(w)), nextstate(ns) {}

  static const std::string &Type() {
    static const std::string type =
        W::Type() == "tropical" ? std::string("standard") : W::Type();
    return type;
  }
};

using StdArc = ArcTpl<TropicalWeight>;

// What an ArcIterator needs to walk the arcs leaving one state.
template <class A>
struct ArcIteratorData {
  const A *arcs = nullptr;
  size_t narcs = 0;
};

// What a StateIterator needs (every minifst type numbers its states 0..nstates-1).
template <class A>
struct StateIteratorData {
  typename A::StateId nstates = 0;
};

// Identifies the CONTENT an FST object holds, for caches keyed on "exactly this graph" (the
// device copy of a decoding graph, faster-decoder.cc): a process-unique number the object gets
// the first time it is asked, keeps while it is only read, hands to its copies, and drops when it
// is modified.  (minifst only: OpenFst has no such thing, and code that must also build against
// OpenFst detects ContentId() before using it.)
class ContentTag {
 public:
  ContentTag() = default;
  ContentTag(const ContentTag &o) : id_(o.id_.load(std::memory_order_relaxed)) {}
  ContentTag(ContentTag &&o) noexcept : id_(o.id_.exchange(0, std::memory_order_relaxed)) {}
  ContentTag &operator=(const ContentTag &o) {
    id_.store(o.id_.load(std::memory_order_relaxed), std::memory_order_relaxed);
    return *this;
  }
  ContentTag &operator=(ContentTag &&o) noexcept {
    if (this != &o) id_.store(o.id_.exchange(0, std::memory_order_relaxed), std::memory_order_relaxed);
    return *this;
  }
  uint64_t Get() const {
    uint64_t v = id_.load(std::memory_order_relaxed);
    if (v == 0) {
      static std::atomic<uint64_t> next{1};
      const uint64_t fresh = next.fetch_add(1, std::memory_order_relaxed);
      v = id_.compare_exchange_strong(v, fresh, std::memory_order_relaxed) ? fresh : v;
    }
    return v;
  }
  void Drop() { id_.store(0, std::memory_order_relaxed); }

 private:
  mutable std::atomic<uint64_t> id_{0};
};

// Abstract read-only FST.
template <class A>
class Fst {
 public:
  using Arc = A;
  using StateId = typename A::StateId;
  using Weight = typename A::Weight;

  virtual ~Fst() = default;

  // see ContentTag; two objects with the same ContentId() hold the same states and arcs
  uint64_t ContentId() const { return content_.Get(); }

 protected:
  // every mutator of a derived class calls this (MutableArcs() too: its caller is about to write)
  void ContentChanged() { content_.Drop(); }
  // for converting constructors: this object now holds exactly what `other` holds
  void SameContentAs(const Fst &other) { content_ = other.content_; }

 private:
  ContentTag content_;

 public:

  virtual StateId Start() const = 0;
  virtual Weight Final(StateId s) const = 0;
  virtual size_t NumArcs(StateId s) const = 0;
  virtual void InitStateIterator(StateIteratorData<A> *data) const = 0;
  virtual void InitArcIterator(StateId s, ArcIteratorData<A> *data) const = 0;
  virtual const std::string &Type() const = 0;
};

// An FST whose states can be counted without walking them.  As in OpenFst, NumStates() is
// NOT part of Fst<A>: code that only has a `const Fst<A>&` uses StateIterator / CountStates.
template <class A>
class ExpandedFst : public Fst<A> {
 public:
  using StateId = typename A::StateId;
  virtual StateId NumStates() const = 0;
  void InitStateIterator(StateIteratorData<A> *data) const override {
    data->nstates = NumStates();
  }
};

template <class A>
class MutableFst : public ExpandedFst<A> {
 public:
  using StateId = typename A::StateId;
  using Weight = typename A::Weight;

  virtual void DeleteStates() = 0;
  virtual StateId AddState() = 0;
  virtual void SetStart(StateId s) = 0;
  virtual void AddArc(StateId s, const A &arc) = 0;
  virtual void SetFinal(StateId s, Weight w) = 0;
  virtual void ReserveArcs(StateId s, size_t n) = 0;
  virtual void DeleteArcs(StateId s) = 0;
  // Direct arc mutation (OpenFst exposes this through MutableArcIterator).
  virtual A *MutableArcs(StateId s) = 0;
};

// Adjacency-list FST: one vector of arcs per state.
template <class A>
class VectorFst : public MutableFst<A> {
 public:
  using Arc = A;
  using StateId = typename A::StateId;
  using Weight = typename A::Weight;

  VectorFst() = default;

  explicit VectorFst(const Fst<A> &other) { CopyFrom(other); }
  VectorFst(const VectorFst &) = default;
  VectorFst(VectorFst &&) = default;
  VectorFst &operator=(const VectorFst &) = default;
  VectorFst &operator=(VectorFst &&) = default;

  StateId Start() const override { return start_; }
  Weight Final(StateId s) const override { return states_.at(s).final; }
  size_t NumArcs(StateId s) const override { return states_.at(s).arcs.size(); }
  StateId NumStates() const override {
    return static_cast<StateId>(states_.size());
  }
  void InitArcIterator(StateId s, ArcIteratorData<A> *data) const override {
    const auto &arcs = states_.at(s).arcs;
    data->arcs = arcs.data();
    data->narcs = arcs.size();
  }
  const std::string &Type() const override {
    static const std::string type = "vector";
    return type;
  }

  void DeleteStates() override {
    this->ContentChanged();
    states_.clear();
    start_ = kNoStateId;
  }
  StateId AddState() override {
    this->ContentChanged();
    states_.emplace_back();
    return static_cast<StateId>(states_.size() - 1);
  }
  void SetStart(StateId s) override {
    this->ContentChanged();
    start_ = s;
  }
  void AddArc(StateId s, const A &arc) override {
    this->ContentChanged();
    states_.at(s).arcs.push_back(arc);
  }
  void SetFinal(StateId s, Weight w) override {
    this->ContentChanged();
    states_.at(s).final = w;
  }
  void ReserveArcs(StateId s, size_t n) override {
    states_.at(s).arcs.reserve(n);
  }
  void DeleteArcs(StateId s) override {
    this->ContentChanged();
    states_.at(s).arcs.clear();
  }
  A *MutableArcs(StateId s) override {
    this->ContentChanged();
    return states_.at(s).arcs.data();
  }

  void ReserveStates(size_t n) { states_.reserve(n); }

  // Keeps only the states with keep[s] == true, renumbering the survivors in
  // increasing order of their old ids and dropping arcs into removed states.
  void KeepStates(const std::vector<bool> &keep) {
    this->ContentChanged();
    std::vector<StateId> renumber(states_.size(), kNoStateId);
    StateId n = 0;
    for (size_t s = 0; s < states_.size(); ++s) {
      if (keep[s]) renumber[s] = n++;
    }
    std::vector<State> out;
    out.reserve(n);
    for (size_t s = 0; s < states_.size(); ++s) {
      if (!keep[s]) continue;
      State st;
      st.final = states_[s].final;
      for (const A &a : states_[s].arcs) {
        if (a.nextstate >= 0 && renumber[a.nextstate] != kNoStateId) {
          A b = a;
          b.nextstate = renumber[a.nextstate];
          st.arcs.push_back(b);
        }
      }
      out.push_back(std::move(st));
    }
    start_ = (start_ >= 0 && renumber[start_] != kNoStateId) ? renumber[start_]
                                                             : kNoStateId;
    states_ = std::move(out);
  }

 private:
  struct State {
    Weight final = Weight::Zero();
    std::vector<A> arcs;
  };

  void CopyFrom(const Fst<A> &other) {
    StateIteratorData<A> sd;
    other.InitStateIterator(&sd);
    StateId n = sd.nstates;
    states_.resize(n);
    start_ = other.Start();
    for (StateId s = 0; s < n; ++s) {
      states_[s].final = other.Final(s);
      ArcIteratorData<A> d;
      other.InitArcIterator(s, &d);
      states_[s].arcs.assign(d.arcs, d.arcs + d.narcs);
    }
    this->SameContentAs(other);
  }

  StateId start_ = kNoStateId;
  std::vector<State> states_;
};

// Immutable CSR FST: one contiguous arc array + per-state offsets.  This is
// the layout OpenFst's ConstFst has on disk and in memory.
template <class A>
class ConstFst : public ExpandedFst<A> {
 public:
  using Arc = A;
  using StateId = typename A::StateId;
  using Weight = typename A::Weight;

  ConstFst() : offsets_(1, 0) {}

  explicit ConstFst(const Fst<A> &other) {
    StateIteratorData<A> sd;
    other.InitStateIterator(&sd);
    StateId n = sd.nstates;
    start_ = other.Start();
    finals_.resize(n);
    offsets_.assign(1, 0);
    offsets_.reserve(n + 1);
    for (StateId s = 0; s < n; ++s) {
      finals_[s] = other.Final(s);
      ArcIteratorData<A> d;
      other.InitArcIterator(s, &d);
      arcs_.insert(arcs_.end(), d.arcs, d.arcs + d.narcs);
      offsets_.push_back(arcs_.size());
    }
    this->SameContentAs(other);
  }

  // Takes ownership of pre-built CSR arrays.  offsets.size() == finals.size()+1
  ConstFst(StateId start, std::vector<size_t> offsets, std::vector<A> arcs,
           std::vector<Weight> finals)
      : start_(start),
        offsets_(std::move(offsets)),
        arcs_(std::move(arcs)),
        finals_(std::move(finals)) {
    if (offsets_.size() != finals_.size() + 1 || offsets_.back() != arcs_.size())
      throw std::runtime_error("ConstFst: inconsistent CSR arrays");
  }

  StateId Start() const override { return start_; }
  Weight Final(StateId s) const override { return finals_.at(s); }
  size_t NumArcs(StateId s) const override {
    return offsets_.at(s + 1) - offsets_.at(s);
  }
  StateId NumStates() const override {
    return static_cast<StateId>(finals_.size());
  }
  void InitArcIterator(StateId s, ArcIteratorData<A> *data) const override {
    data->arcs = arcs_.data() + offsets_[s];
    data->narcs = offsets_[s + 1] - offsets_[s];
  }
  const std::string &Type() const override {
    static const std::string type = "const";
    return type;
  }

  const std::vector<size_t> &Offsets() const { return offsets_; }
  const std::vector<A> &Arcs() const { return arcs_; }

 private:
  StateId start_ = kNoStateId;
  std::vector<size_t> offsets_;
  std::vector<A> arcs_;
  std::vector<Weight> finals_;
};

// ArcIterator<F> for any F derived from Fst<A>: one virtual call per state,
// then plain pointer walking (the same cost model as OpenFst).
template <class F>
class ArcIterator {
 public:
  using Arc = typename F::Arc;
  using StateId = typename Arc::StateId;

  ArcIterator(const F &fst, StateId s) { fst.InitArcIterator(s, &data_); }

  bool Done() const { return i_ >= data_.narcs; }
  const Arc &Value() const { return data_.arcs[i_]; }
  void Next() { ++i_; }
  void Reset() { i_ = 0; }
  void Seek(size_t a) { i_ = a; }
  size_t Position() const { return i_; }

 private:
  ArcIteratorData<Arc> data_;
  size_t i_ = 0;
};

template <class F>
class StateIterator {
 public:
  using StateId = typename F::Arc::StateId;

  explicit StateIterator(const F &fst) {
    StateIteratorData<typename F::Arc> d;
    fst.InitStateIterator(&d);
    n_ = d.nstates;
  }

  bool Done() const { return s_ >= n_; }
  StateId Value() const { return s_; }
  void Next() { ++s_; }
  void Reset() { s_ = 0; }

 private:
  StateId n_ = 0;
  StateId s_ = 0;
};

// Number of states of any FST (OpenFst: expanded-fst.h).
template <class A>
typename A::StateId CountStates(const Fst<A> &fst) {
  typename A::StateId n = 0;
  for (StateIterator<Fst<A>> it(fst); !it.Done(); it.Next()) ++n;
  return n;
}

using StdFst = Fst<StdArc>;
using StdVectorFst = VectorFst<StdArc>;
using StdConstFst = ConstFst<StdArc>;

}  // namespace fst

#endif  // KALDI_DECODER_B200_MINIFST_FST_FST_H_
