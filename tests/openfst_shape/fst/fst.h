// tests/openfst_shape/fst/fst.h  -- TEST INFRASTRUCTURE, compile-only.
//
// The class layout of real OpenFst (1.7 / 1.8: fst/fst.h, fst/expanded-fst.h,
// fst/mutable-fst.h, fst/vector-fst.h), declarations only, written from its public
// interface: the abstract Fst<Arc> has NO NumStates() (that is ExpandedFst's), iterators
// over an abstract FST go through InitStateIterator / InitArcIterator, whose data carry
// an optional polymorphic `base`.  The product's host sources are syntax-checked against
// THIS header instead of csrc/minifst (tests/test_host_logic.py), so a use of anything
// only minifst offers on Fst<Arc> -- as round 1's Fst::NumStates() -- fails the test.
#ifndef TESTS_OPENFST_SHAPE_FST_FST_H_
#define TESTS_OPENFST_SHAPE_FST_FST_H_

#include <cstddef>
#include <cstdint>
#include <limits>
#include <memory>
#include <string>
#include <vector>

namespace fst {

constexpr int kNoStateId = -1;
constexpr int kNoLabel = -1;

template <class T>
class TropicalWeightTpl {
 public:
  using ValueType = T;
  constexpr TropicalWeightTpl() : value_(0) {}
  constexpr TropicalWeightTpl(T v) : value_(v) {}  // NOLINT
  static constexpr TropicalWeightTpl Zero() {
    return TropicalWeightTpl(std::numeric_limits<T>::infinity());
  }
  static constexpr TropicalWeightTpl One() { return TropicalWeightTpl(0); }
  constexpr const T &Value() const { return value_; }
  static const std::string &Type();

 private:
  T value_;
};
template <class T>
bool operator==(const TropicalWeightTpl<T> &a, const TropicalWeightTpl<T> &b) {
  return a.Value() == b.Value();
}
template <class T>
bool operator!=(const TropicalWeightTpl<T> &a, const TropicalWeightTpl<T> &b) {
  return !(a == b);
}
template <class T>
TropicalWeightTpl<T> Times(const TropicalWeightTpl<T> &a, const TropicalWeightTpl<T> &b) {
  return TropicalWeightTpl<T>(a.Value() + b.Value());
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
  ArcTpl() noexcept(std::is_nothrow_default_constructible<Weight>::value) {}
  template <class T>
  ArcTpl(Label il, Label ol, T &&w, StateId ns)
      : ilabel(il), olabel(ol), weight(std::forward<T>(w)), nextstate(ns) {}
  static const std::string &Type();
};
using StdArc = ArcTpl<TropicalWeight>;

template <class Arc>
class StateIteratorBase {
 public:
  using StateId = typename Arc::StateId;
  virtual ~StateIteratorBase() {}
  virtual bool Done() const = 0;
  virtual StateId Value() const = 0;
  virtual void Next() = 0;
  virtual void Reset() = 0;
};
template <class Arc>
struct StateIteratorData {
  std::unique_ptr<StateIteratorBase<Arc>> base;
  typename Arc::StateId nstates;
};

template <class Arc>
class ArcIteratorBase {
 public:
  virtual ~ArcIteratorBase() {}
  virtual bool Done() const = 0;
  virtual const Arc &Value() const = 0;
  virtual void Next() = 0;
  virtual size_t Position() const = 0;
  virtual void Reset() = 0;
  virtual void Seek(size_t) = 0;
};
template <class Arc>
struct ArcIteratorData {
  std::unique_ptr<ArcIteratorBase<Arc>> base;
  const Arc *arcs = nullptr;
  size_t narcs = 0;
  int *ref_count = nullptr;
};

// fst/fst.h: the abstract FST.  No NumStates().
template <class A>
class Fst {
 public:
  using Arc = A;
  using StateId = typename Arc::StateId;
  using Weight = typename Arc::Weight;
  virtual ~Fst() {}
  virtual StateId Start() const = 0;
  virtual Weight Final(StateId) const = 0;
  virtual size_t NumArcs(StateId) const = 0;
  virtual size_t NumInputEpsilons(StateId) const = 0;
  virtual size_t NumOutputEpsilons(StateId) const = 0;
  virtual uint64_t Properties(uint64_t mask, bool test) const = 0;
  virtual const std::string &Type() const = 0;
  virtual Fst *Copy(bool safe = false) const = 0;
  virtual void InitStateIterator(StateIteratorData<Arc> *data) const = 0;
  virtual void InitArcIterator(StateId s, ArcIteratorData<Arc> *data) const = 0;
};

// fst/expanded-fst.h
template <class A>
class ExpandedFst : public Fst<A> {
 public:
  using StateId = typename A::StateId;
  virtual StateId NumStates() const = 0;
};

// fst/mutable-fst.h
template <class A>
class MutableFst : public ExpandedFst<A> {
 public:
  using Arc = A;
  using StateId = typename Arc::StateId;
  using Weight = typename Arc::Weight;
  virtual void SetStart(StateId s) = 0;
  virtual void SetFinal(StateId s, Weight weight = Weight::One()) = 0;
  virtual StateId AddState() = 0;
  virtual void AddArc(StateId s, const Arc &arc) = 0;
  virtual void DeleteStates(const std::vector<StateId> &dstates) = 0;
  virtual void DeleteStates() = 0;
  virtual void DeleteArcs(StateId s, size_t n) = 0;
  virtual void DeleteArcs(StateId s) = 0;
  virtual void ReserveStates(size_t) {}
  virtual void ReserveArcs(StateId, size_t) {}
};

// fst/vector-fst.h (the implementation classes are folded away)
template <class A>
class VectorFst : public MutableFst<A> {
 public:
  using Arc = A;
  using StateId = typename Arc::StateId;
  using Weight = typename Arc::Weight;
  VectorFst();
  explicit VectorFst(const Fst<Arc> &fst);
  VectorFst(const VectorFst &fst, bool safe = false);
  VectorFst &operator=(const VectorFst &fst);
  StateId Start() const override;
  Weight Final(StateId) const override;
  size_t NumArcs(StateId) const override;
  size_t NumInputEpsilons(StateId) const override;
  size_t NumOutputEpsilons(StateId) const override;
  uint64_t Properties(uint64_t mask, bool test) const override;
  const std::string &Type() const override;
  VectorFst *Copy(bool safe = false) const override;
  StateId NumStates() const override;
  void InitStateIterator(StateIteratorData<Arc> *data) const override;
  void InitArcIterator(StateId s, ArcIteratorData<Arc> *data) const override;
  void SetStart(StateId s) override;
  void SetFinal(StateId s, Weight weight = Weight::One()) override;
  StateId AddState() override;
  void AddArc(StateId s, const Arc &arc) override;
  void DeleteStates(const std::vector<StateId> &dstates) override;
  void DeleteStates() override;
  void DeleteArcs(StateId s, size_t n) override;
  void DeleteArcs(StateId s) override;
  void ReserveStates(size_t n) override;
  void ReserveArcs(StateId s, size_t n) override;
};

// fst/const-fst.h
template <class A, class Unsigned = uint32_t>
class ConstFst : public ExpandedFst<A> {
 public:
  using Arc = A;
  using StateId = typename Arc::StateId;
  using Weight = typename Arc::Weight;
  ConstFst();
  explicit ConstFst(const Fst<Arc> &fst);
  StateId Start() const override;
  Weight Final(StateId) const override;
  size_t NumArcs(StateId) const override;
  size_t NumInputEpsilons(StateId) const override;
  size_t NumOutputEpsilons(StateId) const override;
  uint64_t Properties(uint64_t mask, bool test) const override;
  const std::string &Type() const override;
  ConstFst *Copy(bool safe = false) const override;
  StateId NumStates() const override;
  void InitStateIterator(StateIteratorData<Arc> *data) const override;
  void InitArcIterator(StateId s, ArcIteratorData<Arc> *data) const override;
};

// Iterators: class templates specialised per FST type in OpenFst; the generic form goes
// through Init*Iterator and handles a polymorphic base.
template <class FST>
class StateIterator {
 public:
  using Arc = typename FST::Arc;
  using StateId = typename Arc::StateId;
  explicit StateIterator(const FST &fst) : s_(0) { fst.InitStateIterator(&data_); }
  bool Done() const { return data_.base ? data_.base->Done() : s_ >= data_.nstates; }
  StateId Value() const { return data_.base ? data_.base->Value() : s_; }
  void Next() {
    if (data_.base) {
      data_.base->Next();
    } else {
      ++s_;
    }
  }

 private:
  StateIteratorData<Arc> data_;
  StateId s_;
};

template <class FST>
class ArcIterator {
 public:
  using Arc = typename FST::Arc;
  using StateId = typename Arc::StateId;
  ArcIterator(const FST &fst, StateId s) : i_(0) { fst.InitArcIterator(s, &data_); }
  bool Done() const { return data_.base ? data_.base->Done() : i_ >= data_.narcs; }
  const Arc &Value() const { return data_.base ? data_.base->Value() : data_.arcs[i_]; }
  void Next() {
    if (data_.base) {
      data_.base->Next();
    } else {
      ++i_;
    }
  }

 private:
  ArcIteratorData<Arc> data_;
  size_t i_;
};

// fst/expanded-fst.h
template <class Arc>
typename Arc::StateId CountStates(const Fst<Arc> &fst);

using StdFst = Fst<StdArc>;
using StdVectorFst = VectorFst<StdArc>;
using StdConstFst = ConstFst<StdArc>;

}  // namespace fst

#endif  // TESTS_OPENFST_SHAPE_FST_FST_H_
