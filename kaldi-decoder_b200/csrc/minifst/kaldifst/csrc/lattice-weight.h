// minifst/kaldifst/csrc/lattice-weight.h
//
// Stand-in for kaldifst's LatticeWeight (kaldifst v1.8.0 is pinned by the
// reference at cmake/kaldifst.cmake:4 but is not vendored).  A lattice weight
// is a pair (graph cost, acoustic cost); the semiring compares by the sum and
// breaks ties on the first component.  Reference use sites:
// faster-decoder.cc:398-420 (LatticeWeight(graph, ac), LatticeWeight::One()).
#ifndef KALDI_DECODER_B200_MINIFST_KALDIFST_CSRC_LATTICE_WEIGHT_H_
#define KALDI_DECODER_B200_MINIFST_KALDIFST_CSRC_LATTICE_WEIGHT_H_

#include <limits>
#include <string>

#include "fst/fst.h"

namespace fst {

template <class T>
class LatticeWeightTpl {
 public:
  using ValueType = T;

  LatticeWeightTpl() : v1_(0), v2_(0) {}
  LatticeWeightTpl(T graph, T acoustic) : v1_(graph), v2_(acoustic) {}

  static LatticeWeightTpl Zero() {
    return LatticeWeightTpl(std::numeric_limits<T>::infinity(),
                            std::numeric_limits<T>::infinity());
  }
  static LatticeWeightTpl One() { return LatticeWeightTpl(0, 0); }

  T Value1() const { return v1_; }
  T Value2() const { return v2_; }
  void SetValue1(T v) { v1_ = v; }
  void SetValue2(T v) { v2_ = v; }

  static const std::string &Type() {
    static const std::string type = "lattice4";
    return type;
  }

 private:
  T v1_;
  T v2_;
};

template <class T>
inline bool operator==(const LatticeWeightTpl<T> &a,
                       const LatticeWeightTpl<T> &b) {
  return a.Value1() == b.Value1() && a.Value2() == b.Value2();
}

template <class T>
inline bool operator!=(const LatticeWeightTpl<T> &a,
                       const LatticeWeightTpl<T> &b) {
  return !(a == b);
}

// -1 if a is worse (larger total) than b, +1 if better, 0 if equal.
template <class T>
inline int Compare(const LatticeWeightTpl<T> &a, const LatticeWeightTpl<T> &b) {
  T fa = a.Value1() + a.Value2(), fb = b.Value1() + b.Value2();
  if (fa < fb) return 1;
  if (fa > fb) return -1;
  if (a.Value1() < b.Value1()) return 1;
  if (a.Value1() > b.Value1()) return -1;
  return 0;
}

template <class T>
inline LatticeWeightTpl<T> Plus(const LatticeWeightTpl<T> &a,
                                const LatticeWeightTpl<T> &b) {
  return Compare(a, b) >= 0 ? a : b;
}

template <class T>
inline LatticeWeightTpl<T> Times(const LatticeWeightTpl<T> &a,
                                 const LatticeWeightTpl<T> &b) {
  return LatticeWeightTpl<T>(a.Value1() + b.Value1(), a.Value2() + b.Value2());
}

using LatticeWeight = LatticeWeightTpl<float>;
using LatticeArc = ArcTpl<LatticeWeight>;
using Lattice = VectorFst<LatticeArc>;

}  // namespace fst

#endif  // KALDI_DECODER_B200_MINIFST_KALDIFST_CSRC_LATTICE_WEIGHT_H_
