// tests/openfst_shape/kaldifst/csrc/remove-eps-local.h -- compile-only: kaldifst's declaration.
#ifndef TESTS_OPENFST_SHAPE_KALDIFST_REMOVE_EPS_LOCAL_H_
#define TESTS_OPENFST_SHAPE_KALDIFST_REMOVE_EPS_LOCAL_H_
#include "fst/fst.h"
namespace fst {
template <class Arc>
void RemoveEpsLocal(MutableFst<Arc> *fst);
}  // namespace fst
#endif
