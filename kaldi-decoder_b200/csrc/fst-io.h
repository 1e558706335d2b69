// kaldi-decoder_b200/csrc/fst-io.h
//
// Graph input and result output without kaldifst/OpenFst (neither is available
// in this build): the reference's scripts do
//     HLG = kaldifst.StdVectorFst.read(path)          # OpenFst binary
//     ok, best = decoder.get_best_path()
//     ok, isyms, osyms, w = kaldifst.get_linear_symbol_sequence(best)
// (SURVEY.md App. B.2, B.4).  This file provides those steps for the minifst
// types: OpenFst binary ("vector" and "const", standard arcs) and AT&T text
// readers/writers, and GetLinearSymbolSequence.  The binary layout follows
// OpenFst's published format; it is round-trip tested here
// but UNVERIFIED against files written by a real OpenFst (none available).
#ifndef KALDI_DECODER_B200_CSRC_FST_IO_H_
#define KALDI_DECODER_B200_CSRC_FST_IO_H_

#include <cstdint>
#include <iosfwd>
#include <string>
#include <vector>

#include "fst/fst.h"
#include "kaldifst/csrc/lattice-weight.h"

namespace kaldi_decoder {

fst::StdVectorFst ReadFstBinary(std::istream &is);
fst::StdVectorFst ReadFst(const std::string &path);
void WriteFstBinary(const fst::Fst<fst::StdArc> &fst, std::ostream &os);
void WriteFst(const fst::Fst<fst::StdArc> &fst, const std::string &path);

// AT&T text: "src dst ilabel olabel [weight]" (acceptor: "src dst label [weight]"),
// final states "state [weight]"; the source of the first line is the start state.
fst::StdVectorFst ReadFstText(const std::string &text, bool acceptor = false);
std::string WriteFstText(const fst::Fst<fst::StdArc> &fst);

// Follows the single path of a linear FST from its start state: non-epsilon
// input / output labels in order and the product of the weights incl. the final
// weight.  Returns false if the FST is not linear (or is empty).
bool GetLinearSymbolSequence(const fst::Fst<fst::LatticeArc> &fst, std::vector<int32_t> *isyms,
                             std::vector<int32_t> *osyms, fst::LatticeWeight *total);

}  // namespace kaldi_decoder

#endif  // KALDI_DECODER_B200_CSRC_FST_IO_H_
