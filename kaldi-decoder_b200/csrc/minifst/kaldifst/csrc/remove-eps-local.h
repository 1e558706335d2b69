// minifst/kaldifst/csrc/remove-eps-local.h
//
// Stand-in for kaldifst's fst::RemoveEpsLocal (kaldifst v1.8.0, pinned by the
// reference at cmake/kaldifst.cmake:4, not vendored; call site
// faster-decoder.cc:422).  PARITY UNPINNED: no reference test covers it; this
// restates the published Kaldi algorithm (fstext/remove-eps-local-inl.h) for
// the only input shape the decoder ever produces, a chain-like acyclic FST:
//
//   visit states in id order, and every arc position of a state in order
//   (arcs appended while visiting are visited too); for an arc s --a--> n with
//   n != s where n has exactly one way out (one arc, or being final):
//     * n final, a is (eps, eps): fold  final(s) = Plus(final(s), a.w * final(n))
//     * n has one arc b and a, b do not both carry an ilabel nor both an
//       olabel: add the combined arc  s --(a|b, a.w * b.w)--> b.next
//     in both cases a is deleted, and so is n's exit when a was n's only entry;
//   finally drop states that became unreachable/dead and renumber in id order.
//
// "Pattern 1" of the original (next state with one entry and several exits,
// which needs weight pushing) cannot fire on a linear chain and is not
// implemented: such a state is left untouched.
#ifndef KALDI_DECODER_B200_MINIFST_KALDIFST_CSRC_REMOVE_EPS_LOCAL_H_
#define KALDI_DECODER_B200_MINIFST_KALDIFST_CSRC_REMOVE_EPS_LOCAL_H_

#include <vector>

#include "fst/fst.h"

namespace fst {

namespace minifst_internal {

template <class Arc>
inline bool CombineArcs(const Arc &a, const Arc &b, Arc *c) {
  if (a.ilabel != 0 && b.ilabel != 0) return false;
  if (a.olabel != 0 && b.olabel != 0) return false;
  c->ilabel = a.ilabel != 0 ? a.ilabel : b.ilabel;
  c->olabel = a.olabel != 0 ? a.olabel : b.olabel;
  c->weight = Times(a.weight, b.weight);
  c->nextstate = b.nextstate;
  return true;
}

}  // namespace minifst_internal

template <class Arc>
void RemoveEpsLocal(MutableFst<Arc> *fst) {
  using StateId = typename Arc::StateId;
  using Weight = typename Arc::Weight;
  const StateId kDead = kNoStateId;  // marks a deleted arc

  const StateId num_states = fst->NumStates();
  if (fst->Start() == kNoStateId || num_states == 0) return;

  // ways in (arcs + being the start state) and out (arcs + being final).
  std::vector<int> n_in(num_states, 0), n_out(num_states, 0);
  n_in[fst->Start()]++;
  for (StateId s = 0; s < num_states; ++s) {
    ArcIteratorData<Arc> d;
    fst->InitArcIterator(s, &d);
    for (size_t i = 0; i < d.narcs; ++i) {
      n_in[d.arcs[i].nextstate]++;
      n_out[s]++;
    }
    if (fst->Final(s) != Weight::Zero()) n_out[s]++;
  }

  for (StateId s = 0; s < num_states; ++s) {
    for (size_t pos = 0; pos < fst->NumArcs(s); ++pos) {
      Arc arc = fst->MutableArcs(s)[pos];
      const StateId n = arc.nextstate;
      if (n == kDead || n == s) continue;
      if (n_out[n] != 1) continue;  // pattern 2 only (see header comment)

      const bool sole_entry = (n_in[n] == 1);
      bool remove_arc = false;
      Weight n_final = fst->Final(n);
      if (n_final != Weight::Zero()) {
        // n's single way out is its final weight.
        if (arc.ilabel == 0 && arc.olabel == 0) {
          Weight folded = Times(arc.weight, n_final);
          if (fst->Final(s) == Weight::Zero()) n_out[s]++;
          fst->SetFinal(s, Plus(fst->Final(s), folded));
          remove_arc = true;
          if (sole_entry) {
            n_out[n]--;
            fst->SetFinal(n, Weight::Zero());
          }
        }
      } else {
        // n's single way out is one live arc.
        Arc *n_arcs = fst->MutableArcs(n);
        size_t j = 0;
        while (n_arcs[j].nextstate == kDead) ++j;
        Arc next_arc = n_arcs[j];
        Arc combined;
        if (minifst_internal::CombineArcs(arc, next_arc, &combined)) {
          remove_arc = true;
          if (sole_entry) {
            n_out[n]--;
            n_in[next_arc.nextstate]--;
            n_arcs[j].nextstate = kDead;
          }
          fst->AddArc(s, combined);  // may reallocate s's arcs
          n_out[s]++;
          n_in[combined.nextstate]++;
        }
      }
      if (remove_arc) {
        n_out[s]--;
        n_in[n]--;
        fst->MutableArcs(s)[pos].nextstate = kDead;
      }
    }
  }

  // Connect(): keep states that are reachable from the start through live
  // arcs and can reach a final state; renumber in increasing id order.
  std::vector<bool> reach(num_states, false), coreach(num_states, false);
  std::vector<std::vector<StateId>> rev(num_states);
  {
    std::vector<StateId> stack{fst->Start()};
    reach[fst->Start()] = true;
    while (!stack.empty()) {
      StateId s = stack.back();
      stack.pop_back();
      ArcIteratorData<Arc> d;
      fst->InitArcIterator(s, &d);
      for (size_t i = 0; i < d.narcs; ++i) {
        StateId n = d.arcs[i].nextstate;
        if (n == kDead) continue;
        rev[n].push_back(s);
        if (!reach[n]) {
          reach[n] = true;
          stack.push_back(n);
        }
      }
    }
    for (StateId s = 0; s < num_states; ++s) {
      if (reach[s] && fst->Final(s) != Weight::Zero() && !coreach[s]) {
        coreach[s] = true;
        stack.push_back(s);
      }
    }
    while (!stack.empty()) {
      StateId s = stack.back();
      stack.pop_back();
      for (StateId p : rev[s]) {
        if (!coreach[p]) {
          coreach[p] = true;
          stack.push_back(p);
        }
      }
    }
  }

  // Rebuild in place through the MutableFst interface.
  std::vector<StateId> renumber(num_states, kNoStateId);
  StateId kept = 0;
  for (StateId s = 0; s < num_states; ++s) {
    if (reach[s] && coreach[s]) renumber[s] = kept++;
  }
  struct Saved {
    Weight final;
    std::vector<Arc> arcs;
  };
  std::vector<Saved> saved;
  saved.reserve(kept);
  for (StateId s = 0; s < num_states; ++s) {
    if (renumber[s] == kNoStateId) continue;
    Saved sv;
    sv.final = fst->Final(s);
    ArcIteratorData<Arc> d;
    fst->InitArcIterator(s, &d);
    for (size_t i = 0; i < d.narcs; ++i) {
      StateId n = d.arcs[i].nextstate;
      if (n == kDead || renumber[n] == kNoStateId) continue;
      Arc a = d.arcs[i];
      a.nextstate = renumber[n];
      sv.arcs.push_back(a);
    }
    saved.push_back(std::move(sv));
  }
  StateId new_start = renumber[fst->Start()];
  fst->DeleteStates();
  for (StateId s = 0; s < kept; ++s) fst->AddState();
  for (StateId s = 0; s < kept; ++s) {
    fst->SetFinal(s, saved[s].final);
    for (const Arc &a : saved[s].arcs) fst->AddArc(s, a);
  }
  if (new_start != kNoStateId) fst->SetStart(new_start);
}

}  // namespace fst

#endif  // KALDI_DECODER_B200_MINIFST_KALDIFST_CSRC_REMOVE_EPS_LOCAL_H_
