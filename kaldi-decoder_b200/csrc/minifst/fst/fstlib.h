// minifst/fst/fstlib.h
//
// OpenFst's umbrella header.  The decoder hot path needs nothing beyond
// fst/fst.h (see that file for the list of names and the reference use
// sites); this header exists so `#include "fst/fstlib.h"`
// (faster-decoder.h:17 in the reference) resolves.
#ifndef KALDI_DECODER_B200_MINIFST_FST_FSTLIB_H_
#define KALDI_DECODER_B200_MINIFST_FST_FSTLIB_H_

#include "fst/fst.h"

#endif  // KALDI_DECODER_B200_MINIFST_FST_FSTLIB_H_
