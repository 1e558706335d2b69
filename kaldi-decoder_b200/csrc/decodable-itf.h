// kaldi-decoder_b200/csrc/decodable-itf.h
//
// The acoustic-score source the decoder pulls from: same interface as the
// reference's DecodableInterface (kaldi-decoder/csrc/decodable-itf.h:65-102).
// Indices are one-based (index 0 is epsilon in the graph), frames zero-based.
#ifndef KALDI_DECODER_B200_CSRC_DECODABLE_ITF_H_
#define KALDI_DECODER_B200_CSRC_DECODABLE_ITF_H_

#include <cstdint>

#include "kaldi-decoder_b200/csrc/log.h"

namespace kaldi_decoder {

class DecodableInterface {
 public:
  virtual ~DecodableInterface() = default;

  /// Log-likelihood of (frame, index); the decoder negates it.
  virtual float LogLikelihood(int32_t frame, int32_t index) = 0;

  /// True if `frame` is the last frame.
  virtual bool IsLastFrame(int32_t frame) const = 0;

  /// Frames available so far.
  virtual int32_t NumFramesReady() const {
    KALDI_DECODER_ERR << "NumFramesReady() not implemented for this decodable type.";
    return -1;
  }

  /// Number of indices (they run from 1 to NumIndices()).
  virtual int32_t NumIndices() const = 0;
};

}  // namespace kaldi_decoder

#endif  // KALDI_DECODER_B200_CSRC_DECODABLE_ITF_H_
