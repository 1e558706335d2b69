// kaldi-decoder_b200/csrc/decodable-ctc.cc
#include "kaldi-decoder_b200/csrc/decodable-ctc.h"

namespace kaldi_decoder {

DecodableCtc::DecodableCtc(const FloatMatrix &log_probs, int32_t offset)
    : log_probs_(log_probs), offset_(offset) {
  p_ = log_probs_.data.data();
  num_rows_ = log_probs_.rows();
  num_cols_ = log_probs_.cols();
}

DecodableCtc::DecodableCtc(const float *p, int32_t num_rows, int32_t num_cols, int32_t offset)
    : p_(p), num_rows_(num_rows), num_cols_(num_cols), offset_(offset) {}

float DecodableCtc::LogLikelihood(int32_t frame, int32_t index) {
  // the graph's input labels are token ids + 1 (0 is epsilon): column index - 1
  return p_[static_cast<int64_t>(frame - offset_) * num_cols_ + index - 1];
}

int32_t DecodableCtc::NumFramesReady() const { return offset_ + num_rows_; }

int32_t DecodableCtc::NumIndices() const { return num_cols_; }

bool DecodableCtc::IsLastFrame(int32_t frame) const { return frame == NumFramesReady() - 1; }

}  // namespace kaldi_decoder
