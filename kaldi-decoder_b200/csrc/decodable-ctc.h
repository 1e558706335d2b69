// kaldi-decoder_b200/csrc/decodable-ctc.h
//
// DecodableCtc: a row-major [frames x vocab] matrix of log-probs
// (reference: kaldi-decoder/csrc/decodable-ctc.{h,cc}).  LogLikelihood(f, i) =
// p[(f - offset) * cols + i - 1]; NumFramesReady() = offset + rows.  Eigen is
// not available here, so the copying constructor takes a plain pointer (the
// Python binding hands it the numpy buffer); the borrowing constructor has the
// reference's exact signature (decodable-ctc.h:18-24).
#ifndef KALDI_DECODER_B200_CSRC_DECODABLE_CTC_H_
#define KALDI_DECODER_B200_CSRC_DECODABLE_CTC_H_

#include <cstdint>
#include <vector>

#include "kaldi-decoder_b200/csrc/decodable-itf.h"

namespace kaldi_decoder {

// Minimal stand-in for the reference's Eigen FloatMatrix typedef (eigen.h:11-12).
struct FloatMatrix {
  int32_t num_rows = 0;
  int32_t num_cols = 0;
  std::vector<float> data;  // row-major
  FloatMatrix() = default;
  FloatMatrix(const float *p, int32_t r, int32_t c)
      : num_rows(r), num_cols(c), data(p, p + static_cast<size_t>(r) * c) {}
  int32_t rows() const { return num_rows; }
  int32_t cols() const { return num_cols; }
  const float &operator()(int32_t r, int32_t c) const {
    return data[static_cast<size_t>(r) * num_cols + c];
  }
};

class DecodableCtc : public DecodableInterface {
 public:
  // Copies log_probs.
  explicit DecodableCtc(const FloatMatrix &log_probs, int32_t offset = 0);

  // Shares memory with the caller: `p` must outlive this object.
  DecodableCtc(const float *p, int32_t num_rows, int32_t num_cols, int32_t offset = 0);

  float LogLikelihood(int32_t frame, int32_t index) override;
  int32_t NumFramesReady() const override;
  int32_t NumIndices() const override;  // one-based indices
  bool IsLastFrame(int32_t frame) const override;

  // Direct view for the GPU decoder (no per-arc virtual call).
  const float *Data() const { return p_; }
  int32_t NumRows() const { return num_rows_; }
  int32_t NumCols() const { return num_cols_; }
  int32_t Offset() const { return offset_; }

 private:
  FloatMatrix log_probs_;
  const float *p_ = nullptr;
  int32_t num_rows_ = 0;
  int32_t num_cols_ = 0;
  int32_t offset_ = 0;
};

}  // namespace kaldi_decoder

#endif  // KALDI_DECODER_B200_CSRC_DECODABLE_CTC_H_
