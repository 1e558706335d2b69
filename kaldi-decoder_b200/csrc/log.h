// kaldi-decoder_b200/csrc/log.h
//
// Error convention of the reference (kaldi-decoder/csrc/log.h:21-96), restated:
// KALDI_DECODER_LOG / KALDI_DECODER_WARN build a message and drop it (the
// reference's print is commented out, log.h:52); KALDI_DECODER_ERR and a failed
// KALDI_DECODER_ASSERT throw std::runtime_error carrying "file:function:line"
// and the message, which pybind11 turns into a Python RuntimeError.
#ifndef KALDI_DECODER_B200_CSRC_LOG_H_
#define KALDI_DECODER_B200_CSRC_LOG_H_

#include <cstdint>
#include <sstream>
#include <stdexcept>
#include <string>

namespace kaldi_decoder {

enum class LogLevel { kInfo = 0, kWarn = 1, kError = 2 };

class Logger {
 public:
  Logger(const char *file, const char *func, uint32_t line, LogLevel level) : level_(level) {
    static const char *const kTag[] = {"[I] ", "[W] ", "[E] "};
    os_ << file << ":" << func << ":" << line << "\n" << kTag[static_cast<int>(level)];
  }
  template <typename T>
  Logger &operator<<(const T &v) {
    os_ << v;
    return *this;
  }
  ~Logger() noexcept(false) {
    if (level_ == LogLevel::kError) throw std::runtime_error(os_.str());
  }

 private:
  std::ostringstream os_;
  LogLevel level_;
};

struct Voidifier {
  void operator&(const Logger &) const {}
};

}  // namespace kaldi_decoder

#if defined(__GNUC__) || defined(__clang__)
#define KALDI_DECODER_FUNC __PRETTY_FUNCTION__
#else
#define KALDI_DECODER_FUNC __func__
#endif

#define KALDI_DECODER_LOG \
  ::kaldi_decoder::Logger(__FILE__, KALDI_DECODER_FUNC, __LINE__, ::kaldi_decoder::LogLevel::kInfo)
#define KALDI_DECODER_WARN \
  ::kaldi_decoder::Logger(__FILE__, KALDI_DECODER_FUNC, __LINE__, ::kaldi_decoder::LogLevel::kWarn)
#define KALDI_DECODER_ERR \
  ::kaldi_decoder::Logger(__FILE__, KALDI_DECODER_FUNC, __LINE__, ::kaldi_decoder::LogLevel::kError)
#define KALDI_DECODER_ASSERT(x) \
  (x) ? (void)0 : ::kaldi_decoder::Voidifier() & KALDI_DECODER_ERR << "Check failed!\n" << "x: " << #x

#endif  // KALDI_DECODER_B200_CSRC_LOG_H_
