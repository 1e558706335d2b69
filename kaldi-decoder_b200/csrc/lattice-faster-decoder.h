// kaldi-decoder_b200/csrc/lattice-faster-decoder.h
//
// LatticeFasterDecoderConfig, the options struct of the reference's lattice decoder
// (kaldi-decoder/csrc/lattice-faster-decoder.h:23-134): same fields, defaults, ToString()
// and Check().  Only the struct is provided: scripts that build
// or print it keep working; the lattice-generating search itself is not on the accelerated
// path (FasterDecoder's best-path search is).
#ifndef KALDI_DECODER_B200_CSRC_LATTICE_FASTER_DECODER_H_
#define KALDI_DECODER_B200_CSRC_LATTICE_FASTER_DECODER_H_

#include <cstdint>
#include <limits>
#include <sstream>
#include <string>

#include "kaldi-decoder_b200/csrc/log.h"

namespace kaldi_decoder {

struct LatticeFasterDecoderConfig {
  float beam;
  int32_t max_active;
  int32_t min_active;
  float lattice_beam;
  int32_t prune_interval;
  bool determinize_lattice;  // read by callers, not by a decoder
  float beam_delta;
  float hash_ratio;
  float prune_scale;  // (0, 1): how eagerly tokens are pruned as decoding goes
  int32_t memory_pool_tokens_block_size;
  int32_t memory_pool_links_block_size;

  LatticeFasterDecoderConfig(float beam = 16.0,
                             int32_t max_active = std::numeric_limits<int32_t>::max(),
                             int32_t min_active = 200, float lattice_beam = 10.0,
                             int32_t prune_interval = 25, bool determinize_lattice = true,
                             float beam_delta = 0.5, float hash_ratio = 2.0,
                             float prune_scale = 0.1,
                             int32_t memory_pool_tokens_block_size = 1 << 8,
                             int32_t memory_pool_links_block_size = 1 << 8)
      : beam(beam), max_active(max_active), min_active(min_active), lattice_beam(lattice_beam),
        prune_interval(prune_interval), determinize_lattice(determinize_lattice),
        beam_delta(beam_delta), hash_ratio(hash_ratio), prune_scale(prune_scale),
        memory_pool_tokens_block_size(memory_pool_tokens_block_size),
        memory_pool_links_block_size(memory_pool_links_block_size) {}

  std::string ToString() const {
    std::ostringstream os;
    os << "LatticeFasterDecoderConfig(beam=" << beam << ", max_active=" << max_active
       << ", min_active=" << min_active << ", lattice_beam=" << lattice_beam
       << ", prune_interval=" << prune_interval
       << ", determinize_lattice=" << (determinize_lattice ? "True" : "False")
       << ", beam_delta=" << beam_delta << ", hash_ratio=" << hash_ratio
       << ", prune_scale=" << prune_scale
       << ", memory_pool_tokens_block_size=" << memory_pool_tokens_block_size
       << ", memory_pool_links_block_size=" << memory_pool_links_block_size << ")";
    return os.str();
  }

  // lattice-faster-decoder.h:127-133 of the reference
  void Check() const {
    KALDI_DECODER_ASSERT(beam > 0.0 && max_active > 1 && lattice_beam > 0.0 &&
                         min_active <= max_active && prune_interval > 0 && beam_delta > 0.0 &&
                         hash_ratio >= 1.0 && prune_scale > 0.0 && prune_scale < 1.0);
  }
};

}  // namespace kaldi_decoder

#endif  // KALDI_DECODER_B200_CSRC_LATTICE_FASTER_DECODER_H_
