// kaldi-decoder_b200/csrc/simple-decoder.h
//
// kaldi_decoder::SimpleDecoder with the reference's public interface
// (kaldi-decoder/csrc/simple-decoder.h:24-79), on the same device search as
// FasterDecoder (kd_decoder_config.search = KD_SEARCH_SIMPLE): beam-only pruning,
// token cost = prev + float(graph + acoustic), PruneToks after every frame.
// Differences a caller can see:
//   * the acoustic cost of a lattice arc is recovered as float(cost - prev cost) -
//     graph cost (the log-probs are gone when GetBestPath runs): equal to the
//     reference's stored value up to one float rounding of graph + acoustic;
//   * cost ties are broken deterministically (lowest arc index / lowest state id),
//     the reference by unordered_map iteration order.
#ifndef KALDI_DECODER_B200_CSRC_SIMPLE_DECODER_H_
#define KALDI_DECODER_B200_CSRC_SIMPLE_DECODER_H_

#include <cstdint>
#include <memory>

#include "kaldi-decoder_b200/csrc/faster-decoder.h"

namespace kaldi_decoder {

class SimpleDecoder {
 public:
  using StdArc = fst::StdArc;
  using StdWeight = StdArc::Weight;
  using Label = StdArc::Label;
  using StateId = StdArc::StateId;

  SimpleDecoder(const fst::Fst<fst::StdArc> &fst, float beam);
  // additive: share an already uploaded graph / choose device capacities
  SimpleDecoder(std::shared_ptr<DeviceGraph> graph, float beam,
                const DeviceConfig &dev = DeviceConfig());
  SimpleDecoder(const SimpleDecoder &) = delete;
  SimpleDecoder &operator=(const SimpleDecoder &) = delete;
  ~SimpleDecoder();

  // InitDecoding + AdvanceDecoding over every ready frame; true iff some token is alive
  // afterwards, final state or not.
  bool Decode(DecodableInterface *decodable);

  bool ReachedFinal() const;

  // Linear lattice of the cheapest live token's history; false (and an empty FST) when no
  // token is alive.  Final weights take part only if use_final_probs and ReachedFinal().
  bool GetBestPath(fst::Lattice *fst_out, bool use_final_probs = true) const;

  // min(cost + final weight) - min(cost) over the live tokens; +inf without an active
  // final state.
  float FinalRelativeCost() const;

  void InitDecoding();

  void AdvanceDecoding(DecodableInterface *decodable, int32_t max_num_frames = -1);

  int32_t NumFramesDecoded() const;

 private:
  struct Impl;
  std::unique_ptr<Impl> impl_;
};

}  // namespace kaldi_decoder

#endif  // KALDI_DECODER_B200_CSRC_SIMPLE_DECODER_H_
