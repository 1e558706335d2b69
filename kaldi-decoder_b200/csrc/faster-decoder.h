// kaldi-decoder_b200/csrc/faster-decoder.h
//
// kaldi_decoder::FasterDecoder with the reference's public interface
// (kaldi-decoder/csrc/faster-decoder.h:24-107), running on a B200 through the C
// ABI of include/kd_capi.h.  What changes underneath:
//   * the `const fst::Fst<StdArc>&` is read once at construction and copied to
//     the GPU as a split CSR (the reference keeps borrowing it, h:179);
//   * tokens, the state->token map and the backpointers live in device memory
//     between calls; InitDecoding / AdvanceDecoding / GetBestPath may be
//     interleaved exactly as with the reference (streaming);
//   * a DecodableCtc is consumed as a whole matrix; any other
//     DecodableInterface is materialised frame by frame through
//     LogLikelihood() before upload.
// One FasterDecoder is one utterance lane.  BatchFasterDecoder (additive API)
// drives many lanes over one graph replica with one kernel launch.
#ifndef KALDI_DECODER_B200_CSRC_FASTER_DECODER_H_
#define KALDI_DECODER_B200_CSRC_FASTER_DECODER_H_

#include <cstdint>
#include <limits>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "fst/fst.h"
#include "fst/fstlib.h"
#include "kaldi-decoder_b200/csrc/decodable-ctc.h"
#include "kaldi-decoder_b200/csrc/decodable-itf.h"
#include "kaldifst/csrc/lattice-weight.h"

namespace kaldi_decoder {

// faster-decoder.h:24-63 of the reference: same fields, defaults and ToString().
struct FasterDecoderOptions {
  float beam;
  int32_t max_active;
  int32_t min_active;
  float beam_delta;
  float hash_ratio;  // validated (>= 1.0) and otherwise unused on the GPU

  /*implicit*/ FasterDecoderOptions(float beam = 16.0,
                                    int32_t max_active = std::numeric_limits<int32_t>::max(),
                                    int32_t min_active = 20, float beam_delta = 0.5,
                                    float hash_ratio = 2.0)
      : beam(beam),
        max_active(max_active),
        min_active(min_active),
        beam_delta(beam_delta),
        hash_ratio(hash_ratio) {}

  std::string ToString() const {
    std::ostringstream os;
    os << "FasterDecoderOptions(";
    os << "beam=" << beam << ", ";
    os << "max_active=" << max_active << ", ";
    os << "min_active=" << min_active << ", ";
    os << "beam_delta=" << beam_delta << ", ";
    os << "hash_ratio=" << hash_ratio << ")";
    return os.str();
  }
};

// Device capacities of a decoder (0 = library default); see kd_decoder_config.
struct DeviceConfig {
  int32_t device = 0;
  int32_t hash_capacity = 0;
  int64_t arena_records = 0;
  int32_t threads_per_lane = 0;
  int32_t chunk_frames = 0;
};

// The decoding graph resident on one GPU; immutable, shareable between decoders.
class DeviceGraph {
 public:
  DeviceGraph(const fst::Fst<fst::StdArc> &fst, int32_t device = 0);
  ~DeviceGraph();
  DeviceGraph(const DeviceGraph &) = delete;
  DeviceGraph &operator=(const DeviceGraph &) = delete;
  void *Handle() const { return handle_; }
  int32_t Device() const { return device_; }

  // The device copy of `fst`, shared by all decoders that are built from the same FST content.
  // Scripts in the reference's style construct a FasterDecoder from the HLG for every utterance
  // (the reference's constructor only stores a reference, faster-decoder.cc:21-32); with this
  // only the first one converts and uploads the graph.  It needs an FST type that can identify
  // its content (minifst: Fst::ContentId(), which changes when the FST is modified); any other
  // type is uploaded afresh on every call.  The last few graphs are kept alive (environment
  // KD_B200_GRAPH_CACHE = how many, default 4, 0 = no sharing); ClearCache() lets go of them.
  static std::shared_ptr<DeviceGraph> Shared(const fst::Fst<fst::StdArc> &fst, int32_t device = 0);
  static void ClearCache();
  static int64_t NumUploads();  // graphs converted and copied to a device so far (diagnostics)

  // Idle decoders (kd_decoder*) of this graph, handed from one decoder object to the next one
  // with the same capacities (`key`); PutIdle resets the decoder (or destroys it if enough are
  // waiting already).
  void *TakeIdle(const std::string &key);
  void PutIdle(const std::string &key, void *decoder);

 private:
  struct Idle;
  void *handle_ = nullptr;
  int32_t device_ = 0;
  std::unique_ptr<Idle> idle_;
};

class FasterDecoder {
 public:
  typedef fst::StdArc Arc;
  typedef Arc::Label Label;
  typedef Arc::StateId StateId;
  typedef Arc::Weight Weight;

  FasterDecoder(const fst::Fst<fst::StdArc> &fst, const FasterDecoderOptions &config);
  // additive: share an already uploaded graph / choose device capacities
  FasterDecoder(std::shared_ptr<DeviceGraph> graph, const FasterDecoderOptions &config,
                const DeviceConfig &dev = DeviceConfig());

  FasterDecoder(const FasterDecoder &) = delete;
  FasterDecoder &operator=(const FasterDecoder &) = delete;
  ~FasterDecoder();

  void SetOptions(const FasterDecoderOptions &config);

  void Decode(DecodableInterface *decodable);

  /// True if a final state was active on the last frame.
  bool ReachedFinal() const;

  /// Best path as a linear lattice; false (and an empty FST) if no token survived.
  bool GetBestPath(fst::MutableFst<fst::LatticeArc> *fst_out, bool use_final_probs = true);

  void InitDecoding();

  /// Decodes the frames that are ready, at most max_num_frames if it is >= 0.
  void AdvanceDecoding(DecodableInterface *decodable, int32_t max_num_frames = -1);

  int32_t NumFramesDecoded() const;

 private:
  struct Impl;
  std::unique_ptr<Impl> impl_;
};

// Many utterance lanes over one graph replica (additive API; the reference
// decodes one utterance per object per call).
class BatchFasterDecoder {
 public:
  BatchFasterDecoder(const fst::Fst<fst::StdArc> &fst, const FasterDecoderOptions &config,
                     int32_t max_lanes, const DeviceConfig &dev = DeviceConfig());
  BatchFasterDecoder(std::shared_ptr<DeviceGraph> graph, const FasterDecoderOptions &config,
                     int32_t max_lanes, const DeviceConfig &dev = DeviceConfig());
  BatchFasterDecoder(const BatchFasterDecoder &) = delete;
  BatchFasterDecoder &operator=(const BatchFasterDecoder &) = delete;
  ~BatchFasterDecoder();

  int32_t MaxLanes() const;
  void SetOptions(const FasterDecoderOptions &config);
  void InitDecoding(const std::vector<int32_t> &lanes);
  // One DecodableCtc-shaped matrix per lane: mats[i] is rows[i] x cols, first row = frame
  // offsets[i] (offsets may be empty).  device_memory says where the matrices live.
  void AdvanceDecoding(const std::vector<int32_t> &lanes, const std::vector<const float *> &mats,
                       const std::vector<int32_t> &rows, int32_t cols,
                       const std::vector<int32_t> &offsets = {}, int32_t max_num_frames = -1,
                       bool device_memory = false);
  // InitDecoding + AdvanceDecoding over all frames (+ the best-path selection): one launch.
  void Decode(const std::vector<int32_t> &lanes, const std::vector<const float *> &mats,
              const std::vector<int32_t> &rows, int32_t cols, bool device_memory = false);
  // The same, deferred: returns a ticket at once.  Several calls on disjoint lanes may be in
  // flight (the upload and search of one batch overlap the search and download of another);
  // any other method that touches a lane completes the call that owns it first.  Host
  // matrices must stay valid until then.  producer_stream (device matrices): the
  // cudaStream_t the matrices were produced on (nullptr: the legacy default stream).
  int64_t DecodeAsync(const std::vector<int32_t> &lanes, const std::vector<const float *> &mats,
                      const std::vector<int32_t> &rows, int32_t cols, bool device_memory = false,
                      void *producer_stream = nullptr);
  void Wait(int64_t ticket);
  // Best paths of a DecodeAsync call, in call order, without any further kernel.
  void GetResults(int64_t ticket, std::vector<int32_t> *lanes, std::vector<fst::Lattice> *out,
                  std::vector<bool> *ok, bool use_final_probs = true);
  int32_t NumFramesDecoded(int32_t lane) const;
  bool ReachedFinal(int32_t lane) const;
  bool GetBestPath(int32_t lane, fst::MutableFst<fst::LatticeArc> *fst_out,
                   bool use_final_probs = true);
  // All lanes at once: ok[i] as GetBestPath's return value.
  void GetBestPaths(const std::vector<int32_t> &lanes, std::vector<fst::Lattice> *out,
                    std::vector<bool> *ok, bool use_final_probs = true);
  void *Handle() const;

 private:
  struct Impl;
  std::unique_ptr<Impl> impl_;
};

}  // namespace kaldi_decoder

#endif  // KALDI_DECODER_B200_CSRC_FASTER_DECODER_H_
