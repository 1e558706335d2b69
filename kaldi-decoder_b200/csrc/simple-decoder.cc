// kaldi-decoder_b200/csrc/simple-decoder.cc
//
// Host side of SimpleDecoder around the C ABI (kd_capi.h); the search runs in
// kd_kernels.cuh with SIMPLE = true.  Nothing here decodes on the CPU.

#include "kaldi-decoder_b200/csrc/simple-decoder.h"

#include <algorithm>
#include <limits>
#include <utility>
#include <vector>

#include "kaldi-decoder_b200/csrc/log.h"
#include "kaldifst/csrc/remove-eps-local.h"
#include "kd_capi.h"

namespace kaldi_decoder {

namespace {
void CheckRc(int rc) {
  if (rc != KD_OK) KALDI_DECODER_ERR << kd_last_error();
}
}  // namespace

struct SimpleDecoder::Impl {
  std::shared_ptr<DeviceGraph> graph;
  kd_decoder *dec = nullptr;
  std::vector<float> scratch;  // materialised generic decodables
  ~Impl() { kd_decoder_destroy(dec); }
};

SimpleDecoder::SimpleDecoder(const fst::Fst<fst::StdArc> &fst, float beam)
    : SimpleDecoder(DeviceGraph::Shared(fst, 0), beam, DeviceConfig()) {}

SimpleDecoder::SimpleDecoder(std::shared_ptr<DeviceGraph> graph, float beam,
                             const DeviceConfig &dev)
    : impl_(new Impl) {
  impl_->graph = std::move(graph);
  kd_options o;
  o.beam = beam;
  o.max_active = std::numeric_limits<int32_t>::max();
  o.min_active = 0;
  o.beam_delta = 0.5f;
  o.hash_ratio = 2.0f;
  kd_decoder_config c;
  c.max_lanes = 1;
  c.hash_capacity = dev.hash_capacity;
  c.arena_records = dev.arena_records;
  c.threads_per_lane = 0;
  c.chunk_frames = dev.chunk_frames;
  c.search = KD_SEARCH_SIMPLE;
  CheckRc(kd_decoder_create(static_cast<kd_graph *>(impl_->graph->Handle()), &o, &c, &impl_->dec));
}

SimpleDecoder::~SimpleDecoder() = default;

void SimpleDecoder::InitDecoding() {
  const int32_t lane = 0;
  CheckRc(kd_decoder_init(impl_->dec, 1, &lane));
}

int32_t SimpleDecoder::NumFramesDecoded() const {
  int32_t v = -1;
  CheckRc(kd_decoder_num_frames_decoded(impl_->dec, 0, &v));
  return v;
}

bool SimpleDecoder::Decode(DecodableInterface *decodable) {
  InitDecoding();
  AdvanceDecoding(decodable);
  // simple-decoder.cc:27: true iff tokens are alive; the best token always survives
  // PruneToks, so this is "a best token exists"
  int32_t ok = 0, rf = 0;
  int64_t n = 0;
  const int32_t lane = 0;
  CheckRc(kd_decoder_best_path_prepare(impl_->dec, 1, &lane, 1, &ok, &rf, &n));
  return ok != 0;
}

void SimpleDecoder::AdvanceDecoding(DecodableInterface *decodable, int32_t max_num_frames) {
  const int32_t decoded = NumFramesDecoded();
  KALDI_DECODER_ASSERT(decoded >= 0 && "You must call InitDecoding() before AdvanceDecoding()");
  const int32_t lane = 0;
  if (auto *ctc = dynamic_cast<DecodableCtc *>(decodable)) {
    const float *p = ctc->Data();
    const int32_t rows = ctc->NumRows(), offset = ctc->Offset();
    CheckRc(kd_decoder_advance(impl_->dec, 1, &lane, &p, &rows, ctc->NumCols(), &offset,
                               max_num_frames, KD_MEM_HOST));
    return;
  }
  const int32_t ready = decodable->NumFramesReady();
  KALDI_DECODER_ASSERT(ready >= decoded);
  int32_t target = ready;
  if (max_num_frames >= 0) target = std::min(target, decoded + max_num_frames);
  if (target <= decoded) return;
  const int32_t cols = decodable->NumIndices();
  const int32_t rows = target - decoded;
  impl_->scratch.resize(static_cast<size_t>(rows) * cols);
  for (int32_t f = 0; f < rows; ++f)
    for (int32_t i = 0; i < cols; ++i)
      impl_->scratch[static_cast<size_t>(f) * cols + i] =
          decodable->LogLikelihood(decoded + f, i + 1);
  const float *p = impl_->scratch.data();
  CheckRc(kd_decoder_advance(impl_->dec, 1, &lane, &p, &rows, cols, &decoded, -1, KD_MEM_HOST));
}

bool SimpleDecoder::ReachedFinal() const {
  int32_t v = 0;
  CheckRc(kd_decoder_reached_final(impl_->dec, 0, &v));
  return v != 0;
}

float SimpleDecoder::FinalRelativeCost() const {
  float v = std::numeric_limits<float>::infinity();
  CheckRc(kd_decoder_final_relative_cost(impl_->dec, 0, &v));
  return v;
}

bool SimpleDecoder::GetBestPath(fst::Lattice *fst_out, bool use_final_probs) const {
  fst_out->DeleteStates();
  const int32_t lane = 0;
  int32_t ok = 0, rf = 0;
  int64_t n = 0;
  CheckRc(kd_decoder_best_path_prepare(impl_->dec, 1, &lane, use_final_probs ? 1 : 0, &ok, &rf, &n));
  if (!ok) return false;
  std::vector<int32_t> il(n), ol(n);
  std::vector<float> gw(n), aw(n);
  float f2[2] = {0.f, 0.f};
  const int64_t off = 0;
  CheckRc(kd_decoder_best_path_fetch(impl_->dec, 1, &lane, &off, n, il.data(), ol.data(), gw.data(),
                                     aw.data(), f2));
  // simple-decoder.cc:128-147: linear lattice, final weight, RemoveEpsLocal
  auto cur = fst_out->AddState();
  fst_out->SetStart(cur);
  for (int64_t i = 0; i < n; ++i) {
    fst::LatticeArc arc(il[i], ol[i], fst::LatticeWeight(gw[i], aw[i]), 0);
    arc.nextstate = fst_out->AddState();
    fst_out->AddArc(cur, arc);
    cur = arc.nextstate;
  }
  fst_out->SetFinal(cur, fst::LatticeWeight(f2[0], f2[1]));
  fst::RemoveEpsLocal(fst_out);
  return true;
}

}  // namespace kaldi_decoder
