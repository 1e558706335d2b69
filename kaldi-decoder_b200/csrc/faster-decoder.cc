// kaldi-decoder_b200/csrc/faster-decoder.cc
//
// Host side of FasterDecoder / BatchFasterDecoder: graph ingest, decodable
// handling and best-path lattice construction around the C ABI (kd_capi.h).
// The search itself is in kd_kernels.cuh; nothing here decodes on the CPU.

#include "kaldi-decoder_b200/csrc/faster-decoder.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <list>
#include <mutex>
#include <utility>

#include "kaldi-decoder_b200/csrc/log.h"
#include "kaldifst/csrc/remove-eps-local.h"
#include "kd_capi.h"

namespace kaldi_decoder {

namespace {

void Check(int rc) {
  if (rc != KD_OK) KALDI_DECODER_ERR << kd_last_error();
}

kd_options ToC(const FasterDecoderOptions &o) {
  kd_options c;
  c.beam = o.beam;
  c.max_active = o.max_active;
  c.min_active = o.min_active;
  c.beam_delta = o.beam_delta;
  c.hash_ratio = o.hash_ratio;
  return c;
}

kd_decoder_config ToC(const DeviceConfig &d, int32_t max_lanes) {
  kd_decoder_config c;
  c.search = KD_SEARCH_FASTER;
  c.max_lanes = max_lanes;
  c.hash_capacity = d.hash_capacity;
  c.arena_records = d.arena_records;
  c.threads_per_lane = d.threads_per_lane;
  c.chunk_frames = d.chunk_frames;
  if (const char *e = std::getenv("KD_B200_ARENA_RECORDS"))
    if (c.arena_records == 0) c.arena_records = std::atoll(e);
  if (const char *e = std::getenv("KD_B200_HASH_CAPACITY"))
    if (c.hash_capacity == 0) c.hash_capacity = std::atoi(e);
  return c;
}

// What makes two decoders of one graph interchangeable (DeviceGraph::TakeIdle / PutIdle).
std::string IdleKey(const kd_decoder_config &c) {
  char buf[128];
  std::snprintf(buf, sizeof(buf), "%d/%d/%d/%lld/%d/%d", c.search, c.max_lanes, c.hash_capacity,
                static_cast<long long>(c.arena_records), c.threads_per_lane, c.chunk_frames);
  return buf;
}

// ContentId() of FST types that have one (minifst), 0 otherwise (OpenFst: no graph sharing).
template <class F>
auto ContentIdOf(const F &f, int) -> decltype(static_cast<uint64_t>(f.ContentId())) {
  return f.ContentId();
}
template <class F>
uint64_t ContentIdOf(const F &, long) {
  return 0;
}

struct GraphCache {
  struct Slot {
    uint64_t content;
    int32_t device;
    std::shared_ptr<DeviceGraph> graph;
  };
  std::mutex mu;
  std::list<Slot> slots;  // most recently used first
};

// (never destroyed: at process exit the CUDA runtime may already be gone)
GraphCache &TheGraphCache() {
  static GraphCache *c = new GraphCache;
  return *c;
}

size_t GraphCacheSize() {
  if (const char *e = std::getenv("KD_B200_GRAPH_CACHE")) return static_cast<size_t>(std::max(0, std::atoi(e)));
  return 4;
}

// how many idle decoders a graph keeps for the next decoder objects (0 = none: destroy at once)
size_t MaxIdleDecoders() {
  if (const char *e = std::getenv("KD_B200_IDLE_DECODERS")) return static_cast<size_t>(std::max(0, std::atoi(e)));
  return 16;
}

std::atomic<int64_t> g_uploads{0};

// Linear lattice from the raw per-token arcs, then RemoveEpsLocal, as
// faster-decoder.cc:408-422 of the reference.
void BuildLattice(int64_t n, const int32_t *il, const int32_t *ol, const float *gw,
                  const float *aw, const float final2[2], fst::MutableFst<fst::LatticeArc> *out) {
  out->DeleteStates();
  auto cur = out->AddState();
  out->SetStart(cur);
  for (int64_t i = 0; i < n; ++i) {
    fst::LatticeArc arc(il[i], ol[i], fst::LatticeWeight(gw[i], aw[i]), 0);
    arc.nextstate = out->AddState();
    out->AddArc(cur, arc);
    cur = arc.nextstate;
  }
  out->SetFinal(cur, fst::LatticeWeight(final2[0], final2[1]));
  fst::RemoveEpsLocal(out);
}

}  // namespace

// --------------------------------------------------------------- DeviceGraph

struct DeviceGraph::Idle {
  std::mutex mu;
  std::vector<std::pair<std::string, kd_decoder *>> list;
};

DeviceGraph::DeviceGraph(const fst::Fst<fst::StdArc> &fst, int32_t device)
    : device_(device), idle_(new Idle) {
  // Only what OpenFst's abstract Fst<Arc> offers: CountStates, Start, Final, ArcIterator
  // (NumStates() lives on ExpandedFst; the reference binds `const Fst<StdArc>&`).
  const int32_t n = fst::CountStates(fst);
  std::vector<int64_t> off(static_cast<size_t>(std::max(n, 0)) + 1, 0);
  std::vector<int32_t> il, ol, ns;
  std::vector<float> w, fin(static_cast<size_t>(std::max(n, 0)));
  for (int32_t s = 0; s < n; ++s) {
    for (fst::ArcIterator<fst::Fst<fst::StdArc>> aiter(fst, s); !aiter.Done(); aiter.Next()) {
      const fst::StdArc &arc = aiter.Value();
      il.push_back(arc.ilabel);
      ol.push_back(arc.olabel);
      w.push_back(arc.weight.Value());
      ns.push_back(arc.nextstate);
    }
    off[s + 1] = static_cast<int64_t>(il.size());
    fin[s] = fst.Final(s).Value();
  }
  kd_graph *g = nullptr;
  Check(kd_graph_create(device, n, fst.Start(), off.data(), il.data(), ol.data(), w.data(),
                        ns.data(), fin.data(), &g));
  handle_ = g;
  g_uploads.fetch_add(1, std::memory_order_relaxed);
}

int64_t DeviceGraph::NumUploads() { return g_uploads.load(std::memory_order_relaxed); }

DeviceGraph::~DeviceGraph() {
  for (auto &e : idle_->list) kd_decoder_destroy(e.second);
  kd_graph_destroy(static_cast<kd_graph *>(handle_));
}

std::shared_ptr<DeviceGraph> DeviceGraph::Shared(const fst::Fst<fst::StdArc> &fst, int32_t device) {
  const uint64_t content = ContentIdOf(fst, 0);
  const size_t keep = GraphCacheSize();
  if (content == 0 || keep == 0) return std::make_shared<DeviceGraph>(fst, device);
  GraphCache &c = TheGraphCache();
  // (held across the upload: a second thread that wants the same graph waits for it)
  std::lock_guard<std::mutex> lock(c.mu);
  for (auto it = c.slots.begin(); it != c.slots.end(); ++it) {
    if (it->content == content && it->device == device) {
      c.slots.splice(c.slots.begin(), c.slots, it);
      return c.slots.front().graph;
    }
  }
  auto g = std::make_shared<DeviceGraph>(fst, device);
  c.slots.push_front({content, device, g});
  while (c.slots.size() > keep) c.slots.pop_back();
  return g;
}

void DeviceGraph::ClearCache() {
  GraphCache &c = TheGraphCache();
  std::list<GraphCache::Slot> drop;
  {
    std::lock_guard<std::mutex> lock(c.mu);
    drop.swap(c.slots);
  }
}

void *DeviceGraph::TakeIdle(const std::string &key) {
  std::lock_guard<std::mutex> lock(idle_->mu);
  for (size_t i = idle_->list.size(); i-- > 0;) {
    if (idle_->list[i].first == key) {
      kd_decoder *d = idle_->list[i].second;
      idle_->list.erase(idle_->list.begin() + static_cast<long>(i));
      return d;
    }
  }
  return nullptr;
}

void DeviceGraph::PutIdle(const std::string &key, void *decoder) {
  auto *d = static_cast<kd_decoder *>(decoder);
  if (!d) return;
  const size_t keep = MaxIdleDecoders();
  if (keep > 0 && kd_decoder_reset(d) == KD_OK) {
    std::lock_guard<std::mutex> lock(idle_->mu);
    if (idle_->list.size() < keep) {
      idle_->list.emplace_back(key, d);
      return;
    }
  }
  kd_decoder_destroy(d);
}

// ------------------------------------------------------------- FasterDecoder

struct FasterDecoder::Impl {
  std::shared_ptr<DeviceGraph> graph;
  kd_decoder *dec = nullptr;
  std::string idle_key;
  std::vector<float> scratch;  // materialised generic decodables
  // the device side outlives this object: the next FasterDecoder of the graph takes it over
  ~Impl() {
    if (dec) graph->PutIdle(idle_key, dec);
  }
};

FasterDecoder::FasterDecoder(const fst::Fst<fst::StdArc> &fst, const FasterDecoderOptions &config)
    : FasterDecoder(DeviceGraph::Shared(fst, 0), config, DeviceConfig()) {}

FasterDecoder::FasterDecoder(std::shared_ptr<DeviceGraph> graph,
                             const FasterDecoderOptions &config, const DeviceConfig &dev)
    : impl_(new Impl) {
  impl_->graph = std::move(graph);
  kd_options o = ToC(config);
  kd_decoder_config c = ToC(dev, 1);
  impl_->idle_key = IdleKey(c);
  impl_->dec = static_cast<kd_decoder *>(impl_->graph->TakeIdle(impl_->idle_key));
  if (impl_->dec) {
    Check(kd_decoder_set_options(impl_->dec, &o));  // (on failure ~Impl puts it back)
    return;
  }
  Check(kd_decoder_create(static_cast<kd_graph *>(impl_->graph->Handle()), &o, &c, &impl_->dec));
}

FasterDecoder::~FasterDecoder() = default;

void FasterDecoder::SetOptions(const FasterDecoderOptions &config) {
  kd_options o = ToC(config);
  Check(kd_decoder_set_options(impl_->dec, &o));
}

void FasterDecoder::InitDecoding() {
  const int32_t lane = 0;
  Check(kd_decoder_init(impl_->dec, 1, &lane));
}

void FasterDecoder::Decode(DecodableInterface *decodable) {
  // DecodableCtc: InitDecoding, every ready frame and the best-path selection that
  // ReachedFinal() / GetBestPath() will ask for, in ONE kernel launch.
  if (auto *ctc = dynamic_cast<DecodableCtc *>(decodable)) {
    if (ctc->Offset() == 0) {
      const int32_t lane = 0;
      const float *p = ctc->Data();
      const int32_t rows = ctc->NumRows();
      int64_t ticket = -1;
      Check(kd_decoder_advance_async(impl_->dec, 1, &lane, &p, &rows, ctc->NumCols(), nullptr, -1,
                                     KD_MEM_HOST, KD_ADVANCE_INIT | KD_ADVANCE_FINALIZE, nullptr,
                                     &ticket));
      Check(kd_decoder_wait(impl_->dec, ticket));
      return;
    }
  }
  InitDecoding();
  AdvanceDecoding(decodable);
}

int32_t FasterDecoder::NumFramesDecoded() const {
  int32_t v = -1;
  Check(kd_decoder_num_frames_decoded(impl_->dec, 0, &v));
  return v;
}

void FasterDecoder::AdvanceDecoding(DecodableInterface *decodable, int32_t max_num_frames) {
  const int32_t decoded = NumFramesDecoded();
  KALDI_DECODER_ASSERT(decoded >= 0 && "You must call InitDecoding() before AdvanceDecoding()");
  const int32_t lane = 0;
  if (auto *ctc = dynamic_cast<DecodableCtc *>(decodable)) {
    const float *p = ctc->Data();
    const int32_t rows = ctc->NumRows(), offset = ctc->Offset();
    Check(kd_decoder_advance(impl_->dec, 1, &lane, &p, &rows, ctc->NumCols(), &offset,
                             max_num_frames, KD_MEM_HOST));
    return;
  }
  // Any other decodable: pull the ready frames through its virtual interface
  // (what the reference does arc by arc), then decode them in one call.
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
  Check(kd_decoder_advance(impl_->dec, 1, &lane, &p, &rows, cols, &decoded, -1, KD_MEM_HOST));
}

bool FasterDecoder::ReachedFinal() const {
  int32_t v = 0;
  Check(kd_decoder_reached_final(impl_->dec, 0, &v));
  return v != 0;
}

bool FasterDecoder::GetBestPath(fst::MutableFst<fst::LatticeArc> *fst_out, bool use_final_probs) {
  fst_out->DeleteStates();
  const int32_t lane = 0;
  int32_t ok = 0, rf = 0;
  int64_t n = 0;
  Check(kd_decoder_best_path_prepare(impl_->dec, 1, &lane, use_final_probs ? 1 : 0, &ok, &rf, &n));
  if (!ok) return false;
  std::vector<int32_t> il(n), ol(n);
  std::vector<float> gw(n), aw(n);
  float f2[2] = {0.f, 0.f};
  const int64_t off = 0;
  Check(kd_decoder_best_path_fetch(impl_->dec, 1, &lane, &off, n, il.data(), ol.data(), gw.data(),
                                   aw.data(), f2));
  BuildLattice(n, il.data(), ol.data(), gw.data(), aw.data(), f2, fst_out);
  return true;
}

// -------------------------------------------------------- BatchFasterDecoder

struct BatchFasterDecoder::Impl {
  std::shared_ptr<DeviceGraph> graph;
  kd_decoder *dec = nullptr;
  int32_t max_lanes = 0;
  ~Impl() { kd_decoder_destroy(dec); }
};

BatchFasterDecoder::BatchFasterDecoder(const fst::Fst<fst::StdArc> &fst,
                                       const FasterDecoderOptions &config, int32_t max_lanes,
                                       const DeviceConfig &dev)
    : BatchFasterDecoder(DeviceGraph::Shared(fst, dev.device), config, max_lanes, dev) {}

BatchFasterDecoder::BatchFasterDecoder(std::shared_ptr<DeviceGraph> graph,
                                       const FasterDecoderOptions &config, int32_t max_lanes,
                                       const DeviceConfig &dev)
    : impl_(new Impl) {
  KALDI_DECODER_ASSERT(max_lanes >= 1);
  impl_->graph = std::move(graph);
  impl_->max_lanes = max_lanes;
  kd_options o = ToC(config);
  kd_decoder_config c = ToC(dev, max_lanes);
  Check(kd_decoder_create(static_cast<kd_graph *>(impl_->graph->Handle()), &o, &c, &impl_->dec));
}

BatchFasterDecoder::~BatchFasterDecoder() = default;

int32_t BatchFasterDecoder::MaxLanes() const { return impl_->max_lanes; }
void *BatchFasterDecoder::Handle() const { return impl_->dec; }

void BatchFasterDecoder::SetOptions(const FasterDecoderOptions &config) {
  kd_options o = ToC(config);
  Check(kd_decoder_set_options(impl_->dec, &o));
}

void BatchFasterDecoder::InitDecoding(const std::vector<int32_t> &lanes) {
  Check(kd_decoder_init(impl_->dec, static_cast<int32_t>(lanes.size()), lanes.data()));
}

void BatchFasterDecoder::AdvanceDecoding(const std::vector<int32_t> &lanes,
                                         const std::vector<const float *> &mats,
                                         const std::vector<int32_t> &rows, int32_t cols,
                                         const std::vector<int32_t> &offsets,
                                         int32_t max_num_frames, bool device_memory) {
  KALDI_DECODER_ASSERT(mats.size() == lanes.size() && rows.size() == lanes.size());
  KALDI_DECODER_ASSERT(offsets.empty() || offsets.size() == lanes.size());
  Check(kd_decoder_advance(impl_->dec, static_cast<int32_t>(lanes.size()), lanes.data(),
                           mats.data(), rows.data(), cols,
                           offsets.empty() ? nullptr : offsets.data(), max_num_frames,
                           device_memory ? KD_MEM_DEVICE : KD_MEM_HOST));
}

void BatchFasterDecoder::Decode(const std::vector<int32_t> &lanes,
                                const std::vector<const float *> &mats,
                                const std::vector<int32_t> &rows, int32_t cols,
                                bool device_memory) {
  Wait(DecodeAsync(lanes, mats, rows, cols, device_memory));
}

int64_t BatchFasterDecoder::DecodeAsync(const std::vector<int32_t> &lanes,
                                        const std::vector<const float *> &mats,
                                        const std::vector<int32_t> &rows, int32_t cols,
                                        bool device_memory, void *producer_stream) {
  KALDI_DECODER_ASSERT(mats.size() == lanes.size() && rows.size() == lanes.size());
  int64_t ticket = -1;
  Check(kd_decoder_advance_async(impl_->dec, static_cast<int32_t>(lanes.size()), lanes.data(),
                                 mats.data(), rows.data(), cols, nullptr, -1,
                                 device_memory ? KD_MEM_DEVICE : KD_MEM_HOST,
                                 KD_ADVANCE_INIT | KD_ADVANCE_FINALIZE, producer_stream, &ticket));
  return ticket;
}

void BatchFasterDecoder::Wait(int64_t ticket) { Check(kd_decoder_wait(impl_->dec, ticket)); }

void BatchFasterDecoder::GetResults(int64_t ticket, std::vector<int32_t> *lanes,
                                    std::vector<fst::Lattice> *out, std::vector<bool> *ok,
                                    bool use_final_probs) {
  int32_t n = 0;
  const int32_t *lane_ids = nullptr, *words = nullptr;
  const int64_t *woff = nullptr;
  std::vector<int64_t> cnt(impl_->max_lanes);
  std::vector<int32_t> okv(impl_->max_lanes), rf(impl_->max_lanes);
  std::vector<float> f2(2 * static_cast<size_t>(impl_->max_lanes));
  Check(kd_decoder_result_view(impl_->dec, ticket, use_final_probs ? 1 : 0, &n, &lane_ids, &words,
                               &woff, cnt.data(), okv.data(), rf.data(), f2.data()));
  if (lanes) lanes->assign(lane_ids, lane_ids + n);
  out->assign(n, fst::Lattice());
  ok->assign(n, false);
  for (int32_t i = 0; i < n; ++i) {
    if (!okv[i]) continue;
    (*ok)[i] = true;
    BuildLattice(cnt[i], words + woff[4 * i], words + woff[4 * i + 1],
                 reinterpret_cast<const float *>(words + woff[4 * i + 2]),
                 reinterpret_cast<const float *>(words + woff[4 * i + 3]), &f2[2 * i], &(*out)[i]);
  }
}

int32_t BatchFasterDecoder::NumFramesDecoded(int32_t lane) const {
  int32_t v = -1;
  Check(kd_decoder_num_frames_decoded(impl_->dec, lane, &v));
  return v;
}

bool BatchFasterDecoder::ReachedFinal(int32_t lane) const {
  int32_t v = 0;
  Check(kd_decoder_reached_final(impl_->dec, lane, &v));
  return v != 0;
}

bool BatchFasterDecoder::GetBestPath(int32_t lane, fst::MutableFst<fst::LatticeArc> *fst_out,
                                     bool use_final_probs) {
  std::vector<fst::Lattice> out;
  std::vector<bool> ok;
  GetBestPaths({lane}, &out, &ok, use_final_probs);
  fst_out->DeleteStates();
  if (!ok[0]) return false;
  // copy through the MutableFst interface
  const fst::Lattice &l = out[0];
  for (int s = 0; s < l.NumStates(); ++s) fst_out->AddState();
  if (l.Start() != fst::kNoStateId) fst_out->SetStart(l.Start());
  for (int s = 0; s < l.NumStates(); ++s) {
    fst_out->SetFinal(s, l.Final(s));
    for (fst::ArcIterator<fst::Lattice> aiter(l, s); !aiter.Done(); aiter.Next())
      fst_out->AddArc(s, aiter.Value());
  }
  return true;
}

void BatchFasterDecoder::GetBestPaths(const std::vector<int32_t> &lanes,
                                      std::vector<fst::Lattice> *out, std::vector<bool> *ok,
                                      bool use_final_probs) {
  const int32_t n = static_cast<int32_t>(lanes.size());
  std::vector<int32_t> okv(n), rf(n);
  std::vector<int64_t> cnt(n), off(n);
  Check(kd_decoder_best_path_prepare(impl_->dec, n, lanes.data(), use_final_probs ? 1 : 0,
                                     okv.data(), rf.data(), cnt.data()));
  int64_t total = 0;
  for (int32_t i = 0; i < n; ++i) {
    off[i] = total;
    total += cnt[i];
  }
  std::vector<int32_t> il(total), ol(total);
  std::vector<float> gw(total), aw(total), f2(2 * static_cast<size_t>(n));
  Check(kd_decoder_best_path_fetch(impl_->dec, n, lanes.data(), off.data(), total, il.data(),
                                   ol.data(), gw.data(), aw.data(), f2.data()));
  out->assign(n, fst::Lattice());
  ok->assign(n, false);
  for (int32_t i = 0; i < n; ++i) {
    if (!okv[i]) continue;
    (*ok)[i] = true;
    BuildLattice(cnt[i], il.data() + off[i], ol.data() + off[i], gw.data() + off[i],
                 aw.data() + off[i], &f2[2 * i], &(*out)[i]);
  }
}

}  // namespace kaldi_decoder
