// kaldi-decoder_b200/csrc/kd_capi.cu
//
// Host side of the C ABI declared in include/kd_capi.h: graph ingest (the
// reference's `const fst::Fst<StdArc>&`, faster-decoder.cc:21-23, becomes a
// device-resident split CSR), lane/arena/table allocation, kernel launches,
// host<->device staging.  No CPU implementation of the search lives here: if
// there is no CUDA device every entry point fails.

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "kd_capi.h"
#include "kd_kernels.cuh"

namespace {

thread_local std::string g_error;

int Fail(int code, const std::string &msg) {
  g_error = msg;
  return code;
}

#define KD_CUDA(expr)                                                              \
  do {                                                                             \
    cudaError_t kd_e_ = (expr);                                                    \
    if (kd_e_ != cudaSuccess) {                                                    \
      return Fail(KD_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(kd_e_)); \
    }                                                                              \
  } while (0)

constexpr int kNumSlots = 16;  // asynchronous calls in flight per decoder
constexpr int kNumSMsFallback = 148;

template <class T>
int DevAlloc(T **p, size_t n) {
  *p = nullptr;
  if (n == 0) n = 1;
  KD_CUDA(cudaMalloc(reinterpret_cast<void **>(p), n * sizeof(T)));
  return KD_OK;
}

// Calls in flight run on up to 2 x kNumSlots streams.  The driver maps streams onto
// CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8); streams sharing a queue serialise:
// a call's copy stream stuck behind another call's persistent search kernel starves its own
// lanes.  The variable is read when the CUDA context is created, so it is set (if the
// application has not chosen a value) as soon as this library is loaded.
__attribute__((constructor)) void KdRaiseMaxConnections() {
  setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", /*overwrite=*/0);
}

}  // namespace

struct kd_graph {
  int device = 0;
  int64_t num_states = 0, num_arcs = 0, num_emit = 0, num_eps = 0;
  int32_t start = -1, max_ilabel = 0;
  // one device allocation, hottest arrays first: [e_iw | st | e_no | n_arc | fin]
  unsigned char *blob = nullptr;
  size_t blob_bytes = 0;
  int64_t num_label_tables = 0;
  int2 *labtab = nullptr;  // [num_label_tables][max_ilabel]: ilabel-1 -> {weight bits, arc offset in its state or -1}
  int4 *st = nullptr;          // 2 x int4 per state
  int2 *e_iw = nullptr;
  int2 *e_no = nullptr;
  int4 *n_arc = nullptr;
  float *fin = nullptr;
};

// One asynchronous AdvanceDecoding call in flight: its own streams, staging buffer, item
// list, progress words and result buffer, so that the upload and the search of one batch of
// lanes overlap the search and the download of another.
struct kd_slot {
  cudaStream_t sc = nullptr, sx = nullptr;  // search, copy
  cudaEvent_t ev_gate = nullptr, ev_begin = nullptr, ev_end = nullptr, ev_dep = nullptr;
  kd::AdvanceItem *d_items = nullptr, *h_items = nullptr;  // max_lanes each
  int32_t *d_words = nullptr;     // [0] work counter, [1] yield flag, [2] rows delivered
  int32_t *h_progress = nullptr;  // pinned: one value per chunk
  int32_t progress_cap = 0;
  float *d_stage = nullptr;       // host-memory advance: the lanes' rows on the device
  size_t stage_floats = 0;
  int32_t *h_res = nullptr;       // pinned: parked best paths, [lane][4 * path_cap] words (+ spill)
  size_t res_words = 0;
  long long *d_out_off = nullptr, *h_out_off = nullptr;  // spill of unparked paths
  // the call in flight
  bool busy = false;
  int64_t ticket = -1;
  int rc = 0;                 // outcome of the last completed call
  std::string msg;
  std::vector<int32_t> lanes, targets;
  kd::Params P;
  int threads = 0;
  int flags = 0;
  int32_t path_cap = 0;
  bool host_input = false;
  bool res_valid = false;     // h_res holds the results of `ticket`
  std::vector<int64_t> res_off;  // [4n] word offsets into h_res
  float kernel_ms = 0.f;
  int32_t launches = 0;
};

struct kd_decoder {
  kd_graph *g = nullptr;
  kd_options opts;
  int device = 0;
  int num_sms = kNumSMsFallback;
  size_t mem_pitch = 0;
  bool launch_blocking = false;
  int32_t max_lanes = 1;
  uint32_t hcap = 0, lcap = 0, qcap = 0, ccap = 0;
  int64_t arena_cap = 0;
  int32_t threads = 0;  // 0 = auto per launch
  int32_t simple = 0;   // KD_SEARCH_SIMPLE
  int32_t chunk_frames = 128;
  uint32_t region_f = 16, region_min = 16384;  // per-frame table region (kd::Params)
  size_t device_bytes = 0;
  size_t l2_window_bytes = 0, l2_persist_bytes = 0;

  kd::LaneState *lanes = nullptr;
  double *a_cost = nullptr;
  unsigned long long *a_link = nullptr;
  int32_t *a_state = nullptr;
  kd::Entry *table = nullptr;
  uint32_t *bitmap = nullptr;  // hcap / 32 words per lane: table entries in use
  uint32_t *list = nullptr;
  uint2 *queue = nullptr;
  uint4 *cand = nullptr;
  uint4 *front = nullptr;
  kd::AdvanceItem *d_items = nullptr;  // synchronous helpers (init, best path, ...)
  long long *d_out_off = nullptr;
  int32_t *d_flags = nullptr;          // ReachedFinal results, max_lanes

  int32_t *d_path = nullptr;  // best paths: [ilabel | olabel | graph | acoustic], path_cap words each
  int32_t *h_path = nullptr;  // pinned mirror
  int64_t path_cap = 0;

  // pinned host mirrors
  kd::AdvanceItem *h_items = nullptr;
  kd::LaneState *h_lanes = nullptr;
  long long *h_out_off = nullptr;
  int32_t *h_flags = nullptr;

  std::vector<int32_t> frames;  // host mirror of num_frames_decoded_, -1 = not initialised
  std::vector<int32_t> status;
  std::vector<uint8_t> bp_valid;   // the lane's best path is selected for its current tokens
  std::vector<uint8_t> use_final;  // per lane: use_final_probs of its last best-path request
  std::vector<int8_t> rf_cache;    // per lane: ReachedFinal of its current tokens, -1 = unknown
  std::vector<int8_t> lane_slot;   // per lane: slot of the call in flight that owns it, -1 = none

  kd_slot slots[kNumSlots];
  int64_t next_ticket = 0;
  // timing span over several (overlapping) launches: begin = before the first launch after
  // kd_decoder_span_begin, end = behind the most recently enqueued launch
  cudaEvent_t ev_span_begin = nullptr, ev_span_end = nullptr;
  bool span_armed = false, span_open = false;
  int32_t span_launches = 0;
  cudaStream_t stream = nullptr;   // synchronous helpers
  float last_kernel_ms = 0.f;
  int32_t last_launches = 0;
  int32_t last_blocks_per_sm = 0;
};

namespace {

int CheckOptions(const kd_options *o) {
  // faster-decoder.cc:24-28
  if (!o) return Fail(KD_ERR_INVALID, "options is null");
  if (!(o->hash_ratio >= 1.0f))
    return Fail(KD_ERR_INVALID, "Check failed!\nx: config_.hash_ratio >= 1.0");
  if (!(o->max_active > 1))
    return Fail(KD_ERR_INVALID, "Check failed!\nx: config_.max_active > 1");
  if (!(o->min_active >= 0 && o->min_active < o->max_active))
    return Fail(KD_ERR_INVALID,
                "Check failed!\nx: config_.min_active >= 0 && config_.min_active < "
                "config_.max_active");
  return KD_OK;
}

kd::Params MakeParams(const kd_decoder *d) {
  kd::Params P;
  memset(&P, 0, sizeof(P));
  P.st = d->g->st;
  P.labtab = d->g->labtab;
  P.lab_stride = d->g->max_ilabel;
  P.simple = d->simple;
  P.e_iw = d->g->e_iw;
  P.e_no = d->g->e_no;
  P.n_arc = d->g->n_arc;
  P.fin = d->g->fin;
  P.start = d->g->start;
  P.beam = d->opts.beam;
  // SimpleDecoder has no max_active / min_active: GetCutoff degenerates to best + beam
  P.max_active = d->simple ? 0x7FFFFFFF : d->opts.max_active;
  P.min_active = d->simple ? 0 : d->opts.min_active;
  P.beam_delta = d->opts.beam_delta;
  P.lanes = d->lanes;
  P.items = d->d_items;
  P.a_cost = d->a_cost;
  P.a_link = d->a_link;
  P.a_state = d->a_state;
  P.arena_cap = d->arena_cap;
  P.table = d->table;
  P.bitmap = d->bitmap;
  P.list = d->list;
  P.queue = d->queue;
  P.cand = d->cand;
  P.front = d->front;
  P.ccap = d->ccap;
  P.hcap = d->hcap;
  P.hmask = d->hcap - 1;
  P.lcap = d->lcap;
  P.qcap = d->qcap;
  int lg = 0;
  while ((1u << lg) < d->hcap) ++lg;
  P.hshift = 34 - lg;  // (lg - 2) scattered group bits, 2 in-group bits
  P.region_f = d->region_f;
  P.region_min = d->region_min;
  return P;
}

// The widest lane that still lets every lane of the call be resident at once (one wave):
// 512 threads x 1 lane per SM, 384 x 2, 256 x 4, 192 x 5, 160 x 7.
int PickThreads(const kd_decoder *d, int n_items) {
  if (d->threads > 0) return d->threads;
  if (n_items <= d->num_sms) return 512;
  if (n_items <= 2 * d->num_sms) return 384;
  if (n_items <= 4 * d->num_sms) return 256;
  if (n_items <= 5 * d->num_sms) return 192;
  return 160;
}

template <int THREADS, int MIN_BLOCKS, bool ROW_SMEM>
int LaunchAdvanceR(kd_decoder *d, kd::Params P, int n_items, cudaStream_t s) {
  size_t smem = kd::advance_smem_fixed<THREADS>() + 16;
  if (P.row_in_smem) smem += static_cast<size_t>(P.cols) * (sizeof(float) + sizeof(uint16_t));
  if (smem > 32 * 1024)  // static shared memory counts against the 48 KB default limit too
    KD_CUDA(cudaFuncSetAttribute(kd::kd_advance_kernel<THREADS, MIN_BLOCKS, ROW_SMEM>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)));
  int per_sm = 1;
  KD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
      &per_sm, kd::kd_advance_kernel<THREADS, MIN_BLOCKS, ROW_SMEM>, THREADS, smem));
  if (per_sm < 1) per_sm = 1;
  int grid = std::min(n_items, per_sm * d->num_sms);
  kd::kd_advance_kernel<THREADS, MIN_BLOCKS, ROW_SMEM><<<grid, THREADS, smem, s>>>(P);
  KD_CUDA(cudaGetLastError());
  d->last_blocks_per_sm = per_sm;
  return KD_OK;
}

template <int THREADS, int MIN_BLOCKS>
int LaunchAdvanceT(kd_decoder *d, const kd::Params &P, int n_items, cudaStream_t s) {
  return P.row_in_smem ? LaunchAdvanceR<THREADS, MIN_BLOCKS, true>(d, P, n_items, s)
                       : LaunchAdvanceR<THREADS, MIN_BLOCKS, false>(d, P, n_items, s);
}

// SimpleDecoder search: one instantiation (256 threads per lane), the mode is an API
// completeness feature, not a tuned path.
int LaunchAdvanceSimple(kd_decoder *d, kd::Params P, int n_items, cudaStream_t s) {
  constexpr int THREADS = 256;
  size_t smem = kd::advance_smem_fixed<THREADS>() + 16;
  if (P.row_in_smem) smem += static_cast<size_t>(P.cols) * (sizeof(float) + sizeof(uint16_t));
  auto launch = [&](auto kernel) -> int {
    if (smem > 32 * 1024)
      KD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
    int per_sm = 1;
    KD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    const int grid = std::min(n_items, per_sm * d->num_sms);
    kernel<<<grid, THREADS, smem, s>>>(P);
    KD_CUDA(cudaGetLastError());
    d->last_blocks_per_sm = per_sm;
    return KD_OK;
  };
  return P.row_in_smem ? launch(kd::kd_advance_kernel<THREADS, 3, true, true>)
                       : launch(kd::kd_advance_kernel<THREADS, 3, false, true>);
}

int LaunchAdvance(kd_decoder *d, const kd::Params &P, int n_items, int threads,
                  cudaStream_t s) {
  if (d->simple) return LaunchAdvanceSimple(d, P, n_items, s);
  // (second template argument: lanes per SM the register budget is cut for -- the measured
  // best per thread count: profiles/r2_sweeps.txt)
#ifdef KD_ONLY_160
  // (quick A/B builds: one instantiation, -DKD_AB_THREADS=t -DKD_AB_BLOCKS=b to pick another)
#ifndef KD_AB_THREADS
#define KD_AB_THREADS 160
#define KD_AB_BLOCKS 7
#endif
  if (threads != KD_AB_THREADS && threads != 160)
    return Fail(KD_ERR_INVALID, "this build has one lane width only");
  return LaunchAdvanceT<KD_AB_THREADS, KD_AB_BLOCKS>(d, P, n_items, s);
#else
  switch (threads) {
    case 128:
      return LaunchAdvanceT<128, 8>(d, P, n_items, s);
    case 160:
      return LaunchAdvanceT<160, 7>(d, P, n_items, s);
    case 192:
      return LaunchAdvanceT<192, 5>(d, P, n_items, s);
    case 224:
      return LaunchAdvanceT<224, 5>(d, P, n_items, s);
    case 256:
      return LaunchAdvanceT<256, 4>(d, P, n_items, s);
    case 384:
      return LaunchAdvanceT<384, 2>(d, P, n_items, s);
    case 512:
      return LaunchAdvanceT<512, 1>(d, P, n_items, s);
    default:
      return Fail(KD_ERR_INVALID,
                  "threads_per_lane must be 128, 160, 192, 224, 256, 384 or 512");
  }
#endif
}

const char *StatusText(int st) {
  if (st & kd::kStatusHashOverflow)
    return "recombination table overflow (raise kd_decoder_config.hash_capacity)";
  if (st & kd::kStatusArenaOverflow)
    return "backpointer arena overflow (raise kd_decoder_config.arena_records)";
  if (st & kd::kStatusQueueOverflow)
    return "epsilon worklist overflow (raise kd_decoder_config.hash_capacity)";
  if (st & kd::kStatusInputStall) return "streamed log-probs did not arrive (copy stalled)";
  if (st & kd::kStatusCandOverflow)
    return "candidate buffer overflow (raise kd_decoder_config.hash_capacity)";
  return "unknown device status";
}

int CheckLanes(const kd_decoder *d, int32_t n, const int32_t *lanes) {
  if (!d) return Fail(KD_ERR_INVALID, "decoder is null");
  if (n < 0 || (n > 0 && !lanes)) return Fail(KD_ERR_INVALID, "bad lane list");
  if (n > d->max_lanes) return Fail(KD_ERR_INVALID, "more lanes than max_lanes");
  for (int32_t i = 0; i < n; ++i)
    if (lanes[i] < 0 || lanes[i] >= d->max_lanes)
      return Fail(KD_ERR_INVALID, "lane id out of range");
  // a lane may appear once per call (two CTAs must never own the same lane)
  std::vector<char> seen(static_cast<size_t>(d->max_lanes), 0);
  for (int32_t i = 0; i < n; ++i) {
    if (seen[lanes[i]]) return Fail(KD_ERR_INVALID, "duplicate lane id in one call");
    seen[lanes[i]] = 1;
  }
  return KD_OK;
}

// Pulls the device lane states of `lanes` into the pinned mirror.
int FetchLaneStates(kd_decoder *d, int32_t n, const int32_t *lanes, cudaStream_t s, bool sync = true) {
  if (n == 0) return KD_OK;
  int32_t lo = lanes[0], hi = lanes[0];
  for (int32_t i = 1; i < n; ++i) {
    lo = std::min(lo, lanes[i]);
    hi = std::max(hi, lanes[i]);
  }
  KD_CUDA(cudaMemcpyAsync(d->h_lanes + lo, d->lanes + lo,
                          sizeof(kd::LaneState) * static_cast<size_t>(hi - lo + 1),
                          cudaMemcpyDeviceToHost, s));
  if (sync) KD_CUDA(cudaStreamSynchronize(s));
  return KD_OK;
}

}  // namespace

extern "C" {

const char *kd_last_error(void) { return g_error.c_str(); }

int kd_device_count(int *count) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  if (count) *count = n;
  return KD_OK;
}

int kd_graph_create(int device, int32_t num_states, int32_t start, const int64_t *row_offsets,
                    const int32_t *ilabel, const int32_t *olabel, const float *weight,
                    const int32_t *nextstate, const float *final_weight, kd_graph **out) {
  if (!out) return Fail(KD_ERR_INVALID, "out is null");
  *out = nullptr;
  if (num_states <= 0 || !row_offsets || !final_weight)
    return Fail(KD_ERR_INVALID, "empty graph");
  // faster-decoder.cc:47: InitDecoding asserts Start() != kNoStateId
  if (start < 0 || start >= num_states)
    return Fail(KD_ERR_INVALID, "Check failed!\nx: start_state != fst::kNoStateId");
  const int64_t E = row_offsets[num_states];
  if (row_offsets[0] != 0 || E < 0) return Fail(KD_ERR_INVALID, "bad row_offsets");
  if (E > 0 && (!ilabel || !olabel || !weight || !nextstate))
    return Fail(KD_ERR_INVALID, "null arc arrays");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return Fail(KD_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
  }
  if (device < 0 || device >= ndev) return Fail(KD_ERR_INVALID, "device index out of range");
  KD_CUDA(cudaSetDevice(device));

  // st[2s] = {emit_begin, emit_count, eps_begin, eps_count}
  // st[2s+1] = {label-table row or -1, bits of the smallest emitting weight, 0, 0}
  std::vector<int4> st(2 * static_cast<size_t>(num_states));
  int64_t n_emit = 0, n_eps = 0;
  int32_t max_il = 0;
  for (int32_t s = 0; s < num_states; ++s) {
    if (row_offsets[s + 1] < row_offsets[s]) return Fail(KD_ERR_INVALID, "bad row_offsets");
    int64_t ne = 0, nn = 0;
    float wmin = std::numeric_limits<float>::infinity();
    for (int64_t a = row_offsets[s]; a < row_offsets[s + 1]; ++a) {
      if (ilabel[a] < 0) return Fail(KD_ERR_INVALID, "negative ilabel");
      if (nextstate[a] < 0 || nextstate[a] >= num_states)
        return Fail(KD_ERR_INVALID, "arc nextstate out of range");
      if (ilabel[a] != 0) {
        ++ne;
        max_il = std::max(max_il, ilabel[a]);
        if (weight[a] < wmin) wmin = weight[a];  // NaN weights never lower it: no label table then
        if (!(weight[a] == weight[a])) wmin = -std::numeric_limits<float>::infinity();
      } else {
        ++nn;
      }
    }
    int wmin_bits;
    memcpy(&wmin_bits, &wmin, 4);
    st[2 * static_cast<size_t>(s)] = make_int4(static_cast<int>(n_emit), static_cast<int>(ne),
                                               static_cast<int>(n_eps), static_cast<int>(nn));
    st[2 * static_cast<size_t>(s) + 1] = make_int4(-1, wmin_bits, 0, 0);
    n_emit += ne;
    n_eps += nn;
  }
  // Label tables: for states with many emitting arcs and distinct ilabels, a row
  // ilabel-1 -> (weight, offset) of that arc within the state: one 8-byte load gives a
  // looked-up arc's weight, the arc record itself is only read for candidates.  A token whose slack
  // (cutoff - cost - smallest weight) only admits the few best labels of the frame
  // looks those labels up instead of scanning all its arcs.
  std::vector<int2> labtab;
  int64_t n_tab = 0;
  if (max_il > 0 && max_il <= 65535 && getenv("KD_B200_NO_LABEL_TABLES") == nullptr) {
    std::vector<std::pair<int32_t, int32_t>> cands;  // (emit count, state)
    for (int32_t s = 0; s < num_states; ++s) {
      const int ne = st[2 * static_cast<size_t>(s)].y;
      if (ne >= kd::kLabelTableMinDegree) cands.emplace_back(ne, s);
    }
    std::sort(cands.begin(), cands.end(), [](const std::pair<int32_t, int32_t> &a,
                                             const std::pair<int32_t, int32_t> &b) {
      return a.first != b.first ? a.first > b.first : a.second < b.second;
    });
    const size_t row_bytes = static_cast<size_t>(max_il) * sizeof(int2);
    const size_t budget = std::max<size_t>(256u << 20, static_cast<size_t>(n_emit) * 16);
    std::vector<int32_t> stamp(static_cast<size_t>(max_il) + 1, -1);
    for (const auto &c : cands) {
      if ((static_cast<size_t>(n_tab) + 1) * row_bytes > budget) break;
      const int32_t s = c.second;
      bool distinct = true;
      for (int64_t a = row_offsets[s]; a < row_offsets[s + 1] && distinct; ++a) {
        if (ilabel[a] == 0) continue;
        if (stamp[ilabel[a]] == s) distinct = false;
        stamp[ilabel[a]] = s;
      }
      if (!distinct) continue;
      labtab.resize((static_cast<size_t>(n_tab) + 1) * max_il, make_int2(0, -1));
      int2 *row = labtab.data() + static_cast<size_t>(n_tab) * max_il;
      int32_t off = 0;
      for (int64_t a = row_offsets[s]; a < row_offsets[s + 1]; ++a) {
        if (ilabel[a] == 0) continue;
        int wbits;
        memcpy(&wbits, &weight[a], 4);
        row[ilabel[a] - 1] = make_int2(wbits, off++);
      }
      st[2 * static_cast<size_t>(s) + 1].x = static_cast<int>(n_tab);
      ++n_tab;
    }
  }
  if (n_emit >= 0x7FFFFFFFll || n_eps >= 0x7FFFFFFFll)
    return Fail(KD_ERR_INVALID, "graph too large (2^31 arcs)");
  // (nextstate | kEpsFlag) marks destinations that have epsilon arcs of their own
  std::vector<int2> eiw(static_cast<size_t>(n_emit)), eno(static_cast<size_t>(n_emit));
  std::vector<int4> na(static_cast<size_t>(n_eps));
  {
    int64_t ie = 0, in = 0;
    for (int64_t a = 0; a < E; ++a) {
      int wbits;
      memcpy(&wbits, &weight[a], 4);
      const int32_t ns = nextstate[a];
      const int ns_word = st[2 * static_cast<size_t>(ns)].w > 0
                              ? static_cast<int>(static_cast<uint32_t>(ns) | kd::kEpsFlag)
                              : ns;
      if (ilabel[a] != 0) {
        eiw[ie] = make_int2(ilabel[a], wbits);
        eno[ie] = make_int2(ns_word, olabel[a]);
        ++ie;
      } else {
        na[in++] = make_int4(olabel[a], wbits, ns_word, 0);
      }
    }
  }
  auto *g = new kd_graph;
  g->device = device;
  g->num_states = num_states;
  g->num_arcs = E;
  g->num_emit = n_emit;
  g->num_eps = n_eps;
  g->start = start;
  g->max_ilabel = max_il;
  int rc;
  auto round256 = [](size_t b) { return (b + 255) & ~static_cast<size_t>(255); };
  const size_t b_iw = round256(eiw.size() * sizeof(int2)), b_st = round256(st.size() * sizeof(int4)),
               b_no = round256(eno.size() * sizeof(int2)), b_na = round256(na.size() * sizeof(int4)),
               b_fin = round256(static_cast<size_t>(num_states) * sizeof(float)),
               b_tab = round256(labtab.size() * sizeof(int2));
  g->blob_bytes = b_iw + b_st + b_no + b_na + b_fin + b_tab;
  g->num_label_tables = n_tab;
  if ((rc = DevAlloc(&g->blob, g->blob_bytes))) {
    kd_graph_destroy(g);
    return rc;
  }
  g->e_iw = reinterpret_cast<int2 *>(g->blob);
  g->st = reinterpret_cast<int4 *>(g->blob + b_iw);
  g->e_no = reinterpret_cast<int2 *>(g->blob + b_iw + b_st);
  g->n_arc = reinterpret_cast<int4 *>(g->blob + b_iw + b_st + b_no);
  g->fin = reinterpret_cast<float *>(g->blob + b_iw + b_st + b_no + b_na);
  g->labtab = reinterpret_cast<int2 *>(g->blob + b_iw + b_st + b_no + b_na + b_fin);
  auto upload = [&](void *dst, const void *src, size_t bytes) -> int {
    if (bytes == 0) return KD_OK;
    KD_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return KD_OK;
  };
  if ((rc = upload(g->labtab, labtab.data(), labtab.size() * sizeof(int2))) ||
      (rc = upload(g->st, st.data(), st.size() * sizeof(int4))) ||
      (rc = upload(g->e_iw, eiw.data(), eiw.size() * sizeof(int2))) ||
      (rc = upload(g->e_no, eno.data(), eno.size() * sizeof(int2))) ||
      (rc = upload(g->n_arc, na.data(), na.size() * sizeof(int4))) ||
      (rc = upload(g->fin, final_weight, sizeof(float) * num_states))) {
    const std::string keep = g_error;
    kd_graph_destroy(g);
    g_error = keep;
    return rc;
  }
  *out = g;
  return KD_OK;
}

int kd_graph_destroy(kd_graph *g) {
  if (!g) return KD_OK;
  cudaSetDevice(g->device);
  cudaFree(g->blob);
  delete g;
  return KD_OK;
}

int kd_graph_info(const kd_graph *g, int64_t info[5]) {
  if (!g || !info) return Fail(KD_ERR_INVALID, "null argument");
  info[0] = g->num_states;
  info[1] = g->num_arcs;
  info[2] = g->num_eps;
  info[3] = g->max_ilabel;
  info[4] = g->device;
  return KD_OK;
}

int kd_decoder_create(kd_graph *g, const kd_options *opts, const kd_decoder_config *cfg,
                      kd_decoder **out) {
  if (!out) return Fail(KD_ERR_INVALID, "out is null");
  *out = nullptr;
  if (!g) return Fail(KD_ERR_INVALID, "graph is null");
  int rc = CheckOptions(opts);
  if (rc) return rc;
  KD_CUDA(cudaSetDevice(g->device));
  kd_decoder_config c;
  memset(&c, 0, sizeof(c));
  if (cfg) c = *cfg;
  if (c.threads_per_lane != 0 && c.threads_per_lane != 128 && c.threads_per_lane != 160 &&
      c.threads_per_lane != 192 && c.threads_per_lane != 224 && c.threads_per_lane != 256 &&
      c.threads_per_lane != 384 && c.threads_per_lane != 512)
    return Fail(KD_ERR_INVALID, "threads_per_lane must be 0, 128, 160, 192, 224, 256, 384 or 512");
  if (c.search != KD_SEARCH_FASTER && c.search != KD_SEARCH_SIMPLE)
    return Fail(KD_ERR_INVALID,
                "kd_decoder_config.search must be KD_SEARCH_FASTER or KD_SEARCH_SIMPLE");
  auto *d = new kd_decoder;
  // (every early return below goes through kd_decoder_destroy: nothing leaks)
  auto fail = [&](int code) {
    const std::string keep = g_error;
    kd_decoder_destroy(d);
    g_error = keep;
    return code;
  };
#define KD_CUDA_D(expr)                                                                      \
  do {                                                                                       \
    cudaError_t kd_e_ = (expr);                                                              \
    if (kd_e_ != cudaSuccess)                                                                \
      return fail(Fail(KD_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(kd_e_))); \
  } while (0)
  d->g = g;
  d->opts = *opts;
  d->device = g->device;
  d->max_lanes = c.max_lanes > 0 ? c.max_lanes : 1;
  cudaDeviceProp prop;
  KD_CUDA_D(cudaGetDeviceProperties(&prop, g->device));
  d->num_sms = prop.multiProcessorCount;
  d->mem_pitch = prop.memPitch;
  {
    const char *e = getenv("CUDA_LAUNCH_BLOCKING");
    d->launch_blocking = e != nullptr && e[0] != '\0' && e[0] != '0';
  }
  // Default table size: the largest power of two that keeps all lanes' tables (and their
  // slot lists, worklists, candidate buffers: 54 bytes per entry) within ~15% of the free
  // device memory, between 2^15 and 2^22 entries.  A frame may hold capacity / 2 tokens.
  uint32_t hcap;
  if (c.hash_capacity > 0) {
    hcap = static_cast<uint32_t>(c.hash_capacity);
  } else {
    size_t free_b = 0, total_b = 0;
    KD_CUDA_D(cudaMemGetInfo(&free_b, &total_b));
    const double per_lane = 0.15 * static_cast<double>(free_b) / d->max_lanes / 56.0;
    hcap = 1u << 15;
    while (hcap < (1u << 22) && 2.0 * hcap <= per_lane) hcap <<= 1;
  }
  uint32_t p2 = 64;
  while (p2 < hcap && p2 < (1u << 30)) p2 <<= 1;
  d->hcap = p2;
  d->lcap = p2 / 2;
  d->qcap = p2;
  d->ccap = p2 / 4;
  d->threads = c.threads_per_lane > 0 ? c.threads_per_lane : 0;
  d->simple = c.search == KD_SEARCH_SIMPLE ? 1 : 0;
  d->chunk_frames = c.chunk_frames > 0 ? c.chunk_frames : 128;
  // KD_B200_TABLE_REGION="F,MIN": entries per candidate and smallest size of a frame's table
  // region (measured default 16,16384; "0" = always the whole table; tiny values make most
  // frames start over, which is how the tests exercise that path)
  if (const char *e = getenv("KD_B200_TABLE_REGION")) {
    unsigned f = 0, m = 0;
    const int got = sscanf(e, "%u%*1[,:]%u", &f, &m);  // "F,MIN" or "F:MIN"
    if (got >= 1) d->region_f = f;
    if (got >= 2) {
      uint32_t p = 64;
      while (p < m && p < (1u << 30)) p <<= 1;
      d->region_min = p;
    }
  }

  const size_t L = static_cast<size_t>(d->max_lanes);
  const size_t table_bytes_per_lane =
      static_cast<size_t>(d->hcap) * sizeof(kd::Entry) + static_cast<size_t>(d->lcap) * 4 +
      static_cast<size_t>(d->qcap) * 16 + static_cast<size_t>(d->ccap) * 16;
  if (c.arena_records > 0) {
    d->arena_cap = c.arena_records;
  } else {
    size_t free_b = 0, total_b = 0;
    KD_CUDA_D(cudaMemGetInfo(&free_b, &total_b));
    // (0.7 of the free memory for tables + arena: the rest is left to the caller's matrices, the
    // staging buffers of host input, and whatever else lives on the device)
    double budget = 0.7 * static_cast<double>(free_b) - static_cast<double>(table_bytes_per_lane * L);
    long long per_lane = static_cast<long long>(budget / (20.0 * static_cast<double>(L)));
    d->arena_cap = std::max<long long>(1 << 16, std::min<long long>(per_lane, 1ll << 25));
  }
  if (d->arena_cap > 0xFFFFFFF0ll) d->arena_cap = 0xFFFFFFF0ll;

  const size_t A = static_cast<size_t>(d->arena_cap);
  if ((rc = DevAlloc(&d->lanes, L)) || (rc = DevAlloc(&d->a_cost, L * A)) ||
      (rc = DevAlloc(&d->a_link, L * A)) || (rc = DevAlloc(&d->a_state, L * A)) ||
      (rc = DevAlloc(&d->table, L * d->hcap)) || (rc = DevAlloc(&d->bitmap, L * (d->hcap / 32))) ||
      (rc = DevAlloc(&d->list, L * d->lcap)) ||
      (rc = DevAlloc(&d->queue, L * 2 * d->qcap)) || (rc = DevAlloc(&d->cand, L * d->ccap)) ||
      (rc = DevAlloc(&d->front, L * kd::kFrontCap)) || (rc = DevAlloc(&d->d_items, L)) ||
      (rc = DevAlloc(&d->d_out_off, L)) || (rc = DevAlloc(&d->d_flags, L)))
    return fail(rc);
  d->device_bytes = L * (sizeof(kd::LaneState) + A * 20 + table_bytes_per_lane);
  KD_CUDA_D(cudaMemset(d->lanes, 0, L * sizeof(kd::LaneState)));
  KD_CUDA_D(cudaMemset(d->table, 0xFF, L * d->hcap * sizeof(kd::Entry)));
  KD_CUDA_D(cudaMemset(d->bitmap, 0, L * (d->hcap / 32) * sizeof(uint32_t)));
  KD_CUDA_D(cudaMallocHost(reinterpret_cast<void **>(&d->h_items), L * sizeof(kd::AdvanceItem)));
  KD_CUDA_D(cudaMallocHost(reinterpret_cast<void **>(&d->h_lanes), L * sizeof(kd::LaneState)));
  KD_CUDA_D(cudaMallocHost(reinterpret_cast<void **>(&d->h_out_off), L * sizeof(long long)));
  KD_CUDA_D(cudaMallocHost(reinterpret_cast<void **>(&d->h_flags), L * sizeof(int32_t)));
  KD_CUDA_D(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
  KD_CUDA_D(cudaEventCreate(&d->ev_span_begin));
  KD_CUDA_D(cudaEventCreate(&d->ev_span_end));
  for (int i = 0; i < kNumSlots; ++i) {
    kd_slot &s = d->slots[i];
    KD_CUDA_D(cudaStreamCreateWithFlags(&s.sc, cudaStreamNonBlocking));
    KD_CUDA_D(cudaStreamCreateWithFlags(&s.sx, cudaStreamNonBlocking));
    KD_CUDA_D(cudaEventCreateWithFlags(&s.ev_gate, cudaEventDisableTiming));
    KD_CUDA_D(cudaEventCreateWithFlags(&s.ev_dep, cudaEventDisableTiming));
    KD_CUDA_D(cudaEventCreate(&s.ev_begin));
    KD_CUDA_D(cudaEventCreate(&s.ev_end));
    if ((rc = DevAlloc(&s.d_items, L)) || (rc = DevAlloc(&s.d_words, static_cast<size_t>(4))) ||
        (rc = DevAlloc(&s.d_out_off, L)))
      return fail(rc);
    KD_CUDA_D(cudaMallocHost(reinterpret_cast<void **>(&s.h_items), L * sizeof(kd::AdvanceItem)));
    KD_CUDA_D(cudaMallocHost(reinterpret_cast<void **>(&s.h_out_off), L * sizeof(long long)));
  }
  // Optional (KD_B200_L2_PIN=1): an L2 persistence window over the graph.  Measured
  // on the 5M-arc HLG with 1024 lanes it is a loss (253 ms vs 190 ms per 1000
  // frames): the per-lane table entries are as hot as the arcs and the carve-out
  // takes L2 away from them.  Left off by default.
  if (getenv("KD_B200_L2_PIN") != nullptr && prop.persistingL2CacheMaxSize > 0 &&
      prop.accessPolicyMaxWindowSize > 0) {
    const size_t persist = static_cast<size_t>(prop.persistingL2CacheMaxSize);
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist) == cudaSuccess) {
      const size_t win = std::min(g->blob_bytes, static_cast<size_t>(prop.accessPolicyMaxWindowSize));
      cudaStreamAttrValue attr;
      memset(&attr, 0, sizeof(attr));
      attr.accessPolicyWindow.base_ptr = g->blob;
      attr.accessPolicyWindow.num_bytes = win;
      attr.accessPolicyWindow.hitRatio =
          win <= persist ? 1.0f : static_cast<float>(static_cast<double>(persist) / win);
      attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      for (int i = 0; i < kNumSlots; ++i)
        cudaStreamSetAttribute(d->slots[i].sc, cudaStreamAttributeAccessPolicyWindow, &attr);
      d->l2_window_bytes = win;
      d->l2_persist_bytes = persist;
    }
    cudaGetLastError();
  }
  d->frames.assign(L, -1);
  d->status.assign(L, 0);
  d->bp_valid.assign(L, 0);
  d->use_final.assign(L, 1);
  d->rf_cache.assign(L, -1);
  d->lane_slot.assign(L, -1);
  KD_CUDA_D(cudaDeviceSynchronize());
#undef KD_CUDA_D
  *out = d;
  return KD_OK;
}

int kd_decoder_destroy(kd_decoder *d) {
  if (!d) return KD_OK;
  cudaSetDevice(d->device);
  cudaDeviceSynchronize();
  cudaFree(d->lanes);
  cudaFree(d->a_cost);
  cudaFree(d->a_link);
  cudaFree(d->a_state);
  cudaFree(d->table);
  cudaFree(d->bitmap);
  cudaFree(d->list);
  cudaFree(d->queue);
  cudaFree(d->cand);
  cudaFree(d->front);
  cudaFree(d->d_items);
  cudaFree(d->d_out_off);
  cudaFree(d->d_flags);
  cudaFree(d->d_path);
  if (d->h_path) cudaFreeHost(d->h_path);
  if (d->h_items) cudaFreeHost(d->h_items);
  if (d->h_lanes) cudaFreeHost(d->h_lanes);
  if (d->h_out_off) cudaFreeHost(d->h_out_off);
  if (d->h_flags) cudaFreeHost(d->h_flags);
  for (int i = 0; i < kNumSlots; ++i) {
    kd_slot &s = d->slots[i];
    cudaFree(s.d_items);
    cudaFree(s.d_words);
    cudaFree(s.d_stage);
    cudaFree(s.d_out_off);
    if (s.h_items) cudaFreeHost(s.h_items);
    if (s.h_progress) cudaFreeHost(s.h_progress);
    if (s.h_res) cudaFreeHost(s.h_res);
    if (s.h_out_off) cudaFreeHost(s.h_out_off);
    if (s.sc) cudaStreamDestroy(s.sc);
    if (s.sx) cudaStreamDestroy(s.sx);
    if (s.ev_gate) cudaEventDestroy(s.ev_gate);
    if (s.ev_dep) cudaEventDestroy(s.ev_dep);
    if (s.ev_begin) cudaEventDestroy(s.ev_begin);
    if (s.ev_end) cudaEventDestroy(s.ev_end);
  }
  if (d->stream) cudaStreamDestroy(d->stream);
  if (d->ev_span_begin) cudaEventDestroy(d->ev_span_begin);
  if (d->ev_span_end) cudaEventDestroy(d->ev_span_end);
  cudaGetLastError();
  delete d;
  return KD_OK;
}

}  // extern "C"

namespace {

// ---------------------------------------------------------------- calls in flight

// Enqueues the search kernel of slot `s` (items already on the device) and, behind it, the
// copies that bring back the lane states and -- KD_ADVANCE_FINALIZE -- the parked paths.
int EnqueueSearch(kd_decoder *d, kd_slot &s) {
  const int32_t m = static_cast<int32_t>(s.lanes.size());
  KD_CUDA(cudaMemsetAsync(s.d_words, 0, 2 * sizeof(int32_t), s.sc));  // work counter, yield flag
  KD_CUDA(cudaEventRecord(s.ev_begin, s.sc));
  if (d->span_armed) {
    KD_CUDA(cudaEventRecord(d->ev_span_begin, s.sc));
    d->span_armed = false;
    d->span_open = true;
  }
  int rc = LaunchAdvance(d, s.P, m, s.threads, s.sc);
  if (rc) return rc;
  s.launches++;
  KD_CUDA(cudaEventRecord(s.ev_end, s.sc));
  if (d->span_open) {
    KD_CUDA(cudaEventRecord(d->ev_span_end, s.sc));
    d->span_launches++;
  }
  rc = FetchLaneStates(d, m, s.lanes.data(), s.sc, false);
  if (rc) return rc;
  if (s.flags & KD_ADVANCE_FINALIZE) {
    // the parked paths: lane i's four arrays are 4 * path_cap consecutive words at the head
    // of its worklist buffer.  Consecutive lane ids come back with one 2-D copy.
    const size_t row_bytes = static_cast<size_t>(4) * s.path_cap * sizeof(int32_t);
    const size_t lane_pitch = static_cast<size_t>(2) * d->qcap * sizeof(uint2);
    bool consecutive = lane_pitch <= d->mem_pitch;
    for (int32_t i = 1; i < m && consecutive; ++i)
      if (s.lanes[i] != s.lanes[i - 1] + 1) consecutive = false;
    if (consecutive) {
      KD_CUDA(cudaMemcpy2DAsync(s.h_res, row_bytes,
                                d->queue + static_cast<size_t>(s.lanes[0]) * 2 * d->qcap,
                                lane_pitch, row_bytes, m, cudaMemcpyDeviceToHost, s.sc));
    } else {
      for (int32_t i = 0; i < m; ++i)
        KD_CUDA(cudaMemcpyAsync(s.h_res + static_cast<size_t>(i) * 4 * s.path_cap,
                                d->queue + static_cast<size_t>(s.lanes[i]) * 2 * d->qcap,
                                row_bytes, cudaMemcpyDeviceToHost, s.sc));
    }
  }
  return KD_OK;
}

// Completes the call in flight in slot `si`: waits for its streams, launches again if lanes
// yielded (their rows had not been enqueued yet), refreshes the host mirrors of its lanes.
// The outcome is kept in the slot (rc, msg) and returned.
int CompleteSlot(kd_decoder *d, int si) {
  kd_slot &s = d->slots[si];
  if (!s.busy) return s.rc;
  const int32_t m = static_cast<int32_t>(s.lanes.size());
  auto finish = [&](int rc) {
    // whatever happened, nothing of this call is left running
    cudaStreamSynchronize(s.sx);
    cudaStreamSynchronize(s.sc);
    s.busy = false;
    s.rc = rc;
    s.msg = rc ? g_error : std::string();
    for (int32_t i = 0; i < m; ++i) d->lane_slot[s.lanes[i]] = -1;
    d->last_kernel_ms = s.kernel_ms;
    d->last_launches = s.launches;
    return rc;
  };
#define KD_CUDA_S(expr)                                                                        \
  do {                                                                                         \
    cudaError_t kd_e_ = (expr);                                                                \
    if (kd_e_ != cudaSuccess)                                                                  \
      return finish(Fail(KD_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(kd_e_))); \
  } while (0)
  KD_CUDA_S(cudaSetDevice(d->device));
  int stalled_rounds = 0;
  while (true) {
    KD_CUDA_S(cudaStreamSynchronize(s.sx));
    KD_CUDA_S(cudaStreamSynchronize(s.sc));
    float ms = 0.f;
    KD_CUDA_S(cudaEventElapsedTime(&ms, s.ev_begin, s.ev_end));
    s.kernel_ms += ms;
    // lanes that yielded: their rows were not there (launches serialised, or pageable host
    // memory staged by the host thread).  Every copy is enqueued by now, so one more launch
    // finds all rows; InitDecoding has been done by the first one.
    int32_t behind = 0, progressed = 0;
    for (int32_t i = 0; i < m; ++i) {
      const kd::LaneState &L = d->h_lanes[s.lanes[i]];
      if (L.status == 0 && L.frames_decoded < s.targets[i]) ++behind;
      if (L.frames_decoded > s.h_items[i].offset) ++progressed;
    }
    if (behind == 0) break;
    if (!s.host_input || (stalled_rounds > 0 && progressed == 0) || stalled_rounds >= 4) {
      for (int32_t i = 0; i < m; ++i) {
        const int32_t lane = s.lanes[i];
        d->frames[lane] = d->h_lanes[lane].frames_decoded;
        d->status[lane] = d->h_lanes[lane].status | kd::kStatusInputStall;
      }
      return finish(Fail(KD_ERR_OVERFLOW, std::string("AdvanceDecoding: ") +
                                              StatusText(kd::kStatusInputStall)));
    }
    ++stalled_rounds;
    for (int32_t i = 0; i < m; ++i) s.h_items[i].flags &= ~kd::kItemInit;
    KD_CUDA_S(cudaMemcpyAsync(s.d_items, s.h_items, sizeof(kd::AdvanceItem) * m,
                              cudaMemcpyHostToDevice, s.sc));
    int rc = EnqueueSearch(d, s);
    if (rc) return finish(rc);
  }
  int bad = 0;
  for (int32_t i = 0; i < m; ++i) {
    const int32_t lane = s.lanes[i];
    const kd::LaneState &L = d->h_lanes[lane];
    d->frames[lane] = L.frames_decoded;
    d->status[lane] = L.status;
    d->rf_cache[lane] = -1;
    d->bp_valid[lane] = 0;  // a path parked in the candidate buffer by an earlier call is gone
    if (L.status != 0) {
      if (bad == 0) bad = L.status;
    } else if (s.flags & KD_ADVANCE_FINALIZE) {
      d->bp_valid[lane] = 1;
      d->rf_cache[lane] = static_cast<int8_t>(L.bp_final != 0);
    }
  }
  if (bad) return finish(Fail(KD_ERR_OVERFLOW, std::string("AdvanceDecoding: ") + StatusText(bad)));
  if (s.flags & KD_ADVANCE_FINALIZE) {
    // word offsets of every lane's arrays in h_res; paths that could not be parked (longer
    // than path_cap) are written by kd_best_fill_kernel into a spill area behind the block
    s.res_off.assign(static_cast<size_t>(4) * m, 0);
    std::vector<int32_t> spill;
    int64_t spill_words = 0;
    const int64_t block = static_cast<int64_t>(4) * s.path_cap;
    for (int32_t i = 0; i < m; ++i) {
      const kd::LaneState &L = d->h_lanes[s.lanes[i]];
      for (int k = 0; k < 4; ++k) s.res_off[4 * i + k] = i * block + k * s.path_cap;
      if (L.bp_ok && !L.bp_parked) {
        spill.push_back(i);
        spill_words += 4 * L.bp_len;
      }
    }
    if (!spill.empty()) {
      const size_t need = static_cast<size_t>(m) * block + static_cast<size_t>(spill_words);
      if (need > s.res_words) {
        int32_t *bigger = nullptr;
        KD_CUDA_S(cudaMallocHost(reinterpret_cast<void **>(&bigger), need * sizeof(int32_t)));
        memcpy(bigger, s.h_res, static_cast<size_t>(m) * block * sizeof(int32_t));
        cudaFreeHost(s.h_res);
        s.h_res = bigger;
        s.res_words = need;
      }
      int32_t *d_spill = nullptr;
      KD_CUDA_S(cudaMalloc(reinterpret_cast<void **>(&d_spill), spill_words * sizeof(int32_t)));
      int64_t pos = 0;
      const int32_t ns = static_cast<int32_t>(spill.size());
      for (int32_t k = 0; k < ns; ++k) {
        const int32_t i = spill[k];
        const int64_t len = d->h_lanes[s.lanes[i]].bp_len;
        s.h_items[k] = s.h_items[i];  // (items are spent: reused as the lane list)
        s.h_items[k].lane = s.lanes[i];
        s.h_out_off[k] = 0;
        // one lane at a time: four arrays of `len` words each
        for (int a = 0; a < 4; ++a) s.res_off[4 * i + a] = m * block + pos + a * len;
        pos += 4 * len;
      }
      kd::Params P = s.P;
      pos = 0;
      for (int32_t k = 0; k < ns; ++k) {
        const int64_t len = d->h_lanes[s.h_items[k].lane].bp_len;
        KD_CUDA_S(cudaMemcpyAsync(s.d_items, s.h_items + k, sizeof(kd::AdvanceItem),
                                  cudaMemcpyHostToDevice, s.sc));
        KD_CUDA_S(cudaMemsetAsync(s.d_out_off, 0, sizeof(long long), s.sc));
        P.items = s.d_items;
        P.n_items = 1;
        int32_t *base = d_spill + pos;
        kd::kd_best_fill_kernel<<<1, 128, 0, s.sc>>>(P, s.d_out_off, base, base + len,
                                                     reinterpret_cast<float *>(base + 2 * len),
                                                     reinterpret_cast<float *>(base + 3 * len));
        KD_CUDA_S(cudaStreamSynchronize(s.sc));
        pos += 4 * len;
      }
      KD_CUDA_S(cudaMemcpy(s.h_res + static_cast<size_t>(m) * block, d_spill,
                           spill_words * sizeof(int32_t), cudaMemcpyDeviceToHost));
      cudaFree(d_spill);
    }
    s.res_valid = true;
  }
#undef KD_CUDA_S
  return finish(KD_OK);
}

// Deferred synchronisation: whoever touches a lane first completes the call that owns it.
int WaitLanes(kd_decoder *d, int32_t n, const int32_t *lanes) {
  int rc = KD_OK;
  for (int32_t i = 0; i < n; ++i) {
    const int si = d->lane_slot[lanes[i]];
    if (si >= 0) {
      const int r = CompleteSlot(d, si);
      if (r && !rc) rc = r;
    }
  }
  return rc;
}

int WaitAll(kd_decoder *d) {
  int rc = KD_OK;
  // oldest first
  for (int round = 0; round < kNumSlots; ++round) {
    int best = -1;
    for (int i = 0; i < kNumSlots; ++i)
      if (d->slots[i].busy && (best < 0 || d->slots[i].ticket < d->slots[best].ticket)) best = i;
    if (best < 0) break;
    const int r = CompleteSlot(d, best);
    if (r && !rc) rc = r;
  }
  return rc;
}

}  // namespace

extern "C" {

int kd_decoder_set_options(kd_decoder *d, const kd_options *opts) {
  if (!d) return Fail(KD_ERR_INVALID, "decoder is null");
  int rc = CheckOptions(opts);
  if (rc) return rc;
  WaitAll(d);  // calls in flight keep the options they were enqueued with
  d->opts = *opts;
  return KD_OK;
}

int kd_decoder_reset(kd_decoder *d) {
  if (!d) return Fail(KD_ERR_INVALID, "decoder is null");
  WaitAll(d);  // (the outcome of calls nobody waited for goes with them)
  for (int32_t l = 0; l < d->max_lanes; ++l) {
    d->frames[l] = -1;
    d->status[l] = 0;
    d->bp_valid[l] = 0;
    d->use_final[l] = 1;
    d->rf_cache[l] = -1;
    d->lane_slot[l] = -1;
  }
  for (int i = 0; i < kNumSlots; ++i) {
    d->slots[i].rc = KD_OK;
    d->slots[i].msg.clear();
  }
  return KD_OK;
}

int kd_decoder_init(kd_decoder *d, int32_t n, const int32_t *lanes) {
  int rc = CheckLanes(d, n, lanes);
  if (rc) return rc;
  if (n == 0) return KD_OK;
  KD_CUDA(cudaSetDevice(d->device));
  WaitLanes(d, n, lanes);  // (a failed call on these lanes is wiped out by this init)
  cudaStream_t s = d->stream;
  for (int32_t i = 0; i < n; ++i) {
    const int32_t lane = lanes[i];
    if (d->status[lane] != 0) {
      // a lane that overflowed may have left claimed table slots behind
      size_t off = static_cast<size_t>(lane) * d->hcap;
      KD_CUDA(cudaMemsetAsync(d->table + off, 0xFF, d->hcap * sizeof(kd::Entry), s));
      KD_CUDA(cudaMemsetAsync(d->bitmap + off / 32, 0, d->hcap / 8, s));
      d->status[lane] = 0;
    }
    memset(&d->h_items[i], 0, sizeof(kd::AdvanceItem));
    d->h_items[i].lane = lane;
  }
  KD_CUDA(cudaMemcpyAsync(d->d_items, d->h_items, sizeof(kd::AdvanceItem) * n,
                          cudaMemcpyHostToDevice, s));
  kd::Params P = MakeParams(d);
  P.items = d->d_items;
  P.n_items = n;
  kd::kd_init_kernel<256><<<n, 256, 0, s>>>(P);
  KD_CUDA(cudaGetLastError());
  rc = FetchLaneStates(d, n, lanes, s);
  if (rc) return rc;
  for (int32_t i = 0; i < n; ++i) {
    const int32_t lane = lanes[i];
    d->frames[lane] = 0;
    d->bp_valid[lane] = 0;
    d->rf_cache[lane] = -1;
    d->status[lane] = d->h_lanes[lane].status;
    if (d->status[lane] != 0)
      return Fail(KD_ERR_OVERFLOW, std::string("InitDecoding: ") + StatusText(d->status[lane]));
  }
  return KD_OK;
}

int kd_decoder_advance_async(kd_decoder *d, int32_t n, const int32_t *lanes,
                             const float *const *logprobs, const int32_t *rows, int32_t cols,
                             const int32_t *offsets, int32_t max_num_frames, int mem_kind,
                             int flags, void *producer_stream, int64_t *ticket) {
  if (ticket) *ticket = -1;
  int rc = CheckLanes(d, n, lanes);
  if (rc) return rc;
  if (n == 0) return KD_OK;
  if (!logprobs || !rows) return Fail(KD_ERR_INVALID, "null logprobs/rows");
  if (mem_kind != KD_MEM_HOST && mem_kind != KD_MEM_DEVICE)
    return Fail(KD_ERR_INVALID, "bad mem_kind");
  if (flags & ~(KD_ADVANCE_INIT | KD_ADVANCE_FINALIZE)) return Fail(KD_ERR_INVALID, "bad flags");
  if (cols < d->g->max_ilabel)
    return Fail(KD_ERR_INVALID,
                "decodable has fewer columns than the largest ilabel of the graph "
                "(the reference would read out of bounds, decodable-ctc.cc:28)");
  KD_CUDA(cudaSetDevice(d->device));
  // a lane is owned by one call at a time: calls in flight on these lanes complete first
  rc = WaitLanes(d, n, lanes);
  if (rc && !(flags & KD_ADVANCE_INIT)) return rc;

  // per-lane targets (faster-decoder.cc:128-144)
  struct Work {
    int32_t lane, decoded, target, n_rows;
    const float *src;
  };
  std::vector<Work> work;
  work.reserve(n);
  size_t stage_need = 0;
  int32_t max_rows = 0;
  for (int32_t i = 0; i < n; ++i) {
    const int32_t lane = lanes[i];
    const int32_t decoded = (flags & KD_ADVANCE_INIT) ? 0 : d->frames[lane];
    if (decoded < 0)
      return Fail(KD_ERR_INVALID,
                  "Check failed!\nx: num_frames_decoded_ >= 0 && \"You must call "
                  "InitDecoding() before AdvanceDecoding()\"");
    if (d->status[lane] != 0 && !(flags & KD_ADVANCE_INIT))
      return Fail(KD_ERR_OVERFLOW, std::string("lane is in error state: ") +
                                       StatusText(d->status[lane]));
    if (rows[i] < 0) return Fail(KD_ERR_INVALID, "negative rows");
    const int32_t off = offsets ? offsets[i] : 0;
    const int32_t ready = off + rows[i];
    if (ready < decoded)
      return Fail(KD_ERR_INVALID, "Check failed!\nx: num_frames_ready >= num_frames_decoded_");
    int32_t target = ready;
    if (max_num_frames >= 0) target = std::min(target, decoded + max_num_frames);
    // (a lane with nothing to decode still takes part when it is to be initialised or finalized)
    if (target <= decoded && flags == 0) continue;
    if (target < decoded) target = decoded;
    if (target > decoded) {
      if (decoded < off)
        return Fail(KD_ERR_INVALID, "decodable offset is beyond the frames decoded so far");
      if (!logprobs[i]) return Fail(KD_ERR_INVALID, "null log-prob matrix");
    }
    Work w;
    w.lane = lane;
    w.decoded = decoded;
    w.target = target;
    w.n_rows = target - decoded;
    w.src = w.n_rows > 0 ? logprobs[i] + static_cast<size_t>(decoded - off) * cols : nullptr;
    work.push_back(w);
    stage_need += static_cast<size_t>(w.n_rows) * cols;
    max_rows = std::max(max_rows, w.n_rows);
  }
  if (work.empty()) return KD_OK;
  const int32_t m = static_cast<int32_t>(work.size());

  // A free slot.  Slots that have been used before keep their staging and result buffers:
  // they are taken in turn (least recently used first, so the results of a call stay readable
  // for at least the next two calls); a fresh slot is only opened when every used one is in
  // flight (or fewer than three are in use).  If all are in flight the oldest completes first.
  int si = -1, n_used = 0;
  for (int i = 0; i < kNumSlots; ++i) {
    if (d->slots[i].ticket >= 0) ++n_used;
    if (!d->slots[i].busy && d->slots[i].ticket >= 0 &&
        (si < 0 || d->slots[i].ticket < d->slots[si].ticket))
      si = i;
  }
  if (si < 0 || n_used < 3) {
    for (int i = 0; i < kNumSlots; ++i)
      if (d->slots[i].ticket < 0) {
        si = i;
        break;
      }
  }
  if (si < 0) {
    for (int i = 0; i < kNumSlots; ++i)
      if (si < 0 || d->slots[i].ticket < d->slots[si].ticket) si = i;
    CompleteSlot(d, si);  // its outcome stays in the slot for kd_decoder_wait
  }
  kd_slot &s = d->slots[si];
  s.res_valid = false;
  s.kernel_ms = 0.f;
  s.launches = 0;
  s.flags = flags;
  s.host_input = mem_kind == KD_MEM_HOST;
  s.lanes.resize(m);
  s.targets.resize(m);
  s.path_cap = 0;
  if (flags & KD_ADVANCE_FINALIZE) {
    // room for the path of the longest lane: its frames so far plus half as many epsilon
    // arcs (a longer path is fetched through kd_best_fill_kernel instead)
    int32_t longest = 0;
    for (int32_t i = 0; i < m; ++i) longest = std::max(longest, work[i].target);
    int64_t cap = (static_cast<int64_t>(longest) * 3 / 2 + 64 + 31) / 32 * 32;
    cap = std::min<int64_t>(cap, d->qcap / 2);
    s.path_cap = static_cast<int32_t>(cap);
    const size_t need = static_cast<size_t>(m) * 4 * s.path_cap;
    if (need > s.res_words) {
      if (s.h_res) cudaFreeHost(s.h_res);
      s.h_res = nullptr;
      s.res_words = 0;
      KD_CUDA(cudaMallocHost(reinterpret_cast<void **>(&s.h_res), need * sizeof(int32_t)));
      s.res_words = need;
    }
  }

  kd::Params P = MakeParams(d);
  P.cols = cols;
  P.row_in_smem = (static_cast<size_t>(cols) * 6 <= 49152) ? 1 : 0;
  P.items = s.d_items;
  P.n_items = m;
  P.work_counter = s.d_words;
  P.yield_flag = s.d_words + 1;
  P.progress = nullptr;
  s.threads = PickThreads(d, m);

  if (s.host_input && stage_need > s.stage_floats) {
    // (cudaFree waits for the device: growth is rare and only happens while ramping up)
    cudaFree(s.d_stage);
    s.d_stage = nullptr;
    s.stage_floats = 0;
    rc = DevAlloc(&s.d_stage, stage_need);
    if (rc) return rc;
    s.stage_floats = stage_need;
  }
  size_t pos = 0;
  for (int32_t i = 0; i < m; ++i) {
    const int32_t lane = work[i].lane;
    s.lanes[i] = lane;
    s.targets[i] = work[i].target;
    kd::AdvanceItem &it = s.h_items[i];
    it.lane = lane;
    it.rows = work[i].n_rows;
    it.offset = work[i].decoded;
    it.target = work[i].target;
    it.flags = flags;
    it.path_cap = s.path_cap;
    if (s.host_input) {
      it.logp = s.d_stage + pos;
      pos += static_cast<size_t>(work[i].n_rows) * cols;
    } else {
      it.logp = work[i].src;
    }
    if ((flags & KD_ADVANCE_INIT) && d->status[lane] != 0) {
      // a lane that overflowed may have left claimed table slots behind
      KD_CUDA(cudaMemsetAsync(d->table + static_cast<size_t>(lane) * d->hcap, 0xFF,
                              d->hcap * sizeof(kd::Entry), s.sc));
      KD_CUDA(cudaMemsetAsync(d->bitmap + static_cast<size_t>(lane) * (d->hcap / 32), 0,
                              d->hcap / 8, s.sc));
      d->status[lane] = 0;
    }
  }
  KD_CUDA(cudaMemcpyAsync(s.d_items, s.h_items, sizeof(kd::AdvanceItem) * m,
                          cudaMemcpyHostToDevice, s.sc));

  // From here on work is enqueued: the slot is in flight, and a failure completes it.
  s.busy = true;
  s.ticket = d->next_ticket++;
  s.rc = KD_OK;
  s.P = P;
  for (int32_t i = 0; i < m; ++i) {
    d->lane_slot[s.lanes[i]] = static_cast<int8_t>(si);
    d->bp_valid[s.lanes[i]] = 0;
    d->rf_cache[s.lanes[i]] = -1;
    if (flags & KD_ADVANCE_INIT) d->frames[s.lanes[i]] = 0;
  }
  auto enqueue = [&]() -> int {
    if (!s.host_input) {
      // Device matrices: the search must not start before the work that produces them.
      // `producer_stream` is the stream that work was enqueued on (NULL: the legacy default
      // stream, which also orders behind every blocking stream).
      cudaStream_t ps = producer_stream ? static_cast<cudaStream_t>(producer_stream)
                                        : cudaStreamLegacy;
      KD_CUDA(cudaEventRecord(s.ev_dep, ps));
      KD_CUDA(cudaStreamWaitEvent(s.sc, s.ev_dep, 0));
      return EnqueueSearch(d, s);
    }
    // Host matrices: ONE search launch; the copy stream delivers the frames in time chunks
    // (all lanes, `chunk_frames` frames each) and publishes its progress in a device word
    // the lanes poll when they run out of rows.  Lanes are latency bound and independent,
    // so every lane starts as soon as its first frames are there and never waits for other
    // lanes (relaunching per chunk costs the slowest-lane tail once per chunk: measured
    // 189 vs 170 ms per step).  The copies are enqueued BEFORE the search kernel: under
    // serialised launches (CUDA_LAUNCH_BLOCKING, profilers) a kernel launched first would
    // poll for rows nobody can enqueue any more.
    const int32_t F = d->chunk_frames;
    // Chunk ends: the first chunks are small (F/16, F/8, ... frames) so the lanes start
    // after a fraction of a millisecond instead of one full chunk's copy time; the copy
    // engine delivers frames about twice as fast as the lanes consume them, so it stays
    // ahead once it is.
    std::vector<int32_t> ends;
    for (int32_t r = 0, step = std::max(4, F / 16); r < max_rows;) {
      r = std::min(max_rows, r + step);
      ends.push_back(r);
      step = std::min(F, step * 2);
    }
    const int32_t n_chunks = static_cast<int32_t>(ends.size());
    if (n_chunks > s.progress_cap) {
      if (s.h_progress) cudaFreeHost(s.h_progress);
      s.h_progress = nullptr;
      s.progress_cap = 0;
      KD_CUDA(cudaMallocHost(reinterpret_cast<void **>(&s.h_progress),
                             sizeof(int32_t) * (static_cast<size_t>(n_chunks) + 64)));
      s.progress_cap = n_chunks + 64;
    }
    // one 2-D copy per chunk when the host matrices are equally long and equally spaced
    // (and the spacing is a pitch the copy engine takes)
    bool uniform = m > 1;
    ptrdiff_t src_stride = 0;
    if (uniform) {
      src_stride = work[1].src - work[0].src;
      if (src_stride < static_cast<ptrdiff_t>(static_cast<size_t>(work[0].n_rows) * cols) ||
          static_cast<size_t>(src_stride) * sizeof(float) > d->mem_pitch ||
          static_cast<size_t>(work[0].n_rows) * cols * sizeof(float) > d->mem_pitch)
        uniform = false;  // overlapping, unordered, identical or too far apart
      for (int32_t i = 1; i < m && uniform; ++i) {
        if (work[i].n_rows != work[0].n_rows) uniform = false;
        if (work[i].src - work[i - 1].src != src_stride) uniform = false;
      }
    }
    KD_CUDA(cudaMemsetAsync(s.d_words + 2, 0, sizeof(int32_t), s.sc));
    KD_CUDA(cudaEventRecord(s.ev_gate, s.sc));
    KD_CUDA(cudaStreamWaitEvent(s.sx, s.ev_gate, 0));  // copies start after the progress reset
    s.P.progress = s.d_words + 2;
    // Per-lane copies (matrices of different lengths or spacing) cost ~3 us of host time
    // each: only the first ~2000 go out before the launch, the rest behind it; should the
    // launch turn out to be blocking, lanes that run dry yield and are launched again.
    int64_t copies_before_launch = (uniform || d->launch_blocking) ? (1ll << 60) : 2048;
    // (test knob: launch after this many copies whatever the layout -- with serialised
    // launches this exercises the yield-and-launch-again path)
    if (const char *e = getenv("KD_B200_COPIES_BEFORE_LAUNCH")) {
      copies_before_launch = atoll(e);
      if (uniform && copies_before_launch <= 0) copies_before_launch = 1;
    }
    bool launched = false;
    for (int32_t c = 0; c < n_chunks; ++c) {
      const int32_t r0 = c == 0 ? 0 : ends[c - 1];
      const int32_t Fc = ends[c] - r0;  // frames of this chunk
      if (uniform) {
        const int32_t rows_c = std::min(Fc, work[0].n_rows - r0);
        const size_t lane_floats = static_cast<size_t>(work[0].n_rows) * cols;
        KD_CUDA(cudaMemcpy2DAsync(s.d_stage + static_cast<size_t>(r0) * cols,
                                  lane_floats * sizeof(float),
                                  work[0].src + static_cast<size_t>(r0) * cols,
                                  static_cast<size_t>(src_stride) * sizeof(float),
                                  static_cast<size_t>(rows_c) * cols * sizeof(float), m,
                                  cudaMemcpyHostToDevice, s.sx));
        --copies_before_launch;
      } else {
        for (int32_t i = 0; i < m; ++i) {
          if (r0 >= work[i].n_rows) continue;
          const int32_t rows_c = std::min(Fc, work[i].n_rows - r0);
          KD_CUDA(cudaMemcpyAsync(
              const_cast<float *>(s.h_items[i].logp) + static_cast<size_t>(r0) * cols,
              work[i].src + static_cast<size_t>(r0) * cols,
              sizeof(float) * static_cast<size_t>(rows_c) * cols, cudaMemcpyHostToDevice, s.sx));
          --copies_before_launch;
        }
      }
      s.h_progress[c] = ends[c];
      KD_CUDA(cudaMemcpyAsync(s.d_words + 2, s.h_progress + c, sizeof(int32_t),
                              cudaMemcpyHostToDevice, s.sx));
      if (!launched && copies_before_launch <= 0) {
        const int r = EnqueueSearch(d, s);
        if (r) return r;
        launched = true;
      }
    }
    if (!launched) return EnqueueSearch(d, s);
    return KD_OK;
  };
  rc = enqueue();
  if (rc) {
    // nothing of a failed call keeps running; its lanes are unusable until InitDecoding
    const std::string keep = g_error;
    cudaStreamSynchronize(s.sx);
    cudaStreamSynchronize(s.sc);
    for (int32_t i = 0; i < m; ++i) {
      d->lane_slot[s.lanes[i]] = -1;
      d->status[s.lanes[i]] |= kd::kStatusInputStall;
    }
    s.busy = false;
    s.rc = rc;
    s.msg = keep;
    g_error = keep;
    return rc;
  }
  if (ticket) *ticket = s.ticket;
  return KD_OK;
}

int kd_decoder_wait(kd_decoder *d, int64_t ticket) {
  if (!d) return Fail(KD_ERR_INVALID, "decoder is null");
  if (ticket < 0) return WaitAll(d);
  for (int i = 0; i < kNumSlots; ++i) {
    kd_slot &s = d->slots[i];
    if (s.ticket != ticket) continue;
    if (s.busy) return CompleteSlot(d, i);
    if (s.rc) g_error = s.msg;  // completed by a later call on its lanes
    return s.rc;
  }
  return KD_OK;  // long done: its slot has been reused
}

int kd_decoder_advance(kd_decoder *d, int32_t n, const int32_t *lanes,
                       const float *const *logprobs, const int32_t *rows, int32_t cols,
                       const int32_t *offsets, int32_t max_num_frames, int mem_kind) {
  int64_t ticket = -1;
  int rc = kd_decoder_advance_async(d, n, lanes, logprobs, rows, cols, offsets, max_num_frames,
                                    mem_kind, 0, nullptr, &ticket);
  if (rc || ticket < 0) return rc;
  return kd_decoder_wait(d, ticket);
}

int kd_decoder_result_view(kd_decoder *d, int64_t ticket, int use_final_probs,
                           int32_t *num_lanes, const int32_t **lanes, const int32_t **words,
                           const int64_t **word_offsets, int64_t *num_arcs, int32_t *ok,
                           int32_t *reached_final, float *final_weight2) {
  if (!d) return Fail(KD_ERR_INVALID, "decoder is null");
  int si = -1;
  for (int i = 0; i < kNumSlots; ++i)
    if (d->slots[i].ticket == ticket) si = i;
  if (si < 0 || ticket < 0)
    return Fail(KD_ERR_INVALID, "unknown ticket (its slot has been reused by a later call)");
  kd_slot &s = d->slots[si];
  if (s.busy) {
    int rc = CompleteSlot(d, si);
    if (rc) return rc;
  }
  if (s.rc) {
    g_error = s.msg;
    return s.rc;
  }
  if (!(s.flags & KD_ADVANCE_FINALIZE) || !s.res_valid)
    return Fail(KD_ERR_INVALID, "the call was not made with KD_ADVANCE_FINALIZE");
  const int32_t m = static_cast<int32_t>(s.lanes.size());
  if (num_lanes) *num_lanes = m;
  if (lanes) *lanes = s.lanes.data();
  if (words) *words = s.h_res;
  if (word_offsets) *word_offsets = s.res_off.data();
  for (int32_t i = 0; i < m; ++i) {
    const int32_t lane = s.lanes[i];
    if (d->lane_slot[lane] >= 0 || !d->bp_valid[lane])
      return Fail(KD_ERR_INVALID, "a lane of this call has been advanced or initialised since");
    const kd::LaneState &L = d->h_lanes[lane];
    d->use_final[lane] = use_final_probs ? 1 : 0;
    if (ok) ok[i] = L.bp_ok;
    if (reached_final) reached_final[i] = L.bp_final;
    if (num_arcs) num_arcs[i] = L.bp_ok ? L.bp_len : 0;
    if (final_weight2) {
      // faster-decoder.cc:416-421
      const bool fin = L.bp_ok && L.bp_final && use_final_probs;
      final_weight2[2 * i] = fin ? L.bp_final_w : 0.f;
      final_weight2[2 * i + 1] = 0.f;
    }
  }
  return KD_OK;
}

int kd_decoder_num_frames_decoded(kd_decoder *d, int32_t lane, int32_t *out) {
  int rc = CheckLanes(d, 1, &lane);
  if (rc) return rc;
  WaitLanes(d, 1, &lane);
  if (out) *out = d->frames[lane];
  return KD_OK;
}

int kd_decoder_best_path_prepare(kd_decoder *d, int32_t n, const int32_t *lanes,
                                 int use_final_probs, int32_t *ok, int32_t *reached_final,
                                 int64_t *num_arcs) {
  int rc = CheckLanes(d, n, lanes);
  if (rc) return rc;
  if (n == 0) return KD_OK;
  KD_CUDA(cudaSetDevice(d->device));
  rc = WaitLanes(d, n, lanes);
  if (rc) return rc;
  // lanes whose last search launch already selected their path (KD_ADVANCE_FINALIZE) are done
  int32_t todo = 0;
  for (int32_t i = 0; i < n; ++i) {
    if (d->frames[lanes[i]] < 0)
      return Fail(KD_ERR_INVALID, "lane not initialised (call InitDecoding first)");
    if (d->bp_valid[lanes[i]]) continue;
    memset(&d->h_items[todo], 0, sizeof(kd::AdvanceItem));
    d->h_items[todo].lane = lanes[i];
    ++todo;
  }
  if (todo > 0) {
    cudaStream_t s = d->stream;
    KD_CUDA(cudaMemcpyAsync(d->d_items, d->h_items, sizeof(kd::AdvanceItem) * todo,
                            cudaMemcpyHostToDevice, s));
    kd::Params P = MakeParams(d);
    P.items = d->d_items;
    P.n_items = todo;
    kd::kd_best_select_kernel<256><<<todo, 256, 0, s>>>(P);
    KD_CUDA(cudaGetLastError());
    rc = FetchLaneStates(d, n, lanes, s);
    if (rc) return rc;
  }
  for (int32_t i = 0; i < n; ++i) {
    const int32_t lane = lanes[i];
    d->bp_valid[lane] = 1;
    d->use_final[lane] = use_final_probs ? 1 : 0;
    const kd::LaneState &L = d->h_lanes[lane];
    d->rf_cache[lane] = static_cast<int8_t>(L.bp_final != 0);
    if (ok) ok[i] = L.bp_ok;
    if (reached_final) reached_final[i] = L.bp_final;
    if (num_arcs) num_arcs[i] = L.bp_ok ? L.bp_len : 0;
  }
  return KD_OK;
}

int kd_decoder_best_path_view(kd_decoder *d, int32_t n, const int32_t *lanes,
                              const int64_t *out_offsets, int64_t total_arcs,
                              const int32_t **ilabel, const int32_t **olabel,
                              const float **graph_cost, const float **acoustic_cost,
                              float *final_weight2) {
  int rc = CheckLanes(d, n, lanes);
  if (rc) return rc;
  if (ilabel) *ilabel = nullptr;
  if (olabel) *olabel = nullptr;
  if (graph_cost) *graph_cost = nullptr;
  if (acoustic_cost) *acoustic_cost = nullptr;
  if (n == 0) return KD_OK;
  if (!out_offsets || total_arcs < 0) return Fail(KD_ERR_INVALID, "bad output layout");
  KD_CUDA(cudaSetDevice(d->device));
  rc = WaitLanes(d, n, lanes);
  if (rc) return rc;
  for (int32_t i = 0; i < n; ++i) {
    if (!d->bp_valid[lanes[i]])
      return Fail(KD_ERR_INVALID,
                  "kd_decoder_best_path_prepare must run after the last InitDecoding / "
                  "AdvanceDecoding of the lane");
    const kd::LaneState &L = d->h_lanes[lanes[i]];
    if (L.bp_ok && (out_offsets[i] < 0 || out_offsets[i] + L.bp_len > total_arcs))
      return Fail(KD_ERR_INVALID, "best path does not fit the output arrays");
    memset(&d->h_items[i], 0, sizeof(kd::AdvanceItem));
    d->h_items[i].lane = lanes[i];
    d->h_out_off[i] = out_offsets[i];
    if (final_weight2) {
      // faster-decoder.cc:416-421
      const bool fin = L.bp_ok && L.bp_final && d->use_final[lanes[i]];
      final_weight2[2 * i] = fin ? L.bp_final_w : 0.f;
      final_weight2[2 * i + 1] = 0.f;
    }
  }
  if (total_arcs == 0) return KD_OK;
  if (total_arcs > d->path_cap) {
    KD_CUDA(cudaStreamSynchronize(d->stream));
    cudaFree(d->d_path);
    if (d->h_path) cudaFreeHost(d->h_path);
    d->d_path = d->h_path = nullptr;
    d->path_cap = 0;
    const size_t cap = static_cast<size_t>(total_arcs) + static_cast<size_t>(total_arcs) / 4 + 1024;
    if ((rc = DevAlloc(&d->d_path, 4 * cap))) return rc;
    KD_CUDA(cudaMallocHost(reinterpret_cast<void **>(&d->h_path), 4 * cap * sizeof(int32_t)));
    d->path_cap = static_cast<int64_t>(cap);
  }
  cudaStream_t s = d->stream;
  KD_CUDA(cudaMemcpyAsync(d->d_items, d->h_items, sizeof(kd::AdvanceItem) * n,
                          cudaMemcpyHostToDevice, s));
  KD_CUDA(cudaMemcpyAsync(d->d_out_off, d->h_out_off, sizeof(long long) * n,
                          cudaMemcpyHostToDevice, s));
  kd::Params P = MakeParams(d);
  P.items = d->d_items;
  P.n_items = n;
  // the four arrays are packed total_arcs apart: one contiguous copy brings them back
  const size_t nb = static_cast<size_t>(total_arcs);
  int32_t *d_il = d->d_path, *d_ol = d->d_path + nb;
  float *d_gw = reinterpret_cast<float *>(d->d_path + 2 * nb);
  float *d_aw = reinterpret_cast<float *>(d->d_path + 3 * nb);
  kd::kd_best_fill_kernel<<<n, 128, 0, s>>>(P, d->d_out_off, d_il, d_ol, d_gw, d_aw);
  KD_CUDA(cudaGetLastError());
  KD_CUDA(cudaMemcpyAsync(d->h_path, d->d_path, 4 * nb * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  KD_CUDA(cudaStreamSynchronize(s));
  if (ilabel) *ilabel = d->h_path;
  if (olabel) *olabel = d->h_path + nb;
  if (graph_cost) *graph_cost = reinterpret_cast<const float *>(d->h_path + 2 * nb);
  if (acoustic_cost) *acoustic_cost = reinterpret_cast<const float *>(d->h_path + 3 * nb);
  return KD_OK;
}

int kd_decoder_best_path_fetch(kd_decoder *d, int32_t n, const int32_t *lanes,
                               const int64_t *out_offsets, int64_t total_arcs, int32_t *ilabel,
                               int32_t *olabel, float *graph_cost, float *acoustic_cost,
                               float *final_weight2) {
  const int32_t *il = nullptr, *ol = nullptr;
  const float *gw = nullptr, *aw = nullptr;
  int rc = kd_decoder_best_path_view(d, n, lanes, out_offsets, total_arcs, &il, &ol, &gw, &aw,
                                     final_weight2);
  if (rc || total_arcs <= 0 || il == nullptr) return rc;
  const size_t bytes = static_cast<size_t>(total_arcs) * 4;
  if (ilabel) memcpy(ilabel, il, bytes);
  if (olabel) memcpy(olabel, ol, bytes);
  if (graph_cost) memcpy(graph_cost, gw, bytes);
  if (acoustic_cost) memcpy(acoustic_cost, aw, bytes);
  return KD_OK;
}

int kd_decoder_best_path(kd_decoder *d, int32_t lane, int use_final_probs, int64_t cap,
                         int32_t *ilabel, int32_t *olabel, float *graph_cost,
                         float *acoustic_cost, int64_t *num_arcs, float final_weight2[2],
                         int32_t *reached_final, int32_t *ok) {
  int32_t okv = 0, rf = 0;
  int64_t len = 0;
  int rc = kd_decoder_best_path_prepare(d, 1, &lane, use_final_probs, &okv, &rf, &len);
  if (rc) return rc;
  if (num_arcs) *num_arcs = len;
  if (reached_final) *reached_final = rf;
  if (ok) *ok = okv;
  if (final_weight2) final_weight2[0] = final_weight2[1] = 0.f;
  if (!okv) return KD_OK;
  if (len > cap) return Fail(KD_ERR_INVALID, "best path longer than the output capacity");
  int64_t off = 0;
  return kd_decoder_best_path_fetch(d, 1, &lane, &off, len, ilabel, olabel, graph_cost,
                                    acoustic_cost, final_weight2);
}

int kd_decoder_final_relative_cost(kd_decoder *d, int32_t lane, float *out) {
  int32_t okv = 0, rf = 0;
  int64_t len = 0;
  int rc = CheckLanes(d, 1, &lane);
  if (rc) return rc;
  // (the selection does not depend on use_final_probs: the lane keeps its own setting)
  rc = kd_decoder_best_path_prepare(d, 1, &lane, d->use_final[lane], &okv, &rf, &len);
  if (rc) return rc;
  const kd::LaneState &L = d->h_lanes[lane];
  float v = std::numeric_limits<float>::infinity();
  // with a final state active the selection cost is min(cost + final weight)
  if (okv && rf) v = static_cast<float>(L.bp_value - L.best_cost);
  if (v != v) v = std::numeric_limits<float>::infinity();  // simple-decoder.cc:94-98
  if (out) *out = v;
  return KD_OK;
}

int kd_decoder_reached_final(kd_decoder *d, int32_t lane, int32_t *out) {
  int rc = CheckLanes(d, 1, &lane);
  if (rc) return rc;
  KD_CUDA(cudaSetDevice(d->device));
  rc = WaitLanes(d, 1, &lane);
  if (rc) return rc;
  if (d->frames[lane] < 0)
    return Fail(KD_ERR_INVALID, "lane not initialised (call InitDecoding first)");
  // faster-decoder.cc:347-354 only: no best-token selection, no backpointer walk; the
  // answer is kept until the lane's tokens change
  if (d->rf_cache[lane] < 0) {
    cudaStream_t s = d->stream;
    memset(&d->h_items[0], 0, sizeof(kd::AdvanceItem));
    d->h_items[0].lane = lane;
    KD_CUDA(cudaMemcpyAsync(d->d_items, d->h_items, sizeof(kd::AdvanceItem),
                            cudaMemcpyHostToDevice, s));
    kd::Params P = MakeParams(d);
    P.items = d->d_items;
    P.n_items = 1;
    kd::kd_reached_final_kernel<256><<<1, 256, 0, s>>>(P, d->d_flags);
    KD_CUDA(cudaGetLastError());
    KD_CUDA(cudaMemcpyAsync(d->h_flags, d->d_flags, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    KD_CUDA(cudaStreamSynchronize(s));
    d->rf_cache[lane] = static_cast<int8_t>(d->h_flags[0] != 0);
  }
  if (out) *out = d->rf_cache[lane];
  return KD_OK;
}

int kd_decoder_dump_tokens(kd_decoder *d, int32_t lane, int64_t cap, int32_t *states,
                           double *costs, int64_t *n) {
  int rc = CheckLanes(d, 1, &lane);
  if (rc) return rc;
  KD_CUDA(cudaSetDevice(d->device));
  rc = WaitLanes(d, 1, &lane);
  if (rc) return rc;
  if (d->frames[lane] < 0) {
    if (n) *n = 0;
    return KD_OK;
  }
  rc = FetchLaneStates(d, 1, &lane, d->stream);
  if (rc) return rc;
  const kd::LaneState &L = d->h_lanes[lane];
  if (n) *n = L.n_live;
  if (L.n_live == 0 || L.n_live > cap) return KD_OK;
  const size_t base = static_cast<size_t>(lane) * static_cast<size_t>(d->arena_cap) + L.tok_base;
  // the block may hold holes (state -1): records that are not tokens are squeezed out here
  std::vector<int32_t> st(static_cast<size_t>(L.n_tok));
  std::vector<double> co(static_cast<size_t>(L.n_tok));
  KD_CUDA(cudaMemcpy(st.data(), d->a_state + base, sizeof(int32_t) * L.n_tok,
                     cudaMemcpyDeviceToHost));
  KD_CUDA(cudaMemcpy(co.data(), d->a_cost + base, sizeof(double) * L.n_tok,
                     cudaMemcpyDeviceToHost));
  // SimpleDecoder search: PruneToks (simple-decoder.cc:251-279) is applied to the view
  const bool pruned_view = d->simple && L.frames_decoded > 0;
  const double limit = L.best_cost + static_cast<double>(d->opts.beam);
  int64_t k = 0;
  for (int32_t i = 0; i < L.n_tok && k < L.n_live; ++i) {
    if (st[i] < 0) continue;
    if (pruned_view && !(co[i] < limit)) continue;
    if (states) states[k] = st[i];
    if (costs) costs[k] = co[i];
    ++k;
  }
  if (n) *n = k;
  return KD_OK;
}

int kd_decoder_stats(kd_decoder *d, int32_t lane, kd_stats *out) {
  if (!d || !out) return Fail(KD_ERR_INVALID, "null argument");
  if (lane >= d->max_lanes) return Fail(KD_ERR_INVALID, "lane id out of range");
  KD_CUDA(cudaSetDevice(d->device));
  WaitAll(d);
  memset(out, 0, sizeof(*out));
  KD_CUDA(cudaMemcpy(d->h_lanes, d->lanes, sizeof(kd::LaneState) * d->max_lanes,
                     cudaMemcpyDeviceToHost));
  const int32_t lo = lane < 0 ? 0 : lane, hi = lane < 0 ? d->max_lanes : lane + 1;
  for (int32_t l = lo; l < hi; ++l) {
    if (d->frames[l] < 0) continue;
    const kd::LaneState &L = d->h_lanes[l];
    out->frames += L.st_frames;
    out->tokens_in += L.st_tokens_in;
    out->tokens_expanded += L.st_expanded;
    out->emit_arcs += L.st_emit_arcs;
    out->eps_arcs += L.st_eps_arcs;
    out->tokens_out += L.st_tokens_out;
    out->max_tokens = std::max<int64_t>(out->max_tokens, L.st_max_tokens);
    out->eps_sweeps += L.st_sweeps;
    out->cycles_cutoff += L.cyc_cutoff;
    out->cycles_expand += L.cyc_expand;
    out->cycles_closure += L.cyc_closure;
    out->cycles_commit += L.cyc_commit;
    out->slots_claimed += L.st_claimed;
    out->candidates += L.st_cand;
    out->arcs_evaluated += L.st_items;
    out->cycles_scan += L.cyc_scan;
    out->arena_compactions += L.st_compactions;
    out->cycles_input_wait += L.cyc_wait;
    out->table_retries += L.st_redo;
  }
  return KD_OK;
}

int kd_decoder_last_advance_info(kd_decoder *d, float *kernel_ms, int32_t *launches) {
  if (!d) return Fail(KD_ERR_INVALID, "decoder is null");
  if (kernel_ms) *kernel_ms = d->last_kernel_ms;
  if (launches) *launches = d->last_launches;
  return KD_OK;
}

int kd_decoder_span_begin(kd_decoder *d) {
  if (!d) return Fail(KD_ERR_INVALID, "decoder is null");
  d->span_armed = true;
  d->span_open = false;
  d->span_launches = 0;
  return KD_OK;
}

int kd_decoder_span_end(kd_decoder *d, float *ms, int32_t *launches) {
  if (!d) return Fail(KD_ERR_INVALID, "decoder is null");
  KD_CUDA(cudaSetDevice(d->device));
  int rc = WaitAll(d);
  if (ms) *ms = 0.f;
  if (launches) *launches = d->span_launches;
  const bool open = d->span_open;
  d->span_armed = d->span_open = false;
  if (rc) return rc;
  if (open && ms) KD_CUDA(cudaEventElapsedTime(ms, d->ev_span_begin, d->ev_span_end));
  return KD_OK;
}

int kd_decoder_info(kd_decoder *d, int64_t info[6]) {
  if (!d || !info) return Fail(KD_ERR_INVALID, "null argument");
  info[0] = d->max_lanes;
  info[1] = d->hcap;
  info[2] = d->arena_cap;
  info[3] = d->threads;
  info[4] = static_cast<int64_t>(d->device_bytes);
  info[5] = d->chunk_frames;
  return KD_OK;
}

}  // extern "C"
