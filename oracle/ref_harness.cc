// oracle/ref_harness.cc  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C-ABI wrapper around the *unmodified* reference decoder.  The Makefile in
// this directory compiles /root/reference/kaldi-decoder/csrc/faster-decoder.cc
// where it lies (never copied into this repo) together with this file into
// oracle/_ref/libkd_ref.so.  OpenFst/kaldifst are replaced by the repo's
// OpenFst-shaped value types (kaldi-decoder_b200/csrc/minifst); Eigen is
// replaced by a pointer decodable with the arithmetic of
// decodable-ctc.cc:22-31.
//
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
// --impl reference legs may load this library.
//
// The probe subclass only *reads* protected state (FasterDecoder's members
// are `protected`, faster-decoder.h:109) to expose the per-frame token list in
// the reference's own HashList order; it changes no behaviour.

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <exception>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include <unordered_map>

#include "kaldi-decoder/csrc/faster-decoder.h"
#include "kaldi-decoder/csrc/simple-decoder.h"

namespace {

thread_local std::string g_error;

// Same arithmetic as DecodableCtc (decodable-ctc.cc:22-38).
class PtrDecodable : public kaldi_decoder::DecodableInterface {
 public:
  PtrDecodable(const float *p, int32_t rows, int32_t cols, int32_t offset)
      : p_(p), rows_(rows), cols_(cols), offset_(offset) {}
  float LogLikelihood(int32_t frame, int32_t index) override {
    return *(p_ + static_cast<int64_t>(frame - offset_) * cols_ + index - 1);
  }
  int32_t NumFramesReady() const override { return offset_ + rows_; }
  int32_t NumIndices() const override { return cols_; }
  bool IsLastFrame(int32_t frame) const override {
    return frame == NumFramesReady() - 1;
  }

 private:
  const float *p_;
  int32_t rows_, cols_, offset_;
};

class Probe : public kaldi_decoder::FasterDecoder {
 public:
  using kaldi_decoder::FasterDecoder::FasterDecoder;

  int64_t NumTokens() const {
    int64_t n = 0;
    for (const Elem *e = toks_.GetList(); e != nullptr; e = e->tail) ++n;
    return n;
  }
  // tokens in HashList order
  int64_t DumpTokens(int64_t cap, int32_t *states, double *costs) const {
    int64_t n = 0;
    for (const Elem *e = toks_.GetList(); e != nullptr; e = e->tail, ++n) {
      if (n < cap) {
        states[n] = e->key;
        costs[n] = e->val->cost_;
      }
    }
    return n;
  }
};

struct Graph {
  std::unique_ptr<fst::ConstFst<fst::StdArc>> fst;
};

struct Decoder {
  const Graph *graph;
  std::unique_ptr<Probe> dec;
};

kaldi_decoder::FasterDecoderOptions MakeOpts(float beam, int32_t max_active,
                                             int32_t min_active,
                                             float beam_delta,
                                             float hash_ratio) {
  return kaldi_decoder::FasterDecoderOptions(beam, max_active, min_active,
                                             beam_delta, hash_ratio);
}

// Flattens the linear lattice returned by GetBestPath.
int64_t FlattenLinear(const fst::Lattice &lat, int64_t cap, int32_t *il,
                      int32_t *ol, float *graph, float *ac, float *final2) {
  int64_t n = 0;
  final2[0] = final2[1] = 0;
  if (lat.Start() == fst::kNoStateId) return 0;
  int s = lat.Start();
  while (true) {
    if (lat.NumArcs(s) == 0) {
      fst::LatticeWeight f = lat.Final(s);
      final2[0] = f.Value1();
      final2[1] = f.Value2();
      break;
    }
    fst::ArcIterator<fst::Lattice> it(lat, s);
    const fst::LatticeArc &a = it.Value();
    if (n < cap) {
      il[n] = a.ilabel;
      ol[n] = a.olabel;
      graph[n] = a.weight.Value1();
      ac[n] = a.weight.Value2();
    }
    ++n;
    s = a.nextstate;
  }
  return n;
}

}  // namespace

extern "C" {

const char *kdref_last_error() { return g_error.c_str(); }

// CSR graph in original arc order: arcs of state s are [row_off[s], row_off[s+1])
void *kdref_graph_create(int32_t num_states, int32_t start,
                         const int64_t *row_off, const int32_t *ilabel,
                         const int32_t *olabel, const float *weight,
                         const int32_t *nextstate, const float *final_w) {
  try {
    std::vector<size_t> off(row_off, row_off + num_states + 1);
    std::vector<fst::StdArc> arcs;
    arcs.reserve(off.back());
    for (size_t i = 0; i < off.back(); ++i)
      arcs.emplace_back(ilabel[i], olabel[i], fst::TropicalWeight(weight[i]),
                        nextstate[i]);
    std::vector<fst::TropicalWeight> fin(num_states);
    for (int32_t s = 0; s < num_states; ++s) fin[s] = final_w[s];
    auto *g = new Graph;
    g->fst.reset(new fst::ConstFst<fst::StdArc>(
        start, std::move(off), std::move(arcs), std::move(fin)));
    return g;
  } catch (const std::exception &e) {
    g_error = e.what();
    return nullptr;
  }
}

void kdref_graph_destroy(void *g) { delete static_cast<Graph *>(g); }

void *kdref_decoder_create(void *graph, float beam, int32_t max_active,
                           int32_t min_active, float beam_delta,
                           float hash_ratio) {
  try {
    auto *d = new Decoder;
    d->graph = static_cast<Graph *>(graph);
    d->dec.reset(new Probe(
        *d->graph->fst,
        MakeOpts(beam, max_active, min_active, beam_delta, hash_ratio)));
    return d;
  } catch (const std::exception &e) {
    g_error = e.what();
    return nullptr;
  }
}

void kdref_decoder_destroy(void *d) { delete static_cast<Decoder *>(d); }

int kdref_decoder_set_options(void *d, float beam, int32_t max_active,
                              int32_t min_active, float beam_delta,
                              float hash_ratio) {
  static_cast<Decoder *>(d)->dec->SetOptions(
      MakeOpts(beam, max_active, min_active, beam_delta, hash_ratio));
  return 0;
}

int kdref_decoder_init(void *d) {
  try {
    static_cast<Decoder *>(d)->dec->InitDecoding();
    return 0;
  } catch (const std::exception &e) {
    g_error = e.what();
    return -1;
  }
}

int kdref_decoder_advance(void *d, const float *logp, int32_t rows,
                          int32_t cols, int32_t offset,
                          int32_t max_num_frames) {
  try {
    PtrDecodable dec(logp, rows, cols, offset);
    static_cast<Decoder *>(d)->dec->AdvanceDecoding(&dec, max_num_frames);
    return 0;
  } catch (const std::exception &e) {
    g_error = e.what();
    return -1;
  }
}

int kdref_decoder_decode(void *d, const float *logp, int32_t rows,
                         int32_t cols) {
  try {
    PtrDecodable dec(logp, rows, cols, 0);
    static_cast<Decoder *>(d)->dec->Decode(&dec);
    return 0;
  } catch (const std::exception &e) {
    g_error = e.what();
    return -1;
  }
}

int32_t kdref_decoder_num_frames_decoded(void *d) {
  return static_cast<Decoder *>(d)->dec->NumFramesDecoded();
}

int kdref_decoder_reached_final(void *d) {
  return static_cast<Decoder *>(d)->dec->ReachedFinal() ? 1 : 0;
}

int64_t kdref_decoder_dump_tokens(void *d, int64_t cap, int32_t *states,
                                  double *costs) {
  return static_cast<Decoder *>(d)->dec->DumpTokens(cap, states, costs);
}

// Returns the number of arcs of the (RemoveEpsLocal'ed) best path, or -1 if
// GetBestPath returned false, or -2 on exception.  final2 = (graph, acoustic)
// final weight.
int64_t kdref_decoder_best_path(void *d, int use_final_probs, int64_t cap,
                                int32_t *il, int32_t *ol, float *graph,
                                float *ac, float *final2) {
  try {
    fst::Lattice lat;
    bool ok = static_cast<Decoder *>(d)->dec->GetBestPath(&lat,
                                                         use_final_probs != 0);
    if (!ok) return -1;
    return FlattenLinear(lat, cap, il, ol, graph, ac, final2);
  } catch (const std::exception &e) {
    g_error = e.what();
    return -2;
  }
}

// Decodes n_utts utterances, one utterance per thread at a time, one
// FasterDecoder per thread, graph shared read-only.  logp is
// [n_utts][max_rows][cols] row-major (utterance u uses its first rows[u]
// rows).  Outputs per utterance u are written at [u * cap, u * cap + n[u]).
// Returns wall-clock seconds of the decode region (Decode + ReachedFinal +
// GetBestPath for all utterances), or a negative number on error.
double kdref_decode_batch(void *graph, const float *logp, int32_t n_utts,
                          int32_t max_rows, const int32_t *rows, int32_t cols,
                          float beam, int32_t max_active, int32_t min_active,
                          float beam_delta, float hash_ratio,
                          int use_final_probs, int32_t num_threads,
                          int64_t cap, int32_t *il, int32_t *ol, float *gw,
                          float *aw, float *final2, int64_t *n_out,
                          int32_t *reached_final) {
  auto *g = static_cast<Graph *>(graph);
  if (num_threads < 1) num_threads = 1;
  std::atomic<int32_t> next{0};
  std::atomic<int> failed{0};
  std::string err;
  auto worker = [&]() {
    try {
      Probe dec(*g->fst,
                MakeOpts(beam, max_active, min_active, beam_delta, hash_ratio));
      while (true) {
        int32_t u = next.fetch_add(1);
        if (u >= n_utts) break;
        PtrDecodable decodable(
            logp + static_cast<int64_t>(u) * max_rows * cols, rows[u], cols, 0);
        dec.Decode(&decodable);
        int rf = dec.ReachedFinal() ? 1 : 0;
        fst::Lattice lat;
        bool ok = dec.GetBestPath(&lat, use_final_probs != 0);
        if (reached_final) reached_final[u] = rf;
        if (n_out) {
          if (!ok) {
            n_out[u] = -1;
          } else {
            n_out[u] = FlattenLinear(lat, cap, il + u * cap, ol + u * cap,
                                     gw + u * cap, aw + u * cap,
                                     final2 + 2 * u);
          }
        }
      }
    } catch (const std::exception &e) {
      if (failed.exchange(1) == 0) err = e.what();
    }
  };
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> threads;
  for (int32_t t = 1; t < num_threads; ++t) threads.emplace_back(worker);
  worker();
  for (auto &t : threads) t.join();
  auto t1 = std::chrono::steady_clock::now();
  if (failed.load()) {
    g_error = err;
    return -1.0;
  }
  return std::chrono::duration<double>(t1 - t0).count();
}

// ---------------------------------------------------------------- SimpleDecoder
// (kaldi-decoder/csrc/simple-decoder.{h,cc}, unmodified)

struct SimpleHandle {
  const Graph *graph;
  std::unique_ptr<kaldi_decoder::SimpleDecoder> dec;
};

void *kdref_simple_create(void *graph, float beam) {
  try {
    auto *g = static_cast<Graph *>(graph);
    auto *h = new SimpleHandle;
    h->graph = g;
    h->dec.reset(new kaldi_decoder::SimpleDecoder(*g->fst, beam));
    return h;
  } catch (const std::exception &e) {
    g_error = e.what();
    return nullptr;
  }
}

void kdref_simple_destroy(void *d) { delete static_cast<SimpleHandle *>(d); }

int kdref_simple_init(void *d) {
  try {
    static_cast<SimpleHandle *>(d)->dec->InitDecoding();
    return 0;
  } catch (const std::exception &e) {
    g_error = e.what();
    return -1;
  }
}

int kdref_simple_advance(void *d, const float *logp, int32_t rows, int32_t cols, int32_t offset,
                         int32_t max_num_frames) {
  try {
    PtrDecodable dec(logp, rows, cols, offset);
    static_cast<SimpleHandle *>(d)->dec->AdvanceDecoding(&dec, max_num_frames);
    return 0;
  } catch (const std::exception &e) {
    g_error = e.what();
    return -1;
  }
}

int32_t kdref_simple_num_frames_decoded(void *d) {
  return static_cast<SimpleHandle *>(d)->dec->NumFramesDecoded();
}

int kdref_simple_reached_final(void *d) {
  return static_cast<SimpleHandle *>(d)->dec->ReachedFinal() ? 1 : 0;
}

float kdref_simple_final_relative_cost(void *d) {
  return static_cast<SimpleHandle *>(d)->dec->FinalRelativeCost();
}

int64_t kdref_simple_best_path(void *d, int use_final_probs, int64_t cap, int32_t *il,
                               int32_t *ol, float *graph, float *ac, float *final2) {
  try {
    fst::Lattice lat;
    bool ok = static_cast<SimpleHandle *>(d)->dec->GetBestPath(&lat, use_final_probs != 0);
    if (!ok) return -1;
    return FlattenLinear(lat, cap, il, ol, graph, ac, final2);
  } catch (const std::exception &e) {
    g_error = e.what();
    return -2;
  }
}

}  // extern "C"
