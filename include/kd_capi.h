/* include/kd_capi.h
 *
 * C ABI of the B200-native token-passing Viterbi beam search.
 *
 * This is the drop-in boundary for the reference's FasterDecoder hot path
 * (k2-fsa/kaldi-decoder, kaldi-decoder/csrc/faster-decoder.{h,cc}): plain
 * pointers and sizes, no C++ or torch types.  The C++ classes in
 * kaldi-decoder_b200/csrc/ (same names and signatures as the reference's) and
 * the Python bindings are thin layers over exactly these entry points.
 * Every function returns KD_OK (0) or a negative status; kd_last_error() gives
 * the message for the calling thread (the reference signals errors by
 * throwing std::runtime_error from KALDI_DECODER_ERR / KALDI_DECODER_ASSERT,
 * log.h:46-51,86-89 -- the C++ wrapper rethrows from these statuses).
 *
 * There is NO CPU fallback: every call needs a CUDA device (sm_100a build).
 *
 * Execution model.  A kd_decoder owns `max_lanes` independent utterance lanes
 * (one reference FasterDecoder object == one lane).  All per-lane search
 * state (tokens, recombination table, backpointer store) stays resident in
 * device memory between calls, so InitDecoding / AdvanceDecoding streaming
 * works across calls exactly as in the reference (faster-decoder.cc:126-152).
 */
#ifndef KD_CAPI_H_
#define KD_CAPI_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define KD_API __declspec(dllexport)
#else
#define KD_API __attribute__((visibility("default")))
#endif

enum {
  KD_OK = 0,
  KD_ERR_INVALID = -1,   /* bad argument / failed reference-style assertion   */
  KD_ERR_CUDA = -2,      /* CUDA runtime error                                */
  KD_ERR_OVERFLOW = -3,  /* a per-lane device capacity was exceeded           */
  KD_ERR_NO_DEVICE = -4  /* no usable CUDA device                             */
};

enum { KD_MEM_HOST = 0, KD_MEM_DEVICE = 1 };

typedef struct kd_graph kd_graph;
typedef struct kd_decoder kd_decoder;

/* FasterDecoderOptions (faster-decoder.h:24-63).  hash_ratio is accepted for
 * signature compatibility and only validated (>= 1.0, faster-decoder.cc:24). */
typedef struct kd_options {
  float beam;
  int32_t max_active;
  int32_t min_active;
  float beam_delta;
  float hash_ratio;
} kd_options;

/* Device capacities; any field <= 0 picks a default. */
typedef struct kd_decoder_config {
  int32_t max_lanes;        /* number of utterance lanes (default 1)                   */
  int32_t hash_capacity;    /* per-lane recombination-table entries, power of two;
                               tokens alive in one frame must stay <= capacity / 2     */
  int64_t arena_records;    /* per-lane backpointer-store records, 20 bytes each.  The store
                               is garbage-collected on the device when it fills up (the
                               reference frees dead tokens by reference counting,
                               faster-decoder.h:145-155): it has to hold the tokens still
                               reachable from the live ones plus the current frame, not the
                               sum over all frames                                     */
  int32_t threads_per_lane; /* 128/160/192/224/256/384/512; default: the widest
                               that keeps all lanes of a call resident at once (160 with
                               7 lanes per SM is the best when the chip is kept full by
                               several calls in flight)                                */
  int32_t chunk_frames;     /* host-memory advance: frames per copy/search pipeline
                               stage (default 128)                                     */
  int32_t search;           /* KD_SEARCH_FASTER (default) or KD_SEARCH_SIMPLE          */
} kd_decoder_config;

/* Which reference decoder the lanes reproduce.
 *   KD_SEARCH_FASTER  FasterDecoder (faster-decoder.cc), all kd_options fields used.
 *   KD_SEARCH_SIMPLE  SimpleDecoder (simple-decoder.cc:150-279): beam-only pruning
 *                     (kd_options.beam; max_active / min_active are ignored), token cost
 *                     = prev + float(w + ac), PruneToks applied to what ReachedFinal and
 *                     GetBestPath see. */
#define KD_SEARCH_FASTER 0
#define KD_SEARCH_SIMPLE 1

/* Search counters, summed over the frames decoded since kd_decoder_init.  They
 * define the algorithmic bytes of SURVEY.md §8(d):
 *   16*(emit_arcs + eps_arcs) + 16*tokens_in + 24*tokens_out + 4*cols*frames */
typedef struct kd_stats {
  int64_t frames;
  int64_t tokens_in;       /* tokens at frame start                               */
  int64_t tokens_expanded; /* tokens with cost < weight_cutoff                    */
  int64_t emit_arcs;       /* emitting arcs visited                               */
  int64_t eps_arcs;        /* epsilon arcs visited in the closure                 */
  int64_t tokens_out;      /* tokens alive at frame end                           */
  int64_t max_tokens;      /* max tokens alive at a frame end                     */
  int64_t eps_sweeps;      /* closure sweeps                                      */
  /* SM cycles per phase (clock64 of the lane's thread 0, summed over frames) */
  int64_t cycles_cutoff;   /* GetCutoff + row staging + seed                      */
  int64_t cycles_expand;   /* emitting expansion + recombination                  */
  int64_t cycles_closure;  /* epsilon closure                                     */
  int64_t cycles_commit;   /* token block commit + table wipe                     */
  int64_t slots_claimed;   /* recombination-table slots claimed (>= tokens_out)   */
  int64_t candidates;      /* emitting arcs that passed the running-cutoff filter */
  int64_t arcs_evaluated;  /* arcs actually loaded: scanned, or found through a label
                              table; emit_arcs - this = arcs skipped as provably pruned */
  int64_t cycles_scan;     /* part of cycles_expand before the exact cutoff       */
  int64_t arena_compactions; /* garbage collections of the backpointer arena (a lane's records
                              are compacted when a frame's tokens no longer fit)   */
  int64_t cycles_input_wait; /* SM cycles lanes spent waiting for log-prob rows still on their
                              way from the host (KD_MEM_HOST): the share of the search time
                              that the upload, not the search, decides                */
  int64_t table_retries;   /* frames searched twice: a frame hashes into a region of the lane's
                              table sized from its candidates (locality); when the epsilon
                              closure fills that region the frame starts over with the whole
                              table                                                    */
} kd_stats;

KD_API const char *kd_last_error(void);

/* Number of CUDA devices visible (0 when there is none). */
KD_API int kd_device_count(int *count);

/* ---- graph: replaces the `const fst::Fst<fst::StdArc>&` the reference binds
 * (faster-decoder.cc:21-23, faster-decoder.h:179).  Input is the FST as CSR in
 * its original arc order: arcs of state s are [row_offsets[s], row_offsets[s+1]).
 * final_weight[s] = +inf for non-final states (TropicalWeight::Zero()).
 * The graph is copied to `device` as a split (emitting / epsilon) 16-byte-arc
 * CSR; it is immutable and may be shared by any number of decoders. */
KD_API int kd_graph_create(int device, int32_t num_states, int32_t start,
                           const int64_t *row_offsets, const int32_t *ilabel,
                           const int32_t *olabel, const float *weight,
                           const int32_t *nextstate, const float *final_weight,
                           kd_graph **out);
KD_API int kd_graph_destroy(kd_graph *g);
/* info[0..4] = num_states, num_arcs, num_epsilon_arcs, max_ilabel, device */
KD_API int kd_graph_info(const kd_graph *g, int64_t info[5]);

/* ---- decoder: FasterDecoder::FasterDecoder (faster-decoder.cc:21-32); the
 * option checks of lines 24-28 are enforced here and in kd_decoder_set_options. */
KD_API int kd_decoder_create(kd_graph *g, const kd_options *opts,
                             const kd_decoder_config *cfg, kd_decoder **out);
KD_API int kd_decoder_destroy(kd_decoder *d);
/* FasterDecoder::SetOptions (faster-decoder.h:78) */
KD_API int kd_decoder_set_options(kd_decoder *d, const kd_options *opts);
/* Back to the state kd_decoder_create left it in -- every lane uninitialised, calls in flight
 * completed and forgotten -- keeping the device buffers, streams and pinned memory: what a host
 * wrapper calls to hand a decoder from one short-lived FasterDecoder object to the next (the
 * reference's scripts construct one per utterance, faster-decoder.cc:21-32) instead of paying
 * kd_decoder_destroy + kd_decoder_create each time. */
KD_API int kd_decoder_reset(kd_decoder *d);

/* FasterDecoder::InitDecoding (faster-decoder.cc:42-56) for n lanes. */
KD_API int kd_decoder_init(kd_decoder *d, int32_t n, const int32_t *lanes);

/* FasterDecoder::AdvanceDecoding (faster-decoder.cc:126-152) with a
 * DecodableCtc per lane (decodable-ctc.cc:22-31): lane lanes[i] sees the
 * row-major float matrix logprobs[i] of rows[i] x cols whose first row is frame
 * offsets[i] (offsets may be NULL = all 0); NumFramesReady = offsets[i] + rows[i].
 * max_num_frames < 0 means "all ready frames".  mem_kind says where the
 * matrices live; host matrices are copied (pinned memory makes the copy
 * asynchronous and overlapped with the search).  Returns when all lanes have
 * advanced. */
KD_API int kd_decoder_advance(kd_decoder *d, int32_t n, const int32_t *lanes,
                              const float *const *logprobs, const int32_t *rows,
                              int32_t cols, const int32_t *offsets,
                              int32_t max_num_frames, int mem_kind);

/* ---- the same call without waiting, and with InitDecoding / GetBestPath folded in.
 *
 * kd_decoder_advance_async enqueues the call and returns; kd_decoder_wait(ticket) completes
 * it and reports its outcome (ticket < 0: every call in flight).  Up to 16 calls on
 * disjoint lanes are in flight at once, each on its own streams: the upload and the search
 * of one batch of lanes overlap the search and the download of another, and the lanes of
 * the next batch take over the SMs as the slowest lanes of the previous one finish.  Any
 * other call that touches a lane first completes the call that owns it (deferred
 * synchronisation), so the synchronous API can be mixed in freely.  Host matrices must
 * stay valid and unchanged until the call has completed.
 *
 * flags:
 *   KD_ADVANCE_INIT      InitDecoding (faster-decoder.cc:42-56) of every listed lane first,
 *                        inside the same kernel launch.
 *   KD_ADVANCE_FINALIZE  after the last frame, select each lane's best path
 *                        (faster-decoder.cc:356-402) inside the same launch and bring its
 *                        arcs back behind it: kd_decoder_result_view then needs no kernel,
 *                        and kd_decoder_best_path_* / kd_decoder_reached_final on these
 *                        lanes reuse the selection.
 * With both flags a whole Decode() + GetBestPath() of a batch is ONE kernel launch.
 *
 * producer_stream (KD_MEM_DEVICE only): the CUDA stream (cudaStream_t) on which the work
 * producing the matrices was enqueued; the search is ordered behind it.  NULL means the
 * legacy default stream (which in turn waits for all blocking streams).  The decoder's own
 * streams are non-blocking: without this ordering a log_softmax still running on another
 * stream would race with the search.  kd_decoder_advance orders behind the legacy default
 * stream. */
#define KD_ADVANCE_INIT 1
#define KD_ADVANCE_FINALIZE 2
KD_API int kd_decoder_advance_async(kd_decoder *d, int32_t n, const int32_t *lanes,
                                    const float *const *logprobs, const int32_t *rows,
                                    int32_t cols, const int32_t *offsets,
                                    int32_t max_num_frames, int mem_kind, int flags,
                                    void *producer_stream, int64_t *ticket);
KD_API int kd_decoder_wait(kd_decoder *d, int64_t ticket);

/* The best paths of a KD_ADVANCE_FINALIZE call (completes it if it is still in flight).
 * *lanes lists the *num_lanes lanes the call advanced, in call order; lane i's arcs are
 * four arrays of num_arcs[i] words -- ilabel, olabel (int32), graph cost, acoustic cost
 * (float) -- starting at (*words)[(*word_offsets)[4 * i + 0..3]].  The pointers address
 * pinned host memory of the decoder and stay valid until the third-next
 * kd_decoder_advance_async call or the decoder's destruction.  ok / reached_final /
 * final_weight2 (2 per lane) as kd_decoder_best_path_prepare / _fetch; any may be NULL,
 * arrays must hold *num_lanes entries (at most the n of the call). */
KD_API int kd_decoder_result_view(kd_decoder *d, int64_t ticket, int use_final_probs,
                                  int32_t *num_lanes, const int32_t **lanes,
                                  const int32_t **words, const int64_t **word_offsets,
                                  int64_t *num_arcs, int32_t *ok, int32_t *reached_final,
                                  float *final_weight2);

/* FasterDecoder::NumFramesDecoded (faster-decoder.h:107); -1 before init. */
KD_API int kd_decoder_num_frames_decoded(kd_decoder *d, int32_t lane, int32_t *out);

/* FasterDecoder::ReachedFinal (faster-decoder.cc:347-354). */
KD_API int kd_decoder_reached_final(kd_decoder *d, int32_t lane, int32_t *out);

/* SimpleDecoder::FinalRelativeCost (simple-decoder.cc:78-101): best cost with final
 * weights minus best cost over the live tokens; +inf if no final state is active or
 * no token is alive.  Either search mode. */
KD_API int kd_decoder_final_relative_cost(kd_decoder *d, int32_t lane, float *out);

/* FasterDecoder::GetBestPath (faster-decoder.cc:356-424) up to, and not
 * including, RemoveEpsLocal: one (ilabel, olabel, graph cost, acoustic cost)
 * arc per token on the best path, in time order; final_weight = (graph,
 * acoustic) of the final state as the reference sets it at lines 416-421.
 * Step 1 selects the best token of each lane and measures its path; step 2
 * writes lane lanes[i]'s arcs at out_offsets[i].. of the host arrays. */
KD_API int kd_decoder_best_path_prepare(kd_decoder *d, int32_t n, const int32_t *lanes,
                                        int use_final_probs, int32_t *ok,
                                        int32_t *reached_final, int64_t *num_arcs);
KD_API int kd_decoder_best_path_fetch(kd_decoder *d, int32_t n, const int32_t *lanes,
                                      const int64_t *out_offsets, int64_t total_arcs,
                                      int32_t *ilabel, int32_t *olabel, float *graph_cost,
                                      float *acoustic_cost, float *final_weight2);
/* Step 2 without the copy into caller arrays: the pointers returned address the
 * decoder's own pinned host buffer (same layout: lane lanes[i]'s arcs at
 * out_offsets[i]..) and stay valid until the next best-path call on this decoder
 * or its destruction. */
KD_API int kd_decoder_best_path_view(kd_decoder *d, int32_t n, const int32_t *lanes,
                                     const int64_t *out_offsets, int64_t total_arcs,
                                     const int32_t **ilabel, const int32_t **olabel,
                                     const float **graph_cost, const float **acoustic_cost,
                                     float *final_weight2);
/* Single-lane convenience: both steps; *num_arcs is the path length even if it
 * exceeds cap (then nothing is written and KD_ERR_INVALID is returned). */
KD_API int kd_decoder_best_path(kd_decoder *d, int32_t lane, int use_final_probs, int64_t cap,
                                int32_t *ilabel, int32_t *olabel, float *graph_cost,
                                float *acoustic_cost, int64_t *num_arcs,
                                float final_weight2[2], int32_t *reached_final, int32_t *ok);

/* The live tokens of a lane (state, cost), in device order (unordered): the
 * contents of the reference's `toks_` list.  *n is the token count even if it
 * exceeds cap (then nothing is written).  KD_SEARCH_SIMPLE: what is written is
 * SimpleDecoder's cur_toks_ after PruneToks, and *n is then the number written. */
KD_API int kd_decoder_dump_tokens(kd_decoder *d, int32_t lane, int64_t cap, int32_t *states,
                                  double *costs, int64_t *n);

/* Counters of one lane, or summed over all lanes when lane < 0. */
KD_API int kd_decoder_stats(kd_decoder *d, int32_t lane, kd_stats *out);

/* Device time (ms, CUDA events on the decoder's streams) of the search kernels
 * of the last kd_decoder_advance call, and how many kernels it launched. */
KD_API int kd_decoder_last_advance_info(kd_decoder *d, float *kernel_ms, int32_t *launches);

/* Device time of a run of search launches that may overlap (calls in flight on several
 * streams): kd_decoder_span_begin arms the measurement, the next launch opens it;
 * kd_decoder_span_end completes every call in flight and returns the CUDA-event time from
 * the start of the first launch to the end of the most recently enqueued one, and the
 * number of search kernels launched in between. */
KD_API int kd_decoder_span_begin(kd_decoder *d);
KD_API int kd_decoder_span_end(kd_decoder *d, float *ms, int32_t *launches);

/* info[0..5] = max_lanes, hash_capacity, arena_records, threads_per_lane,
 *              device bytes allocated, chunk_frames */
KD_API int kd_decoder_info(kd_decoder *d, int64_t info[6]);

#ifdef __cplusplus
}
#endif

#endif /* KD_CAPI_H_ */
