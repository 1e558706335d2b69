// kaldi-decoder_b200/csrc/kd_kernels.cuh
//
// sm_100a kernels of the token-passing Viterbi beam search.  One CTA owns one
// utterance lane for a whole AdvanceDecoding call (all its frames), so the
// frame loop needs no grid-wide synchronisation: every phase boundary is a
// __syncthreads().  Reference semantics implemented:
//
//   GetCutoff           faster-decoder.cc:244-336   -> lane_cutoff()
//   ProcessEmitting     faster-decoder.cc:155-241   -> lane_expand_emitting()
//   ProcessNonemitting  faster-decoder.cc:59-119    -> lane_closure_and_commit()
//   token list / Token  faster-decoder.h:110-156,
//                       hash-list-inl.h:127-173     -> per-lane open-addressing
//                       table (key = state, 128-bit value = ordered fp64 cost |
//                       arc | backpointer; entries claimed through an occupancy
//                       bitmap) + frame-major backpointer arena, collected in place
//   InitDecoding        faster-decoder.cc:42-56     -> first pass of a lane in
//                                                      kd_advance_kernel, kd_init_kernel
//   ReachedFinal /      faster-decoder.cc:347-354,
//   GetBestPath         356-424                     -> last pass of a lane in
//                                                      kd_advance_kernel (lane_finalize),
//                                                      kd_best_select_kernel,
//                                                      kd_best_fill_kernel,
//                                                      kd_reached_final_kernel
//
// Costs are fp64 sums of fp32 addends, added in the reference's order
// ((w + cost) + ac, faster-decoder.cc:210), so they are bit-identical to the
// CPU's.  The reference's running next-frame cutoff is order dependent; its
// final value C* = min(new_weight) + adaptive_beam is not.  The kernel keeps a
// running cutoff in shared memory only as a filter (anything admitted with
// cost >= C* is ignored later), which makes the surviving token set exactly
// {new_weight < C*}, independent of thread scheduling.  Equal-cost arrivals at
// a state are resolved towards the lowest emitting-arc index; in the epsilon
// closure an incumbent from the emitting phase stays (as in the reference) and
// equal-cost epsilon arrivals go to the lowest epsilon-arc index: every
// backpointer is deterministic.
//
// What bounds the search (profiles/r2_*): not bandwidth but (a) the
// dependent memory round trips of a lane-frame (~24, plus ~27 barriers, whatever the
// token count -- and half of the frames have fewer than 350 tokens), executed by 5
// warps, (b) the work of the 6 other lanes on the SM, and (c) the instruction cache: 7
// lanes per SM are in 7 different phases of this kernel, and its hot body must stay
// small.  Hence: little unrolling, cold paths out of line, label tables so that a tenth
// of the arcs is evaluated, candidates filtered by the exact cutoff before they touch
// the table, first arrivals that write their whole entry with one 256-bit store, a
// single-pass commit, the next frame's row fetched by a TMA bulk copy one frame ahead.
#ifndef KD_KERNELS_CUH_
#define KD_KERNELS_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

// Build-time switches of individual optimisations (A/B builds: -DKD_OPT_...=0).
#ifndef KD_OPT_QSTATE
#define KD_OPT_QSTATE 1       // epsilon worklist entries carry the state: its record is requested with the entry
#endif
#ifndef KD_OPT_GRAPH_EL
#define KD_OPT_GRAPH_EL 1     // graph loads (state records, label tables, arcs) carry an L2 evict-last policy
#endif
#ifndef KD_OPT_SINGLE_PASS
#define KD_OPT_SINGLE_PASS 1  // token blocks of at most one scan tile are scanned in one pass
#endif
#ifndef KD_OPT_CAND_CACHED
#define KD_OPT_CAND_CACHED 0  // candidate buffer / front list with default (write-back) caching instead of streaming
#endif
#ifndef KD_SINGLE_PASS_TILES
#define KD_SINGLE_PASS_TILES 1   // blocks of at most this many scan tiles take one pass
#endif
#ifndef KD_OPT_PREFETCH_NO
#define KD_OPT_PREFETCH_NO 0  // L2 prefetch of a candidate's (nextstate, olabel) record when it is appended
#endif
#ifndef KD_OPT_DEFER
#define KD_OPT_DEFER 1         // arrivals at states already in the table are recombined in a second pass
#endif
#ifndef KD_OPT_COARSE
#define KD_OPT_COARSE 1       // item -> token search: top level in its own bank-conflict-free array
#endif
#ifndef KD_OPT_ST256
#define KD_OPT_ST256 1        // state records loaded with one 256-bit load
#endif
#ifndef KD_OPT_SLIST
#define KD_OPT_SLIST 1        // head of the frame's slot list in shared memory
#endif

namespace kd {

constexpr int kStatusHashOverflow = 1;
constexpr int kStatusArenaOverflow = 2;
constexpr int kStatusQueueOverflow = 4;
constexpr int kStatusInputStall = 8;  // streamed log-probs never arrived (set by the host)
// ~1 s of polling before a lane yields.  (Long on purpose: with many calls in flight a call's
// rows queue behind the other calls' copies for tens of milliseconds.)
constexpr long long kYieldCycles = 2000000000ll;
constexpr int kStatusCandOverflow = 16;  // SimpleDecoder search: candidate buffer full

// In an arc field: "epsilon arc".  In a nextstate field: "state has epsilon arcs".
constexpr uint32_t kEpsFlag = 0x80000000u;
constexpr uint32_t kNoArc = 0x7FFFFFFFu;  // the start token's "arc"
constexpr uint32_t kNoPrev = 0xFFFFFFFFu;
constexpr int32_t kEmptyKey = -1;
constexpr unsigned long long kEmptyCost = 0xFFFFFFFFFFFFFFFFull;
constexpr unsigned long long kEmptyArg = 0xFFFFFFFFFFFFFFFFull;
constexpr uint32_t kNoIdx = 0xFFFFFFFFu;
constexpr uint32_t kRetry = 0xFFFFFFFEu;  // table_arrive: the entry is claimed but not written yet
constexpr int kLabelTableMinDegree = 16;  // states with at least this many emitting arcs get a label table
constexpr int kOrderBins = 512;           // buckets of the per-frame label order (1/16 wide)
constexpr int kMaxOrderCols = 2048;       // widest log-prob row for which the label order is built
// 32-item windows a warp takes per scan step.  2 or 3 (more loads in flight per warp) make
// the scan 5% faster and the other phases slower by as much (code size, registers).
#ifndef KD_WIN
#define KD_WIN 1
#endif
constexpr int kWin = KD_WIN;
#ifndef KD_TILE_TOKENS
#define KD_TILE_TOKENS 2
#endif
constexpr int kTileTokens = KD_TILE_TOKENS;  // tokens per thread in one scan tile
#ifndef KD_LIST_SMEM
#define KD_LIST_SMEM 1024
#endif
constexpr int kListSmem = KD_OPT_SLIST ? KD_LIST_SMEM : 0;  // slot-list entries kept in shared memory
constexpr int kFrontCap = 2048;           // records of the per-lane front list (>= the largest scan tile)
constexpr uint32_t kLookupFlag = 0x80000000u;  // in t_beg: expand this token by label lookup  // commit numbering: token goes behind the "good" ones

struct __align__(16) HVal {
  unsigned long long cost;  // order-preserving image of the fp64 cost
  unsigned long long arg;   // (arc << 32) | prev
};

// One recombination-table entry = one 32-byte sector: a probe, the 128-bit
// value CAS, the commit numbering and the wipe of a state all touch the same
// sector (the first layout used three arrays = three sectors per state and
// thrashed L2: profiles/r1_v0_ncu_summary.txt).
struct __align__(32) Entry {
  HVal val;        // 16-byte aligned: target of the 128-bit CAS
  int32_t key;     // state id
  uint32_t idx;    // position in this frame's slot list = index of the token in the next block
  uint32_t epoch;  // the lane-frame that wrote the entry: any other value = stale contents
  uint32_t pad;
};

// Which entries are in use is kept in a per-lane BITMAP (one bit per entry), not in the
// entries: claiming an entry is one atomicOr on a word that lives in L2 (32 KB per lane at
// 2^18 entries, re-used every frame), and the thread that flips the bit writes the whole
// 32-byte entry -- key, slot-list position, epoch and ITS OWN value -- with one 256-bit store
// (STG.256: a full sector, so L2 does not fetch the old contents from DRAM).  The first
// arrival at a state, 80 % of all arrivals, thus costs one L2 round trip; before, it cost a
// probe load that missed to DRAM (the table is 8 MB per lane), a CAS on the key and a CAS on
// the value.  Entries are never wiped: the commit clears the bits, and a later arrival that
// finds a bit set validates the entry by its epoch (a thread that finds the bit set before
// the owner's store has landed re-reads until it has).

// Everything the device keeps per lane between calls.
// Everything the device keeps per lane between calls.
struct __align__(16) LaneState {
  int32_t n_tok;           // records in the current token block (n_live tokens + dead records)
  int32_t frames_decoded;  // num_frames_decoded_; -1 before InitDecoding
  int32_t status;          // kStatus* bits; non-zero = lane unusable until init
  int32_t best_idx;        // index (in the current token block) of a best token
  int32_t best_state;      // its state (-1 if none)
  int32_t n_live;          // tokens alive (the reference's toks_ list length)
  int32_t n_front;         // tokens below good_cut; the first kFrontCap of them are in the front list
  int32_t n_mid;           // tokens below mid_cut
  uint32_t tok_base;       // arena index of the current token block
  uint32_t arena_used;     // arena records in use
  double best_cost;        // min cost over the current tokens (+inf if none)
  double good_cut;         // tokens below it are scanned first in the next frame
  double mid_cut;          // n_mid tokens lie below it (lets GetCutoff skip its counting pass)
  // counters (kd_stats)
  long long st_frames, st_tokens_in, st_expanded, st_emit_arcs, st_eps_arcs,
      st_tokens_out, st_max_tokens, st_sweeps;
  // SM cycles spent per phase (clock64 of thread 0), for the phase breakdown
  long long cyc_cutoff, cyc_expand, cyc_closure, cyc_commit, cyc_scan;
  long long st_claimed;  // table slots claimed (tokens + arrivals later found >= C*)
  long long st_compactions;  // arena garbage collections
  long long cyc_wait;        // SM cycles spent waiting for streamed log-prob rows (host input)
  uint32_t epoch;          // stamps the table entries of the current pass; survives InitDecoding
  uint32_t st_redo;        // frames searched a second time because their table region filled up
  long long st_cand;     // emitting arcs that passed the running-cutoff filter
  long long st_items;    // arcs actually evaluated (scanned + looked up)
  // best-path selection results
  int32_t bp_ok, bp_final, bp_best_state;
  uint32_t bp_best_tok;    // arena index
  long long bp_len;
  float bp_final_w;
  double bp_value;         // selection cost of the best token (cost, or cost + final weight)
  int32_t bp_stored;       // the path's token indices are in the lane's candidate buffer
  int32_t bp_parked;       // the path's arcs are in the lane's worklist buffer: 4 arrays, this far apart
};

constexpr int32_t kItemInit = 1;      // InitDecoding first (faster-decoder.cc:42-56), in the same launch
constexpr int32_t kItemFinalize = 2;  // then select the best path and park its arcs (GetBestPath)

struct AdvanceItem {
  int32_t lane;
  int32_t rows;
  int32_t offset;
  int32_t target;       // frames_decoded to reach
  const float *logp;    // device pointer, row-major rows x cols
  int32_t flags;        // kItem*
  int32_t path_cap;     // kItemFinalize: arcs per parked array (il | ol | graph | acoustic)
};

struct Params {
  // graph (device): CSR split into emitting and epsilon arcs.  The scan of the
  // emitting arcs only needs (ilabel, weight): they are an 8-byte array of
  // their own; (nextstate, olabel) are read for admitted arcs only.
  const int4 *st;      // [2S] {emit_begin, emit_count, eps_begin, eps_count},
                       //      {label-table row or -1, smallest emitting weight bits, 0, 0}
  const int2 *labtab;  // [rows][lab_stride]: ilabel-1 -> {weight bits, arc offset within the state or -1}
  int32_t lab_stride;
  int32_t simple;      // 1: SimpleDecoder semantics (simple-decoder.cc), else FasterDecoder
  const int2 *e_iw;    // [Ee] {ilabel, weight bits}
  const int2 *e_no;    // [Ee] {nextstate | kEpsFlag if that state has eps arcs, olabel}
  const int4 *n_arc;   // [En] {olabel, weight bits, nextstate | kEpsFlag ..., 0}
  const float *fin;    // [S]
  int32_t start;
  // options
  float beam;
  int32_t max_active;
  int32_t min_active;
  float beam_delta;
  // lanes
  LaneState *lanes;
  const AdvanceItem *items;
  int32_t n_items;
  int32_t *work_counter;
  // per-lane storage: lane L uses [L * stride, (L + 1) * stride)
  double *a_cost;
  unsigned long long *a_link;
  int32_t *a_state;
  long long arena_cap;
  Entry *table;
  uint32_t *bitmap;  // hcap / 32 words per lane: entries in use
  uint32_t *list;
  uint2 *queue;     // 2 * qcap per lane: {table slot, state}
  uint4 *cand;      // ccap per lane: arcs that passed the running-cutoff filter
  uint4 *front;     // kFrontCap per lane: {cost, state, number} of the tokens close to the best
  uint32_t hcap, hmask, lcap, qcap, ccap;
  int32_t hshift;
  // a frame's region of the table: region_f entries per candidate, at least region_min
  // (a power of two); region_f = 0: always the whole table
  uint32_t region_f, region_min;
  int32_t cols;
  int32_t row_in_smem;
  // host-memory advance: rows [0, *progress) of every lane's staged matrix have
  // arrived (written by the copy stream while the kernel runs); nullptr = all present
  const int32_t *progress;
  int32_t *yield_flag;  // set by the first lane that gave up waiting for rows
};

// ------------------------------------------------------------------ helpers

__device__ __forceinline__ unsigned long long dkey(double x) {
  x = x + 0.0;  // -0.0 -> +0.0 so that key order == fp64 order
  unsigned long long u = static_cast<unsigned long long>(__double_as_longlong(x));
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

__device__ __forceinline__ double dunkey(unsigned long long k) {
  unsigned long long u = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
  return __longlong_as_double(static_cast<long long>(u));
}

__device__ __forceinline__ uint32_t fkey(float x) {
  x = x + 0.0f;
  uint32_t u = __float_as_uint(x);
  return (u >> 31) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ float funkey(uint32_t k) {
  uint32_t u = (k >> 31) ? (k & 0x7FFFFFFFu) : ~k;
  return __uint_as_float(u);
}

// float -> double (exact).  The hardware conversion (F2F.F64.F32, XU pipe) is one
// instruction; when every visited arc was evaluated it saturated the XU pipe and an
// integer-pipe bit rebuild was faster (profiles/r1_v5_ncu_summary.txt).  With label
// lookups a tenth of the arcs is evaluated and the single instruction wins again
// (measured 108.8 -> 103.5 ms per launch).
__device__ __forceinline__ double widen(float f) { return static_cast<double>(f); }

// L2 eviction policy of the graph.  The recombination tables are scattered over megabytes
// per lane and every entry is touched a handful of times within one frame; left alone their
// lines push the graph out of L2 every couple of frames.  Graph loads ask to be evicted last
// (createpolicy lands in a uniform register).  Measured on the bench workload: 71.6 -> 70.7
// ms per step.  (The converse -- evict-first on the table's loads and wipes -- is a loss:
// an entry's probe, CAS, commit read and wipe do find each other in L2, and evict-first
// makes them miss: recombination 40 k -> 49 k cycles per frame.  ptxas does not take a
// cache policy on atom.cas.)
__device__ __forceinline__ unsigned long long l2_evict_last() {
  unsigned long long p;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// ---- recombination-table accesses
struct EntryWords {  // one entry as loaded by a 256-bit load
  HVal val;
  int32_t key;
  uint32_t idx;
  uint32_t epoch;
};

__device__ __forceinline__ EntryWords ld_entry(const Entry *e) {
  unsigned long long a, b, c, d;
  asm volatile("ld.global.cg.v4.b64 {%0, %1, %2, %3}, [%4];"
               : "=l"(a), "=l"(b), "=l"(c), "=l"(d)
               : "l"(e)
               : "memory");
  EntryWords w;
  w.val.cost = a;
  w.val.arg = b;
  w.key = static_cast<int32_t>(c);
  w.idx = static_cast<uint32_t>(c >> 32);
  w.epoch = static_cast<uint32_t>(d);
  return w;
}

__device__ __forceinline__ void st_entry(Entry *e, HVal v, int32_t key, uint32_t idx,
                                         uint32_t epoch) {
  const unsigned long long c =
      static_cast<unsigned long long>(static_cast<uint32_t>(key)) |
      (static_cast<unsigned long long>(idx) << 32);
  const unsigned long long d = epoch;
  asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(e), "l"(v.cost), "l"(v.arg),
               "l"(c), "l"(d)
               : "memory");
}

__device__ __forceinline__ HVal ld_hval(const HVal *p) {
  HVal r;
  ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2 *>(p));
  r.cost = v.x;
  r.arg = v.y;
  return r;
}

// ---- graph loads (read-only path)
__device__ __forceinline__ int4 gld(const int4 *p) {
#if KD_OPT_GRAPH_EL
  int4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.s32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p), "l"(l2_evict_last()));
  return v;
#else
  return __ldg(p);
#endif
}
__device__ __forceinline__ int2 gld(const int2 *p) {
#if KD_OPT_GRAPH_EL
  int2 v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.s32 {%0, %1}, [%2], %3;"
               : "=r"(v.x), "=r"(v.y)
               : "l"(p), "l"(l2_evict_last()));
  return v;
#else
  return __ldg(p);
#endif
}

// a 32-byte state record in one 256-bit load (one L1 tag lookup instead of two)
__device__ __forceinline__ void gld_state(const int4 *p, int4 *a, int4 *b) {
#if KD_OPT_ST256 && KD_OPT_GRAPH_EL
  asm volatile("ld.global.nc.L2::cache_hint.v8.s32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8], %9;"
               : "=r"(a->x), "=r"(a->y), "=r"(a->z), "=r"(a->w), "=r"(b->x), "=r"(b->y), "=r"(b->z),
                 "=r"(b->w)
               : "l"(p), "l"(l2_evict_last()));
#else
  *a = gld(p);
  *b = gld(p + 1);
#endif
}

// 128-bit compare-and-swap (ATOMG.E.CAS.128); returns the previous value.
__device__ __forceinline__ HVal cas_hval(HVal *p, HVal cmp, HVal val) {
  HVal out;
  asm volatile(
      "{\n\t"
      ".reg .b128 c, v, o;\n\t"
      "mov.b128 c, {%2, %3};\n\t"
      "mov.b128 v, {%4, %5};\n\t"
      "atom.global.relaxed.gpu.cas.b128 o, [%6], c, v;\n\t"
      "mov.b128 {%0, %1}, o;\n\t"
      "}"
      : "=l"(out.cost), "=l"(out.arg)
      : "l"(cmp.cost), "l"(cmp.arg), "l"(val.cost), "l"(val.arg), "l"(p)
      : "memory");
  return out;
}

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t *total) {
  const int lane = threadIdx.x & 31;
  uint32_t s = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xFFFFFFFFu, s, o);
    if (lane >= o) s += t;
  }
  *total = __shfl_sync(0xFFFFFFFFu, s, 31);
  return s - v;
}


// ---- bulk asynchronous copy (TMA, cp.async.bulk) of one log-prob row into shared memory,
// completing on an mbarrier.  The copy of frame t+1's row is issued by one thread as soon as
// frame t's scan is done and lands while the frame's recombination, closure and commit run,
// so the next frame does not start with a DRAM round trip.
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// One thread: expect `bytes` on the barrier and start the copy (16-byte aligned, multiple of 16).
__device__ __forceinline__ void bulk_load_row(float *dst_smem, const float *src_gmem,
                                              uint32_t bytes, unsigned long long *bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "KD_MBAR_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra KD_MBAR_DONE;\n\t"
      "bra KD_MBAR_WAIT;\n\t"
      "KD_MBAR_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// A row can go through the bulk-copy engine iff its address and size are multiples of 16.
__device__ __forceinline__ bool row_bulk_ok(const float *row_g, int cols) {
  return ((reinterpret_cast<uintptr_t>(row_g) | (static_cast<uintptr_t>(cols) * 4u)) & 15u) == 0;
}

struct Shared {
  double red_d[32];
  int red_i[32];
  unsigned long long red_k[2][16];  // block_min: per-warp minima, two buffers used alternately
  uint32_t hist[256];
  uint32_t warp_sums[32];
  uint32_t ex_coarse[33];  // t_ex[32 * q]: the top level of the item -> token search, bank by bank
  // This frame's region of the lane's table: a power-of-two prefix sized from the frame's
  // candidates (the table is empty between frames, so every frame may choose anew).
  uint32_t hmask;
  int32_t hshift;
  uint32_t region_fail;    // the region filled up: the frame is searched again with the whole table
  long long t_mark;
  uint32_t cut_fkey;  // running next-frame cutoff, rounded UP to float (a filter only)
  uint32_t acc_emit, acc_eps, acc_expanded;  // per-frame counters
  uint32_t list_n;
  uint32_t cand_n;
  uint32_t q_n[2];
  double wc;        // this frame's weight_cutoff
  uint32_t n_dead;  // commit: slot-list entries that are not tokens
  uint32_t n_front; // commit: tokens below good_cut
  uint32_t n_mid;   // commit: tokens below mid_cut
  uint32_t count;
  uint32_t sel_bin, sel_k;
  // the frame's labels ordered by bucket of their acoustic cost (-log-prob - minimum):
  // labels with cost below a bound are lab_order[0 .. bin_start[bucket(bound) + 1])
  float ac_min;
  int order_ok;
  uint32_t acc_items;
  int status;
  int item;
  int32_t rows_ready;
  uint32_t park_n;         // arrivals parked for the second recombination pass
  int load_first;          // table arrivals read the entry before they try to claim it (see table_arrive)
  int any_final;           // best-path selection scratch
  uint32_t best_tok;
  int row_pending;         // a bulk copy of the next frame's row is in flight
  uint32_t row_parity;     // phase of row_bar the next wait observes
  unsigned long long row_bar;  // mbarrier the bulk copy of a log-prob row completes on
  int yield;  // the lane ran out of streamed rows and gives its CTA back
};

// min over the block of a double, value only: ordered 64-bit keys, the warp minimum by
// two 32-bit REDUX.MIN (high words, then low words among the lanes holding the minimal
// high word), one barrier, every thread folds the per-warp minima itself.  `phase` (0/1)
// selects the buffer; consecutive calls alternate, which makes the single barrier enough
// (a buffer is rewritten two calls later, behind the barrier of the call in between).
template <int THREADS>
__device__ __forceinline__ double block_min(double v, Shared &sh, int phase) {
  constexpr int NW = THREADS / 32;
  const unsigned long long k = dkey(v);
  const uint32_t hi = static_cast<uint32_t>(k >> 32);
  const uint32_t mh = __reduce_min_sync(0xFFFFFFFFu, hi);
  const uint32_t ml =
      __reduce_min_sync(0xFFFFFFFFu, hi == mh ? static_cast<uint32_t>(k) : 0xFFFFFFFFu);
  if ((threadIdx.x & 31) == 0)
    sh.red_k[phase][threadIdx.x >> 5] = (static_cast<unsigned long long>(mh) << 32) | ml;
  __syncthreads();
  unsigned long long m = sh.red_k[phase][0];
#pragma unroll
  for (int w = 1; w < NW; ++w) {
    const unsigned long long o = sh.red_k[phase][w];
    m = o < m ? o : m;
  }
  return dunkey(m);
}

// min over the block of (v, idx); ties -> lowest idx.  All threads get it.
template <int THREADS>
__device__ __forceinline__ void block_min_arg(double v, int idx, Shared &sh, double *out_v,
                                              int *out_i) {
  constexpr int NW = THREADS / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    double ov = __shfl_xor_sync(0xFFFFFFFFu, v, o);
    int oi = __shfl_xor_sync(0xFFFFFFFFu, idx, o);
    if (ov < v || (ov == v && static_cast<uint32_t>(oi) < static_cast<uint32_t>(idx))) {
      v = ov;
      idx = oi;
    }
  }
  if (lane == 0) {
    sh.red_d[warp] = v;
    sh.red_i[warp] = idx;
  }
  __syncthreads();
  if (warp == 0) {
    v = lane < NW ? sh.red_d[lane] : __longlong_as_double(0x7FF0000000000000ll);
    idx = lane < NW ? sh.red_i[lane] : -1;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      double ov = __shfl_xor_sync(0xFFFFFFFFu, v, o);
      int oi = __shfl_xor_sync(0xFFFFFFFFu, idx, o);
      if (ov < v || (ov == v && static_cast<uint32_t>(oi) < static_cast<uint32_t>(idx))) {
        v = ov;
        idx = oi;
      }
    }
    if (lane == 0) {
      sh.red_d[0] = v;
      sh.red_i[0] = idx;
    }
  }
  __syncthreads();
  *out_v = sh.red_d[0];
  *out_i = sh.red_i[0];
  __syncthreads();
}

// Per-lane views of the global buffers.
struct LaneBuf {
  double *a_cost;
  unsigned long long *a_link;
  int32_t *a_state;
  Entry *table;
  uint32_t *bitmap;
  uint32_t *list;
  uint2 *queue;
  uint4 *cand;
  uint4 *front;
  uint32_t *s_list;  // shared memory: the first kListSmem entries of the slot list
};

__device__ __forceinline__ LaneBuf lane_buffers(const Params &P, int lane) {
  LaneBuf b;
  size_t L = static_cast<size_t>(lane);
  b.a_cost = P.a_cost + L * P.arena_cap;
  b.a_link = P.a_link + L * P.arena_cap;
  b.a_state = P.a_state + L * P.arena_cap;
  b.table = P.table + L * P.hcap;
  b.bitmap = P.bitmap + L * (P.hcap >> 5);
  b.list = P.list + L * P.lcap;
  b.queue = P.queue + L * 2 * P.qcap;
  b.cand = P.cand + L * P.ccap;
  b.front = P.front + L * kFrontCap;
  return b;
}

__device__ __forceinline__ uint32_t table_hash(const Shared &sh, int32_t state) {
  // groups of 4 consecutive states share a 128-byte line; groups are scattered
  return ((((static_cast<uint32_t>(state) >> 2) * 0x9E3779B1u) >> sh.hshift) << 2) |
         (static_cast<uint32_t>(state) & 3u);
}

// The whole table for the frame to come (the default; lane_expand_emitting may shrink it).
__device__ __forceinline__ void table_region_reset(const Params &P, Shared &sh) {
  sh.hmask = P.hmask;
  sh.hshift = P.hshift;
  sh.region_fail = 0;
}

// A slot that was just claimed joins this frame's slot list (its position there is the
// token's number in the next block: the return value) and, when `eps_queue` is given (the
// state has epsilon arcs), that queue: the closure only visits those.
__device__ __forceinline__ uint32_t register_claim(const Params &P, const LaneBuf &B, Shared &sh,
                                                   uint32_t h, int32_t state, uint2 *eps_queue,
                                                   uint32_t *eps_queue_n) {
  const uint32_t pos = atomicAdd(&sh.list_n, 1u);
  // (a region holds at most lcap entries unless it is the whole table: every claim made in a
  // region is in the list, which is what the roll-back of a failed region walks)
  if (pos < P.lcap) {
    if (pos < static_cast<uint32_t>(kListSmem)) {
      B.s_list[pos] = h;
    } else {
      B.list[pos] = h;
    }
    // half full: the whole table is the hard limit it always was, a region just gives up
    if (pos > (sh.hmask >> 1) && sh.hmask != P.hmask) sh.region_fail = 1;
  } else {
    atomicOr(&sh.status, kStatusHashOverflow);
  }
  if (eps_queue != nullptr) {
    const uint32_t qp = atomicAdd(eps_queue_n, 1u);
    if (qp < P.qcap) {
      eps_queue[qp] = make_uint2(h, static_cast<uint32_t>(state));
    } else {
      atomicOr(&sh.status, kStatusQueueOverflow);
    }
  }
  return pos;
}

// The arrival `mine` at `state`.  Returns the state's slot, kNoIdx on overflow, or kRetry.
//   *owner = true : the state was not in the table; this thread took an entry for it and has
//                   written it, value included -- nothing else to do;
//   *owner = false: the state has an entry; *cur is its value as just read (the caller
//                   recombines with a CAS on Entry::val).
// Probing is linear over the bitmap: the first entry of the probe sequence whose bit this
// thread flips is its own; an entry whose bit was already set belongs to the state it names.
// An entry can be claimed (bit set) and not written yet: its owner is between the atomicOr
// and the store.  With `may_wait` false such an entry makes the call return kRetry and the
// caller comes back after a barrier (the epsilon closure: the source token is expanded again
// in the next sweep).  With `may_wait` true the thread re-reads until the store has landed --
// nanoseconds; if the owner is a thread of the same warp on the other side of the divergent
// branch this relies on independent thread scheduling (sm_70+) to let it run, and the wait is
// bounded: it ends in a loud table-overflow status, never in a hang.  The two-pass
// recombination below keeps the bulk of the arrivals away from this case.
__device__ __forceinline__ uint32_t table_arrive(const Params &P, const LaneBuf &B, Shared &sh,
                                                 uint32_t epoch, int32_t state, HVal mine,
                                                 uint2 *eps_queue, uint32_t *eps_queue_n,
                                                 bool may_wait, bool *owner, HVal *cur,
                                                 uint32_t h_start = kNoIdx) {
  uint32_t h = h_start == kNoIdx ? table_hash(sh, state) : h_start;
  // Frames in which most arrivals meet a state that is already there (H-like graphs: every
  // state is reached through hundreds of arcs) look at the entry first: a valid entry of
  // this state saves the atomic.  The bitmap stays the authority for everything else.
  if (sh.load_first != 0 && h_start == kNoIdx) {
    const EntryWords w = ld_entry(B.table + h);
    if (w.epoch == epoch && w.key == state) {
      *owner = false;
      *cur = w.val;
      return h;
    }
  }
  for (uint32_t probe = 0; probe <= sh.hmask; ++probe) {
    // (a region that has given up is not probed to the bitter end: the frame is redone)
    if (probe >= 64u && *reinterpret_cast<volatile uint32_t *>(&sh.region_fail) != 0) return kNoIdx;
    const uint32_t bit = 1u << (h & 31u);
    const uint32_t old = atomicOr(B.bitmap + (h >> 5), bit);
    if ((old & bit) == 0) {
      const uint32_t pos = register_claim(P, B, sh, h, state, eps_queue, eps_queue_n);
      st_entry(B.table + h, mine, state, pos, epoch);
      *owner = true;
      return h;
    }
    // in use: by whom?  (its owner's store may still be on its way: the epoch tells)
    EntryWords w = ld_entry(B.table + h);
    if (w.epoch != epoch) {
      if (!may_wait) return kRetry;
      for (uint32_t spins = 0; w.epoch != epoch; ++spins) {
        if (spins > (1u << 22)) {  // never in a correct run: fail instead of hanging
          atomicOr(&sh.status, kStatusHashOverflow);
          return kNoIdx;
        }
        __nanosleep(64);
        w = ld_entry(B.table + h);
      }
    }
    if (w.key == state) {
      *owner = false;
      *cur = w.val;
      return h;
    }
    h = (h + 1) & sh.hmask;
  }
  if (sh.hmask != P.hmask) {
    sh.region_fail = 1;
  } else {
    atomicOr(&sh.status, kStatusHashOverflow);
  }
  return kNoIdx;
}

// Block-wide count of tokens with float(cost) <= bound (used to skip the exact
// order statistic when min_active cannot bind).
template <int THREADS>
__device__ uint32_t block_count_le(const double *cost, int n, double bound, Shared &sh) {
  if (threadIdx.x == 0) sh.count = 0;
  __syncthreads();
  uint32_t c = 0;
#pragma unroll 1
  for (int i = threadIdx.x; i < n; i += THREADS)
    c += (static_cast<double>(static_cast<float>(cost[i])) <= bound) ? 1u : 0u;
  c = __reduce_add_sync(0xFFFFFFFFu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&sh.count, c);
  __syncthreads();
  const uint32_t r = sh.count;
  __syncthreads();
  return r;
}

// GetCutoff's order statistic: the k-th smallest (0-based) of float(cost[i]).
template <int THREADS>
__device__ __noinline__ float select_kth(const double *cost, int n, uint32_t k, Shared &sh) {
  uint32_t prefix = 0, mask = 0;
  for (int pass = 3; pass >= 0; --pass) {
    const int shift = pass * 8;
    for (int i = threadIdx.x; i < 256; i += THREADS) sh.hist[i] = 0;
    __syncthreads();
#pragma unroll 1
    for (int i = threadIdx.x; i < n; i += THREADS) {
      uint32_t u = fkey(static_cast<float>(cost[i]));
      if ((u & mask) == prefix) atomicAdd(&sh.hist[(u >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t loc[8], s = 0;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        loc[b] = sh.hist[threadIdx.x * 8 + b];
        s += loc[b];
      }
      uint32_t tot;
      uint32_t ex = warp_excl_scan(s, &tot);
      if (k >= ex && k < ex + s) {
        uint32_t c = ex;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          if (k >= c && k < c + loc[b]) {
            sh.sel_bin = threadIdx.x * 8 + b;
            sh.sel_k = k - c;
          }
          c += loc[b];
        }
      }
    }
    __syncthreads();
    prefix |= sh.sel_bin << shift;
    mask |= 255u << shift;
    k = sh.sel_k;
    __syncthreads();
  }
  return funkey(prefix);
}

// faster-decoder.cc:244-336.  n records holding n_live tokens (holes cost +inf, so they
// sort behind every token), best = min cost.  All threads return
// the same (weight_cutoff, adaptive_beam).
template <int THREADS>
__device__ void lane_cutoff(const Params &P, const double *cost, int n, const LaneState &ls,
                            Shared &sh, double *weight_cutoff, float *adaptive_beam) {
  const int n_live = ls.n_live;
  const double best = ls.best_cost;
  const double inf = __longlong_as_double(0x7FF0000000000000ll);
  if (P.max_active == 0x7FFFFFFF && P.min_active == 0) {
    *adaptive_beam = P.beam;
    *weight_cutoff = best + static_cast<double>(P.beam);
    return;
  }
  const double beam_cutoff = best + static_cast<double>(P.beam);
  double max_cut = inf, min_cut = inf;
  if (n_live > P.max_active)
    max_cut = static_cast<double>(select_kth<THREADS>(cost, n, P.max_active, sh));
  if (max_cut < beam_cutoff) {
    *adaptive_beam = static_cast<float>(max_cut - best + static_cast<double>(P.beam_delta));
    *weight_cutoff = max_cut;
    return;
  }
  if (n_live > P.min_active) {
    if (P.min_active == 0) {
      min_cut = best;
    } else {
      // The (min_active+1)-th smallest exceeds beam_cutoff iff at most
      // min_active values are <= beam_cutoff: one counting pass decides
      // whether the exact order statistic is needed at all.
      // The commit already counted the tokens below mid_cut: when that bound lies safely
      // (float rounding of the costs) below beam_cutoff they all count, and no pass is needed.
      uint32_t c;
      if (ls.n_mid > P.min_active && ls.mid_cut <= beam_cutoff - 1e-5 * (fabs(beam_cutoff) + 1.0)) {
        c = static_cast<uint32_t>(ls.n_mid);
      } else {
        c = block_count_le<THREADS>(cost, n, beam_cutoff, sh);
      }
      if (c > static_cast<uint32_t>(P.min_active)) {
        min_cut = beam_cutoff;  // some value <= beam_cutoff: min_active does not bind
      } else {
        min_cut = static_cast<double>(select_kth<THREADS>(cost, n, P.min_active, sh));
      }
    }
  }
  if (min_cut > beam_cutoff) {
    *adaptive_beam = static_cast<float>(min_cut - best + static_cast<double>(P.beam_delta));
    *weight_cutoff = min_cut;
    return;
  }
  *adaptive_beam = P.beam;
  *weight_cutoff = beam_cutoff;
}

// One epsilon arrival (faster-decoder.cc:90-116): dropped if worse than the
// cutoff (done by the caller), inserted if the state is new, replaces the
// incumbent only if strictly better.  An incumbent left over from the emitting
// phase with cost >= C* is not a token (see file comment) and is overwritten.
// A token that was created or improved is queued for expansion when its state
// has epsilon arcs (flag in the arc's nextstate word).
template <bool SIMPLE>
__device__ __forceinline__ bool eps_arrival(const Params &P, const LaneBuf &B, Shared &sh,
                                            uint32_t epoch, uint32_t dst_word,
                                            unsigned long long cost_key, uint32_t arc,
                                            uint32_t src_number, unsigned long long cstar_key,
                                            uint2 *q_next, uint32_t *q_next_n) {
  // (returns false if the destination's entry is claimed but not written yet: the caller
  // expands the source token again in the next sweep)
  const int32_t state = static_cast<int32_t>(dst_word & ~kEpsFlag);
  HVal mine;
  mine.cost = cost_key;
  mine.arg = (static_cast<unsigned long long>(arc | kEpsFlag) << 32) | src_number;
  bool owner;
  HVal cur;
  const uint32_t h =
      table_arrive(P, B, sh, epoch, state, mine, nullptr, nullptr, false, &owner, &cur);
  if (h == kRetry) return false;
  if (h == kNoIdx) return true;
  if (!owner) {
    bool tie_only;
    while (true) {
      const bool cur_is_eps = (cur.arg >> 63) != 0;
      // Two epsilon arrivals with bit-equal cost: the lower epsilon-arc index wins, so the
      // backpointer does not depend on thread scheduling (the reference keeps whichever came
      // first in its LIFO order, faster-decoder.cc:107-112; an incumbent from the emitting
      // phase stays on a tie in both).
      tie_only = cur_is_eps && mine.cost == cur.cost && mine.arg < cur.arg;
      // (SimpleDecoder search: every table entry is a token, simple-decoder.cc:224-231)
      const bool replace = mine.cost < cur.cost || tie_only ||
                           (!SIMPLE && !cur_is_eps && !(cur.cost < cstar_key));
      if (!replace) return true;
      HVal got = cas_hval(&B.table[h].val, cur, mine);
      if (got.cost == cur.cost && got.arg == cur.arg) break;
      cur = got;
    }
    if (tie_only) return true;  // same cost: nothing new to expand
  }
  if (dst_word & kEpsFlag) {
    const uint32_t pos = atomicAdd(q_next_n, 1u);
    if (pos < P.qcap) {
      q_next[pos] = make_uint2(h, static_cast<uint32_t>(state));
    } else {
      atomicOr(&sh.status, kStatusQueueOverflow);
    }
  }
  return true;
}

// Expands the epsilon arcs of the token in table slot `slot`
// (faster-decoder.cc:71-117).
template <bool SIMPLE>
__device__ __forceinline__ void expand_eps(const Params &P, const LaneBuf &B, Shared &sh,
                                           uint32_t epoch, uint2 entry,
                                           unsigned long long cstar_key, double cstar,
                                           uint2 *q_next, uint32_t *q_next_n,
                                           uint32_t *eps_count) {
  const uint32_t slot = entry.x;
  // the worklist entry names the state: its record is requested together with the token
  // (one round trip instead of two)
#if KD_OPT_QSTATE
  const int4 st = gld(P.st + 2 * static_cast<size_t>(entry.y));
#endif
  // value, state and number of the token: one sector, one 256-bit load
  const EntryWords w = ld_entry(B.table + slot);
  const HVal v = w.val;
  const bool is_eps = (v.arg >> 63) != 0;
  // a token iff cost < C*, or it came from an epsilon arc (then cost <= C*)
  if (!(SIMPLE || v.cost < cstar_key || is_eps)) return;
#if !KD_OPT_QSTATE
  const int4 st = gld(P.st + 2 * static_cast<size_t>(w.key));
#endif
  if (st.w == 0) return;
  const double cost = dunkey(v.cost);
  *eps_count += static_cast<uint32_t>(st.w);
  bool again = false;
  for (int a = st.z; a < st.z + st.w; ++a) {
    const int4 arc = gld(P.n_arc + a);
    const double nc = cost + widen(__int_as_float(arc.y));
    if (nc > cstar) continue;  // faster-decoder.cc:92
    if (!eps_arrival<SIMPLE>(P, B, sh, epoch, static_cast<uint32_t>(arc.z), dkey(nc),
                             static_cast<uint32_t>(a), w.idx, cstar_key, q_next, q_next_n))
      again = true;
  }
  if (again) {
    // some destination was being written by another thread: this token is expanded again in
    // the next sweep (arrivals that already went through repeat with equal cost: no effect)
    const uint32_t pos = atomicAdd(q_next_n, 1u);
    if (pos < P.qcap) {
      q_next[pos] = entry;
    } else {
      atomicOr(&sh.status, kStatusQueueOverflow);
    }
  }
}

// Garbage collection of the backpointer arena (the reference frees a token when the last
// token pointing at it dies, faster-decoder.h:145-155, so its memory follows the live
// history; here the arena is compacted when a frame's tokens no longer fit).  The records
// below `tok_base` are kept iff a token of the current block reaches them:
//   1. mark   every token of the block walks its backpointers until it meets a marked record
//             (marks live in a_state of the old records, which nothing reads any more);
//   2. scan   marked records get their new, order-preserving index (left in a_state);
//   3. move   old records slide down tile by tile (all reads of a tile before its writes;
//             targets never lie above their sources), links rewritten through the new indices;
//   4. block  the current block's links are rewritten in place, then it slides down behind
//             the kept records.
// Returns how far the current block moved (the caller still holds its old indices).  Cold
// path: once per several thousand frames of a long stream.
template <int THREADS>
__device__ __noinline__ uint32_t lane_compact_arena(const LaneBuf &B, Shared &sh, LaneState &ls) {
  constexpr int32_t kMark = -2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = THREADS / 32;
  const uint32_t base = ls.tok_base;
  const uint32_t n = static_cast<uint32_t>(ls.n_tok);
  // ---- 1. mark
  for (uint32_t i = tid; i < n; i += THREADS) {
    uint32_t p = static_cast<uint32_t>(B.a_link[base + i]);
    while (p != kNoPrev && p < base) {
      if (atomicExch(&B.a_state[p], kMark) == kMark) break;
      p = static_cast<uint32_t>(B.a_link[p]);
    }
  }
  __syncthreads();
  // ---- 2. scan: a_state[i] = new index of a marked record, -1 otherwise
  uint32_t running = 0;  // kept records below the current tile (uniform)
  for (uint32_t t0 = 0; t0 < base; t0 += THREADS) {
    const uint32_t i = t0 + tid;
    const uint32_t keep = (i < base && B.a_state[i] == kMark) ? 1u : 0u;
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, keep != 0);
    if (lane == 0) sh.warp_sums[warp] = __popc(bal);
    __syncthreads();
    uint32_t ex = running + __popc(bal & ((1u << lane) - 1u));
    uint32_t tile = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const uint32_t c = sh.warp_sums[w];
      if (w < warp) ex += c;
      tile += c;
    }
    if (i < base) B.a_state[i] = keep ? static_cast<int32_t>(ex) : -1;
    running += tile;
    __syncthreads();
  }
  const uint32_t kept = running;
  // ---- 3. move the kept old records
  for (uint32_t t0 = 0; t0 < base; t0 += THREADS) {
    const uint32_t i = t0 + tid;
    int32_t dst = -1;
    double c = 0.0;
    unsigned long long link = 0;
    if (i < base) {
      dst = B.a_state[i];
      if (dst >= 0) {
        c = B.a_cost[i];
        link = B.a_link[i];
        const uint32_t p = static_cast<uint32_t>(link);
        if (p != kNoPrev)
          link = (link & 0xFFFFFFFF00000000ull) | static_cast<uint32_t>(B.a_state[p]);
      }
    }
    __syncthreads();
    if (dst >= 0) {
      B.a_cost[dst] = c;
      B.a_link[dst] = link;
    }
    __syncthreads();
  }
  // ---- 4. the current block: links in place, then down to `kept`
  for (uint32_t i = tid; i < n; i += THREADS) {
    unsigned long long link = B.a_link[base + i];
    const uint32_t p = static_cast<uint32_t>(link);
    if (p != kNoPrev) {
      const uint32_t np = p >= base ? p - base + kept : static_cast<uint32_t>(B.a_state[p]);
      B.a_link[base + i] = (link & 0xFFFFFFFF00000000ull) | np;
    }
  }
  __syncthreads();
  if (kept < base) {
    for (uint32_t t0 = 0; t0 < n; t0 += THREADS) {
      const uint32_t i = t0 + tid;
      double c = 0.0;
      unsigned long long link = 0;
      int32_t st = -1;
      if (i < n) {
        c = B.a_cost[base + i];
        link = B.a_link[base + i];
        st = B.a_state[base + i];
      }
      __syncthreads();
      if (i < n) {
        B.a_cost[kept + i] = c;
        B.a_link[kept + i] = link;
        B.a_state[kept + i] = st;
      }
      __syncthreads();
    }
  }
  if (tid == 0) {
    ls.tok_base = kept;
    ls.arena_used = kept + n;
    ls.st_compactions += 1;
  }
  __syncthreads();
  return base - kept;
}

// Epsilon closure (faster-decoder.cc:59-119) followed by the commit of the
// frame: live table entries become the next token block in the arena, the
// table is wiped, and the block's min cost is recorded for the next GetCutoff.
// On entry queue 0 holds the slots, claimed during the emitting phase (or by
// InitDecoding), whose states have epsilon arcs.
//
// good_cut is handed to the next frame's scan, which takes the tokens below it
// first: its running cutoff tightens early and few arcs that the exact cutoff
// rejects become candidates.
// Returns true if the frame has to be searched again: its table region filled up (the
// closure can multiply the tokens of a frame many times over at word boundaries, which the
// candidate count does not show); the table is back to empty then and nothing was committed.
template <int THREADS, bool SIMPLE>
__device__ bool lane_closure_and_commit(const Params &P, const LaneBuf &B, Shared &sh,
                                        LaneState &ls, double cstar, double good_cut,
                                        double mid_cut) {
  const int tid = threadIdx.x;
  const unsigned long long cstar_key = dkey(cstar);
  // good_cut <= C*: every entry below it is a token
  const unsigned long long good_key = dkey(fmin(good_cut, cstar));
  const unsigned long long mid_key = dkey(fmin(mid_cut, cstar));
  const double inf = __longlong_as_double(0x7FF0000000000000ll);
  const long long t_begin = clock64();
  if (tid == 0) {
    sh.q_n[1] = 0;
    sh.n_dead = 0;
    sh.n_front = 0;
    sh.n_mid = 0;
    sh.acc_eps = 0;
  }
  __syncthreads();
  // ---- closure: sweep 0 expands the candidates, later sweeps the improved ones
  uint32_t eps_count = 0;
  uint2 *q0 = B.queue, *q1 = B.queue + P.qcap;
  int cur = 0;
  long long sweeps = 0;
  while (true) {
    const uint32_t qn = min(sh.q_n[cur], P.qcap);
    if (qn == 0 || sh.status != 0 || sh.region_fail != 0) break;
    __syncthreads();
    if (tid == 0) sh.q_n[cur ^ 1] = 0;
    __syncthreads();
    uint2 *qc = cur ? q1 : q0, *qx = cur ? q0 : q1;
    for (uint32_t p = tid; p < qn; p += THREADS)
      expand_eps<SIMPLE>(P, B, sh, ls.epoch, qc[p], cstar_key, cstar, qx, &sh.q_n[cur ^ 1], &eps_count);
    __syncthreads();
    cur ^= 1;
    ++sweeps;
  }
  __syncthreads();
  if (sh.region_fail != 0) {
    // Roll the frame back: release every entry claimed (all of them are in the slot list), let
    // their contents go stale (new epoch), forget the worklists.
    const uint32_t claimed = min(sh.list_n, P.lcap);
    for (uint32_t p = tid; p < claimed; p += THREADS) {
      const uint32_t h = p < static_cast<uint32_t>(kListSmem) ? B.s_list[p] : B.list[p];
      B.bitmap[h >> 5] = 0u;
    }
    __syncthreads();
    if (tid == 0) {
      ls.epoch = ls.epoch + 1u == 0xFFFFFFFFu ? 0u : ls.epoch + 1u;
      ls.st_redo += 1u;
      ls.cyc_closure += clock64() - t_begin;
      sh.list_n = 0;
      sh.q_n[0] = 0;
      sh.q_n[1] = 0;
    }
    __syncthreads();
    return true;
  }
  const long long t_mid = clock64();
  // ---- commit: one pass over the slot list.  Tokens are numbered in claim order
  // (Entry::idx, written when the slot was claimed), so the predecessor number that
  // an epsilon arrival carries is final and nothing has to be counted first.  A
  // thread wipes (key, val) of its own entries after reading them.
  const uint32_t m = min(sh.list_n, P.lcap);
  // The block does not fit behind the records in use: drop the records no live token
  // reaches any more.  The table's emitting arrivals still name their predecessors by the
  // old indices of the current block, which moved down by `moved`.
  uint32_t moved = 0;
  if (static_cast<long long>(ls.arena_used) + m > P.arena_cap && ls.tok_base > 0)
    moved = lane_compact_arena<THREADS>(B, sh, ls);
  const uint32_t new_base = ls.arena_used;
  if (static_cast<long long>(new_base) + m > P.arena_cap) {
    if (tid == 0) sh.status |= kStatusArenaOverflow;
  }
  __syncthreads();
  const bool write_ok = (sh.status & kStatusArenaOverflow) == 0;
  double my_min = inf;
  int my_arg = -1;
  int32_t my_state = -1;
  uint32_t dead = 0, below_mid = 0;
  for (uint32_t p0 = 0; p0 < m; p0 += THREADS * 4) {
    uint32_t h[4];
    HVal v[4];
    int32_t key[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t p = p0 + u * THREADS + tid;
      h[u] = p < m ? (p < static_cast<uint32_t>(kListSmem) ? B.s_list[p] : B.list[p]) : kNoIdx;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v[u].cost = kEmptyCost;
      v[u].arg = kEmptyArg;
      key[u] = kEmptyKey;
      if (h[u] != kNoIdx) {
        const EntryWords w = ld_entry(B.table + h[u]);
        v[u] = w.val;
        key[u] = w.key;
      }
    }
    // the tokens close to the best also go to the front list (warp-aggregated append)
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool good = v[u].cost < good_key;  // an empty value compares above every key
      const uint32_t gmask = __ballot_sync(0xFFFFFFFFu, good);
      if (gmask == 0) continue;  // warp-uniform
      uint32_t gbase = 0;
      if ((tid & 31) == 0) gbase = atomicAdd(&sh.n_front, __popc(gmask));
      gbase = __shfl_sync(0xFFFFFFFFu, gbase, 0);
      if (good) {
        const uint32_t pos = gbase + __popc(gmask & ((1u << (tid & 31)) - 1u));
        if (pos < static_cast<uint32_t>(kFrontCap))
          B.front[pos] = make_uint4(static_cast<uint32_t>(v[u].cost),
                                    static_cast<uint32_t>(v[u].cost >> 32),
                                    static_cast<uint32_t>(key[u]), p0 + u * THREADS + tid);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (h[u] == kNoIdx) continue;
      const uint32_t idx = p0 + u * THREADS + tid;
      // Not a token: an arrival recombined while the candidate buffer was full (against
      // the running cutoff) that the exact cutoff rejects.  Its record stays in the
      // block as a hole (state -1, cost +inf) that every reader skips.
      const bool live = v[u].cost != kEmptyCost &&
                        (SIMPLE || v[u].cost < cstar_key || (v[u].arg >> 63) != 0);
      if (!live) ++dead;
      if (v[u].cost < mid_key) ++below_mid;
      if (write_ok) {
        const uint32_t arc = static_cast<uint32_t>(v[u].arg >> 32);
        uint32_t prev = static_cast<uint32_t>(v[u].arg);
        if (arc & kEpsFlag) {
          prev += new_base;  // predecessor is a token of this block
        } else if (arc != kNoArc) {
          prev -= moved;     // predecessor is a token of the previous block
        }
        const double c = live ? dunkey(v[u].cost) : inf;
        // written once, read once next frame (cost, state) or at traceback (link)
        __stcs(B.a_cost + new_base + idx, c);
        __stcs(B.a_link + new_base + idx, (static_cast<unsigned long long>(arc) << 32) | prev);
        __stcs(B.a_state + new_base + idx, live ? key[u] : -1);
        if (live && c < my_min) {
          my_min = c;
          my_arg = static_cast<int>(idx);
          my_state = key[u];
        }
      }
      // the entry is free again (its contents stay: the next owner overwrites them).  Every
      // entry in use is released by this pass, so the whole word can go: a plain store
      B.bitmap[h[u] >> 5] = 0u;
    }
  }
  double bmin;
  int barg;
  block_min_arg<THREADS>(my_min, my_arg, sh, &bmin, &barg);
  if (barg >= 0 && my_arg == barg) ls.best_state = my_state;  // one thread: token numbers are unique
  // accumulate counters
  eps_count = __reduce_add_sync(0xFFFFFFFFu, eps_count);
  if ((tid & 31) == 0 && eps_count) atomicAdd(&sh.acc_eps, eps_count);
  dead = __reduce_add_sync(0xFFFFFFFFu, dead);
  if ((tid & 31) == 0 && dead) atomicAdd(&sh.n_dead, dead);
  below_mid = __reduce_add_sync(0xFFFFFFFFu, below_mid);
  if ((tid & 31) == 0 && below_mid) atomicAdd(&sh.n_mid, below_mid);
  __syncthreads();
  if (tid == 0) {
    if (write_ok) {
      ls.tok_base = new_base;
      ls.n_tok = static_cast<int32_t>(m);
      ls.n_live = static_cast<int32_t>(m - sh.n_dead);
      ls.arena_used = new_base + m;
      ls.best_cost = bmin;
      ls.best_idx = barg;
      ls.good_cut = fmin(good_cut, cstar);
      ls.n_front = static_cast<int32_t>(sh.n_front);
      ls.mid_cut = fmin(mid_cut, cstar);
      ls.n_mid = static_cast<int32_t>(sh.n_mid);
    } else {
      ls.n_tok = 0;
      ls.n_live = 0;
      ls.best_cost = inf;
      ls.best_idx = -1;
      ls.best_state = -1;
    }
    ls.epoch = ls.epoch + 1u == 0xFFFFFFFFu ? 0u : ls.epoch + 1u;  // (0xFFFFFFFF: never-written entries)
    // the next frame resembles this one: more than two arrivals per state -> entry first
    sh.load_first = sh.cand_n > 2u * m ? 1 : 0;
    ls.st_sweeps += sweeps;
    ls.st_claimed += m;
    ls.st_eps_arcs += sh.acc_eps;
    ls.cyc_closure += t_mid - t_begin;
    ls.cyc_commit += clock64() - t_mid;
    sh.list_n = 0;
    sh.q_n[0] = 0;
  }
  __syncthreads();
  return false;
}

// Recombines one emitting arc that survived pruning at its destination state: the entry
// keeps the lexicographic minimum of (cost, arc index).
__device__ __forceinline__ void insert_arc(const Params &P, const LaneBuf &B, Shared &sh,
                                           uint32_t epoch, uint32_t a, unsigned long long nk,
                                           uint32_t tok_abs) {
  const int2 no = gld(P.e_no + a);
  HVal mine;
  mine.cost = nk;
  mine.arg = (static_cast<unsigned long long>(a) << 32) | tok_abs;
  bool owner;
  HVal cur;
  const uint32_t h = table_arrive(P, B, sh, epoch, no.x & 0x7FFFFFFF, mine,
                                  no.x < 0 ? B.queue : nullptr, &sh.q_n[0], true, &owner, &cur);
  if (h == kNoIdx || owner) return;
  while (mine.cost < cur.cost || (mine.cost == cur.cost && mine.arg < cur.arg)) {
    HVal got = cas_hval(&B.table[h].val, cur, mine);
    if (got.cost == cur.cost && got.arg == cur.arg) return;
    cur = got;
  }
}



// faster-decoder.cc:155-241 for one lane-frame.  Returns C*.
//
// Three steps: scan -> exact cutoff -> recombine.
//
// Scan.  Two passes over the token block (front list first, then the rest), the
// tokens taken a chunk (kTileTokens per thread) at a time.  Tokens that will be
// expanded (cost < weight_cutoff, at least one emitting arc) are compacted into
// shared memory together with the exclusive prefix sum of their work items -- all
// their emitting arcs, or just the frame's best labels looked up in the state's label
// table -- so the chunk's items form one flat index space.  The warps walk it one
// 32-item window at a time, round-robin.  The token owning each item of a window
// comes from a bit mask of the token boundaries falling inside the window (one
// shared-memory load, one warp OR-reduction, one popc).  new_weight = (w + cost) + ac.
// Most items fail the pruning test and cost nothing more.  The others are only
// *candidates*: the test is made against a running cutoff (the reference's
// next_weight_cutoff, faster-decoder.cc:172-217) kept in shared memory as a float
// rounded UP and lowered with one native 32-bit atomicMin; it is looser than the
// final cutoff.  Candidates are appended to a per-lane buffer, not recombined:
// recombining them on the spot makes a warp wait for one lane's table round trips,
// and fills the table (and L2) with arrivals that the final cutoff rejects -- 3x more
// slots claimed than tokens kept (profiles/r1_v5_ncu_summary.txt).
//
// Exact cutoff.  C* = min(new_weight) + adaptive_beam (faster-decoder.cc:240)
// from the per-thread fp64 minima: the arc with the globally smallest
// new_weight always passes the filter, so the minimum over candidates is the
// minimum over all arcs.
//
// Recombine.  All threads walk the candidate buffer; only candidates with
// new_weight < C* touch the table, so the table holds exactly the tokens.
template <int THREADS, bool ROW_SMEM, bool SIMPLE>
__device__ double lane_expand_emitting(const Params &P, const LaneBuf &B, Shared &sh,
                                       LaneState &ls, const float *row_g,
                                       const float *next_row_g, float *s_row,
                                       double *t_cost, uint32_t *t_ex, uint32_t *t_beg,
                                       int32_t *t_tab, uint32_t *t_tok, uint16_t *lab_order,
                                       uint16_t *bin_start, bool whole_table) {
  constexpr int KT = kTileTokens;
  constexpr int TT = THREADS * KT;
  constexpr int NW = THREADS / 32;
  // 32-ary search over up to TT compacted tokens: strides 1024, 32, 1 (or 32, 1)
  constexpr uint32_t kSearchTop = TT > 1024 ? 1024u : 32u;
  static_assert(TT <= 32768 && NW <= 16, "tile size out of range");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double inf = __longlong_as_double(0x7FF0000000000000ll);
  const int n = ls.n_tok;
  const uint32_t base = ls.tok_base;
  const double *cost = B.a_cost + base;
  const int32_t *state = B.a_state + base;
  const long long t_begin = clock64();

  // the log-prob row of this frame -> shared memory (decodable-ctc.cc:22-29),
  // already negated (faster-decoder.cc:209); widened to fp64 where it is used.
  // (Shared memory is kept small on purpose: what it does not take is L1.)
  // the best token's state record is needed for the seed below: requested first, it
  // arrives while the row is staged and the labels are ordered
  int4 seed_sa = make_int4(0, 0, 0, 0), seed_sb = make_int4(-1, 0, 0, 0);
  if (n > 0 && ls.best_state >= 0) {
    gld_state(P.st + 2 * static_cast<size_t>(ls.best_state), &seed_sa, &seed_sb);
  }
  // The row is kept as it comes (log-probs); every use negates it (faster-decoder.cc:209).
  // Usually it is already on its way: the previous frame started a bulk copy (TMA) of it
  // after its scan.  Rows that cannot go through the bulk-copy engine (address or size not
  // a multiple of 16 bytes) are loaded by the threads; streamed rows are written by the copy
  // engine while this kernel runs, so nothing here reads them through the non-coherent path.
  float amin = __int_as_float(0x7F800000);  // smallest acoustic cost of the frame (ROW_SMEM)
  bool sh_bulk_done = false;
  if (ROW_SMEM) {
    const bool bulk = sh.row_pending != 0 || row_bulk_ok(row_g, P.cols);  // uniform
    if (bulk) {
      if (sh.row_pending == 0 && tid == 0)
        bulk_load_row(s_row, row_g, static_cast<uint32_t>(P.cols) * 4u, &sh.row_bar);
      mbar_wait(&sh.row_bar, sh.row_parity);
    } else {
      for (int i = tid; i < P.cols; i += THREADS) s_row[i] = __ldcg(row_g + i);
    }
#pragma unroll 1
    for (int i = tid; i < P.cols; i += THREADS) amin = fminf(amin, -s_row[i]);
    if (bulk) sh_bulk_done = true;
  }
  if (tid == 0) {
    sh.cut_fkey = fkey(__int_as_float(0x7F800000));
    sh.acc_emit = sh.acc_expanded = sh.acc_items = 0;
    sh.cand_n = 0;
    sh.park_n = 0;
    table_region_reset(P, sh);  // (arcs recombined during the scan -- candidate buffer full -- see the whole table)
  }
  double wc;
  float abf;
  lane_cutoff<THREADS>(P, cost, n, ls, sh, &wc, &abf);
  // SimpleDecoder does not prune after InitDecoding: its first frame expands every token
  if (SIMPLE && ls.frames_decoded == 0) wc = inf;
  const double ab = static_cast<double>(abf);
  if (tid == 0) sh.wc = wc;
  __syncthreads();
  if (sh_bulk_done && tid == 0) {  // every thread is past the wait: the barrier's next phase
    sh.row_parity ^= 1u;
    sh.row_pending = 0;
  }
  // Order the frame's labels by acoustic cost (counting sort into 1/16-wide buckets
  // above the minimum).  A token whose slack admits few labels looks those labels up
  // in its state's label table instead of scanning all its arcs.
  if (ROW_SMEM && P.cols <= kMaxOrderCols && P.labtab != nullptr) {
    // the tile arrays (t_cost .. t_tok, contiguous, 6 * TT words) are free until the scan
    uint32_t *hist = reinterpret_cast<uint32_t *>(t_cost);
    static_assert(6 * TT >= kOrderBins, "tile arrays too small to hold the label histogram");
#pragma unroll 1
    for (int b = tid; b < kOrderBins; b += THREADS) hist[b] = 0;
    amin = static_cast<float>(block_min<THREADS>(static_cast<double>(amin), sh, 0));
    // (these loops touch shared memory only: not unrolled, the kernel is short of
    // instruction cache, not of latency hiding here)
#pragma unroll 1
    for (int i = tid; i < P.cols; i += THREADS) {
      const float d = (-s_row[i] - amin) * 16.0f;
      const int b = d < static_cast<float>(kOrderBins - 1) ? static_cast<int>(d) : kOrderBins - 1;
      atomicAdd(&hist[b], 1u);
    }
    __syncthreads();
    {
      // exclusive prefix over the buckets, kOrderBins / THREADS consecutive ones per thread
      constexpr int PER = (kOrderBins + THREADS - 1) / THREADS;
      uint32_t loc[PER], sum = 0;
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int b = tid * PER + k;
        loc[k] = b < kOrderBins ? hist[b] : 0;
        sum += loc[k];
      }
      uint32_t wtot;
      uint32_t ex = warp_excl_scan(sum, &wtot);
      if (lane == 0) sh.warp_sums[warp] = wtot;
      __syncthreads();
#pragma unroll
      for (int w = 0; w < NW; ++w)
        if (w < warp) ex += sh.warp_sums[w];
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int b = tid * PER + k;
        if (b < kOrderBins) {
          bin_start[b] = static_cast<uint16_t>(ex);
          hist[b] = ex;  // scatter cursor
          ex += loc[k];
        }
      }
      if (tid == 0) {
        bin_start[kOrderBins] = static_cast<uint16_t>(P.cols);
        bin_start[kOrderBins + 1] = static_cast<uint16_t>(P.cols);
        sh.ac_min = amin;
        sh.order_ok = 1;
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int i = tid; i < P.cols; i += THREADS) {
      const float d = (-s_row[i] - amin) * 16.0f;
      const int b = d < static_cast<float>(kOrderBins - 1) ? static_cast<int>(d) : kOrderBins - 1;
      lab_order[atomicAdd(&hist[b], 1u)] = static_cast<uint16_t>(i + 1);
    }
    __syncthreads();
  } else {
    if (tid == 0) sh.order_ok = 0;
    __syncthreads();
  }
  // Seed the running cutoff from the best token's arcs (faster-decoder.cc:176-189).  Any
  // subset of its arcs gives a valid (if looser) bound: a state with a label table
  // only looks up the THREADS labels with the best acoustic cost.
  double seed = inf;
  if (n > 0 && ls.best_cost < wc) {
    const int4 st = seed_sa;
    const int4 sb = seed_sb;
    if (sh.order_ok != 0 && sb.x >= 0) {
      if (tid < P.cols) {
        const uint32_t lab = lab_order[tid];
        const int2 ent = gld(P.labtab + static_cast<size_t>(sb.x) * P.lab_stride + (lab - 1));
        if (ent.y >= 0)
          seed = (widen(__int_as_float(ent.x)) + ls.best_cost) + widen(-s_row[lab - 1]);
      }
    } else {
#pragma unroll 1
      for (int a = tid; a < st.y; a += THREADS) {
        const int2 iw = gld(P.e_iw + st.x + a);
        const double ac = widen(-(ROW_SMEM ? s_row[iw.x - 1] : __ldcg(row_g + iw.x - 1)));
        const double nw = (widen(__int_as_float(iw.y)) + ls.best_cost) + ac;
        seed = fmin(seed, nw);
      }
    }
  }
  {
    const double smin = block_min<THREADS>(seed, sh, 1);
    if (tid == 0) sh.cut_fkey = fkey(__double2float_ru(smin + ab));
    __syncthreads();
  }
  if (tid == 0) sh.t_mark = clock64();

  // ---------------------------------------------------------------- scan
  double my_min = inf;
  // Two passes over the token block.  Pass 0 takes the tokens close to the best
  // (cost < good_cut, fixed by the previous commit): scanning them first makes the
  // running cutoff tight before the bulk of the tokens is classified (scan vs label
  // lookup) and filtered in pass 1.  Within a pass the tokens are read a chunk
  // (kTileTokens per thread) at a time; the ones to expand are compacted into the tile
  // arrays and worked off before the next chunk is classified.  Small chunks win:
  // later chunks are classified against a tighter cutoff, and the tile arrays take
  // little shared memory away from L1 (measured: 2 per thread beats 1, 3 and 4).
  // (Loop state is kept in shared memory where it can be: the item loop below needs
  // every register it can get.)
  static_assert(TT <= kFrontCap, "front list smaller than a scan tile");
  // (a block that fits one tile is scanned in one pass: the second pass would cost three
  // more dependent round trips and has little left to tighten)
  const bool single_pass = KD_OPT_SINGLE_PASS && ls.n_tok <= KD_SINGLE_PASS_TILES * TT;
  for (int pass = single_pass ? 1 : 0; pass < 2; ++pass) {
    // pass 0 reads the front list the commit wrote, unless it overflowed (then it filters the block)
    const bool front_list =
        pass == 0 && static_cast<uint32_t>(ls.n_front) <= static_cast<uint32_t>(kFrontCap);
    const uint32_t un = static_cast<uint32_t>(front_list ? ls.n_front : ls.n_tok);
    for (uint32_t tile0 = 0; tile0 < un; tile0 += TT) {
    const double wcut = sh.wc;
    const double good = single_pass ? -inf : fmin(ls.good_cut, wcut);
    const uint32_t tile_end = min(un, tile0 + TT);
    // chunk setup: 4 consecutive tokens per thread -> (cost, arc range or label
    // count) of the tokens to expand
    uint32_t cnt[KT], beg[KT], ti[KT];
    int32_t tab[KT];
    double tc[KT];
    uint32_t ch_expanded = 0, ch_arcs = 0;  // this chunk's share of the counters
    {
      int32_t ts[KT];
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        const uint32_t i = tile0 + KT * tid + k;
        tc[k] = inf;
        ts[k] = -1;
        ti[k] = i;
        if (i < tile_end) {
          if (front_list) {
            const uint4 rec = KD_OPT_CAND_CACHED ? __ldcg(B.front + i) : __ldcs(B.front + i);
            tc[k] = dunkey((static_cast<unsigned long long>(rec.y) << 32) | rec.x);
            ts[k] = static_cast<int32_t>(rec.z);
            ti[k] = rec.w;
          } else {
            tc[k] = __ldcs(cost + i);
            ts[k] = __ldcs(state + i);
          }
        }
        // faster-decoder.cc:202, split over the two passes (a hole has state -1, cost +inf)
        if (!(pass == 0 ? tc[k] < good : (tc[k] >= good && tc[k] < wcut))) ts[k] = -1;
      }
      int4 sa[KT], sb[KT];
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        sa[k] = make_int4(0, 0, 0, 0);
        sb[k] = make_int4(-1, 0, 0, 0);
        if (ts[k] >= 0) {
          gld_state(P.st + 2 * static_cast<size_t>(ts[k]), &sa[k], &sb[k]);
        }
      }
      // a valid bound on this frame's final cutoff (the seeded running cutoff)
      const double cut_seed = widen(funkey(*reinterpret_cast<volatile uint32_t *>(&sh.cut_fkey)));
      const float amin = sh.ac_min;
      const bool order_ok = sh.order_ok != 0;
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        beg[k] = static_cast<uint32_t>(sa[k].x);
        cnt[k] = static_cast<uint32_t>(sa[k].y);  // work items: arcs, or labels to look up
        tab[k] = -1;
        if (ts[k] >= 0) {
          ++ch_expanded;
          ch_arcs += cnt[k];
          if (order_ok && sb[k].x >= 0) {
            // an arc can only pass if ac < cutoff - cost - w <= slack (margin for fp rounding)
            const double slack = (cut_seed - tc[k]) - widen(__int_as_float(sb[k].y));
            const float sf = __double2float_ru(slack + 1e-6 * (fabs(cut_seed) + 1.0));
            // labels with ac < sf lie in buckets <= bucket(sf): a prefix of lab_order
            const float d = (sf - amin) * 16.0f;
            uint32_t kk = 0;
            if (d >= 0.0f) {
              const int b = d < static_cast<float>(kOrderBins) ? static_cast<int>(d) : kOrderBins;
              kk = bin_start[b + 1];
            }
            if (2 * kk < static_cast<uint32_t>(sa[k].y)) {
              tab[k] = sb[k].x;
              cnt[k] = kk;
            }
          }
        } else {
          cnt[k] = 0;
        }
      }
    }
    uint32_t my_arcs = 0, my_toks = 0;
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      my_arcs += cnt[k];
      my_toks += cnt[k] != 0;
    }
    uint32_t w_arcs, w_toks;
    uint32_t ex_arcs = warp_excl_scan(my_arcs, &w_arcs);
    uint32_t ex_toks = warp_excl_scan(my_toks, &w_toks);
    if (lane == 0) {
      sh.warp_sums[warp] = w_arcs;
      sh.warp_sums[16 + warp] = w_toks;
    }
    __syncthreads();
    uint32_t n_flat = 0, n_comp = 0;  // work items / tokens in the tile (uniform)
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const uint32_t a = sh.warp_sums[w], t = sh.warp_sums[16 + w];
      if (w < warp) {
        ex_arcs += a;
        ex_toks += t;
      }
      n_flat += a;
      n_comp += t;
    }
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      if (cnt[k] != 0) {
        t_ex[ex_toks] = ex_arcs;
        if (KD_OPT_COARSE && TT <= 1024 && (ex_toks & 31u) == 0) sh.ex_coarse[ex_toks >> 5] = ex_arcs;
        t_beg[ex_toks] = beg[k] | (tab[k] >= 0 ? kLookupFlag : 0u);
        t_tab[ex_toks] = tab[k];
        t_cost[ex_toks] = tc[k];
        t_tok[ex_toks] = ti[k];
        ex_arcs += cnt[k];
        ++ex_toks;
      }
    }
    ch_expanded = __reduce_add_sync(0xFFFFFFFFu, ch_expanded);
    ch_arcs = __reduce_add_sync(0xFFFFFFFFu, ch_arcs);
    if (lane == 0 && ch_expanded) {
      atomicAdd(&sh.acc_expanded, ch_expanded);
      atomicAdd(&sh.acc_emit, ch_arcs);
    }
    if (tid == 0) {
      t_ex[n_comp] = n_flat;
      sh.acc_items += n_flat;
    }
    __syncthreads();
    // flat item loop: 32-item windows are dealt round-robin to the warps, so all warps
    // start at the front of the flat space; kWin windows per step.
    {
      const uint32_t jw1 = n_flat;
#pragma unroll 1
      for (uint32_t jb0 = warp * 32; jb0 < jw1; jb0 += NW * 32 * kWin) {
        int2 iw[kWin];      // {ilabel, weight bits} of the item's arc
        uint32_t tt[kWin];  // compacted token of the item, kNoIdx if none
        uint32_t aa[kWin];  // emitting arc index of the item
#pragma unroll
        for (int u = 0; u < kWin; ++u) {
          const uint32_t jb = jb0 + u * (NW * 32);
          iw[u] = make_int2(1, 0);
          tt[u] = kNoIdx;
          aa[u] = kNoIdx;
          if (jb >= jw1) continue;  // warp-uniform
          // compacted token owning item jb = largest t with t_ex[t] <= jb: 32-ary search,
          // every lane probes one position per level (t_ex[0] = 0 <= jb always)
          uint32_t t_lo = 0;
#pragma unroll
          for (uint32_t stride = kSearchTop; stride; stride >>= 5) {
            const uint32_t pos = t_lo + lane * stride;
            // (the 32 probes of the stride-32 level would all fall into one bank of t_ex:
            // that level reads its own copy, one word per lane)
            const bool le =
                pos < n_comp &&
                ((KD_OPT_COARSE && TT <= 1024 && stride == 32) ? sh.ex_coarse[pos >> 5] : t_ex[pos]) <= jb;
            t_lo += (__popc(__ballot_sync(0xFFFFFFFFu, le)) - 1u) * stride;
          }
          // item -> token: bit mask of the token boundaries (first item index of the 32
          // tokens after t_lo) that fall inside the window
          const uint32_t j = jb + lane;
          const uint32_t bnd = t_ex[min(t_lo + 1 + lane, n_comp)];
          const uint32_t p = bnd - jb;  // >= 1: token t_lo owns item jb
          const uint32_t mask = __reduce_or_sync(0xFFFFFFFFu, p < 32 ? (1u << p) : 0u);
          const uint32_t t = t_lo + __popc(mask & ((2u << lane) - 1u));
          if (j < jw1) {
            const uint32_t b = t_beg[t];
            const uint32_t k = j - t_ex[t];
            if (b & kLookupFlag) {
              // the k-th best label of the frame, looked up in the state's label table:
              // one load gives the arc's weight and its offset within the state
              const uint32_t lab = lab_order[k];
              const int2 ent =
                  gld(P.labtab + static_cast<size_t>(t_tab[t]) * P.lab_stride + (lab - 1));
              if (ent.y >= 0) {  // else: the state has no arc with this label
                tt[u] = t;
                aa[u] = (b & ~kLookupFlag) + static_cast<uint32_t>(ent.y);
                iw[u] = make_int2(static_cast<int>(lab), ent.x);
              }
            } else {
              tt[u] = t;
              aa[u] = b + k;
              iw[u] = gld(P.e_iw + aa[u]);
            }
          }
        }
        const double cut_d = widen(funkey(*reinterpret_cast<volatile uint32_t *>(&sh.cut_fkey)));
#pragma unroll
        for (int u = 0; u < kWin; ++u) {
          const double ac = widen(-(ROW_SMEM ? s_row[iw[u].x - 1] : __ldcg(row_g + iw[u].x - 1)));
          const double tcst = t_cost[min(tt[u], static_cast<uint32_t>(TT - 1))];
          const double nw = (widen(__int_as_float(iw[u].y)) + tcst) + ac;
          // faster-decoder.cc:211 against the running cutoff
          const bool is_cand = tt[u] != kNoIdx && nw < cut_d;
          // Candidates are appended warp-aggregated: about a third of the looked-up arcs
          // pass the filter, and one shared-memory atomic per candidate on the single
          // counter serialised the whole CTA.
          const uint32_t cmask = __ballot_sync(0xFFFFFFFFu, is_cand);
          if (cmask == 0) continue;  // warp-uniform
          uint32_t cbase = 0;
          if (lane == 0) cbase = atomicAdd(&sh.cand_n, __popc(cmask));
          cbase = __shfl_sync(0xFFFFFFFFu, cbase, 0);
          if (is_cand) {
#if KD_OPT_PREFETCH_NO
            asm volatile("prefetch.global.L2 [%0];" ::"l"(P.e_no + aa[u]));
#endif
            const uint32_t tok_abs = base + t_tok[tt[u]];
            const unsigned long long nk = dkey(nw);
            const uint32_t e = cbase + __popc(cmask & ((1u << lane) - 1u));
            if (e < P.ccap) {
              if (KD_OPT_CAND_CACHED) {
                B.cand[e] = make_uint4(static_cast<uint32_t>(nk),
                                            static_cast<uint32_t>(nk >> 32), aa[u], tok_abs);
              } else {
                __stcs(B.cand + e, make_uint4(static_cast<uint32_t>(nk),
                                            static_cast<uint32_t>(nk >> 32), aa[u], tok_abs));
              }
            } else if (SIMPLE) {
              atomicOr(&sh.status, kStatusCandOverflow);
            } else {
              insert_arc(P, B, sh, ls.epoch, aa[u], nk, tok_abs);  // buffer full: recombine now
            }
            if (nw < my_min) {
              my_min = nw;
              // faster-decoder.cc:215-217
              const uint32_t fk = fkey(__double2float_ru(nw + ab));
              if (fk < *reinterpret_cast<volatile uint32_t *>(&sh.cut_fkey))
                atomicMin(&sh.cut_fkey, fk);
            }
          }
        }
      }
    }
    // every warp's items count before the next chunk is classified (tighter cutoff)
    __syncthreads();
    }
  }
  const long long t_scan_end = clock64();
  // ---------------------------------------------------------------- exact cutoff
  const double bmin = block_min<THREADS>(fmin(my_min, seed), sh, 0);
  const double cstar = bmin + ab;
  const unsigned long long cstar_key = dkey(cstar);
  // every thread is done with this frame's row (the barrier in block_min): the next frame's
  // row can take its place while the candidates are recombined and the frame is closed
  if (!SIMPLE && ROW_SMEM && next_row_g != nullptr && tid == 0) {
    bulk_load_row(s_row, next_row_g, static_cast<uint32_t>(P.cols) * 4u, &sh.row_bar);
    sh.row_pending = 1;
  }
  // ---------------------------------------------------------------- recombine
  const uint32_t n_cand = min(sh.cand_n, P.ccap);
  if (tid == 0) ls.st_cand += sh.cand_n;
  // The frame's region of the table: P.region_f entries per candidate.  An entry access is a
  // random 32-byte sector of 8 MB per lane; hashing a frame of ~1000 candidates into the first
  // 16 k entries instead keeps the bitmap words and most of the entries of the lanes in flight
  // in L2 (profiles/r2_sweeps.txt, section 14).  The candidates bound the emitting arrivals, not
  // what the epsilon closure adds: a region that fills up makes the frame start over with the
  // whole table (lane_closure_and_commit).
  if (P.region_f > 0 && !whole_table && sh.cand_n <= P.ccap) {
    __syncthreads();  // (the table is still empty: nothing was recombined during the scan)
    if (tid == 0) {
      uint32_t r = P.region_min;
      while (r < P.region_f * n_cand && r < P.hcap) r <<= 1;
      if (r < P.hcap) {
        sh.hmask = r - 1u;
        sh.hshift = 34 - (31 - __clz(r));
      }
    }
    __syncthreads();
  }
  // (Measured and rejected: taking 2-4 candidates per thread through the dependent
  // loads together -- the recombination gets faster, the other phases of the co-resident
  // lanes slower by as much; claiming the slot with the CAS before any probe load.)
  double min_stored = inf;  // SIMPLE only
#if KD_OPT_DEFER
  // Two passes.  Pass A: every arrival tries to claim its state's entry; the ones that do
  // (80 %) write it and are done.  An arrival that finds the entry in use is parked in
  // shared memory (the scan's tile arrays are idle now) instead of being recombined on the
  // spot: within a warp the two cases would run one after the other, and every step of the
  // warp would pay the entry load and the CAS of its few latecomers.  Pass B recombines the
  // parked arrivals, all threads on the same path; the barrier in between also means that
  // every entry they look at has been written.
  uint4 *parked = reinterpret_cast<uint4 *>(t_cost);
  constexpr uint32_t kParkCap = (24u * TT) / 16u;
  const bool defer = !SIMPLE && sh.load_first == 0;
  for (uint32_t e = tid; e < n_cand; e += THREADS) {
    const uint4 c = KD_OPT_CAND_CACHED ? __ldcg(B.cand + e) : __ldcs(B.cand + e);
    unsigned long long nk = (static_cast<unsigned long long>(c.y) << 32) | c.x;
    if (!(nk < cstar_key)) continue;  // faster-decoder.cc:211 / simple-decoder.cc:170, final cutoff
    if (SIMPLE) {
      // SimpleDecoder prunes on (cost + w) + ac but stores cost + float(w + ac)
      // (simple-decoder.cc:168 vs simple-decoder.h:96)
      const int2 iw = gld(P.e_iw + c.z);
      const float ac = -(ROW_SMEM ? s_row[iw.x - 1] : __ldcg(row_g + iw.x - 1));
      const double stored = B.a_cost[c.w] + static_cast<double>(__fadd_rn(__int_as_float(iw.y), ac));
      min_stored = fmin(min_stored, stored);
      nk = dkey(stored);
    }
    if (!defer) {
      insert_arc(P, B, sh, ls.epoch, c.z, nk, c.w);
      continue;
    }
    const int2 no = gld(P.e_no + c.z);
    const int32_t state = no.x & 0x7FFFFFFF;
    const uint32_t h = table_hash(sh, state);
    const uint32_t bit = 1u << (h & 31u);
    const uint32_t old = atomicOr(B.bitmap + (h >> 5), bit);
    if ((old & bit) == 0) {
      HVal mine;
      mine.cost = nk;
      mine.arg = (static_cast<unsigned long long>(c.z) << 32) | c.w;
      const uint32_t pos =
          register_claim(P, B, sh, h, state, no.x < 0 ? B.queue : nullptr, &sh.q_n[0]);
      st_entry(B.table + h, mine, state, pos, ls.epoch);
    } else {
      const uint32_t slot = atomicAdd(&sh.park_n, 1u);
      if (slot < kParkCap) {
        parked[slot] = make_uint4(h, static_cast<uint32_t>(no.x), e, 0u);
      } else {
        insert_arc(P, B, sh, ls.epoch, c.z, nk, c.w);  // no room: on the spot
      }
    }
  }
  __syncthreads();
  if (defer) {
    const uint32_t n_park = min(sh.park_n, kParkCap);
    for (uint32_t i = tid; i < n_park; i += THREADS) {
      const uint4 r = parked[i];
      // the arrival and the entry it met: two independent loads, one round trip
      const uint4 c = KD_OPT_CAND_CACHED ? __ldcg(B.cand + r.z) : __ldcs(B.cand + r.z);
      const EntryWords w = ld_entry(B.table + r.x);
      const int32_t state = static_cast<int32_t>(r.y & 0x7FFFFFFFu);
      HVal mine;
      mine.cost = (static_cast<unsigned long long>(c.y) << 32) | c.x;
      mine.arg = (static_cast<unsigned long long>(c.z) << 32) | c.w;
      uint32_t h = r.x;
      HVal cur = w.val;
      if (w.key != state || w.epoch != ls.epoch) {
        // the entry belongs to another state (or, parked by the overflow path of another
        // thread, is still on its way): the regular walk, from this entry on
        bool owner;
        h = table_arrive(P, B, sh, ls.epoch, state, mine,
                         static_cast<int32_t>(r.y) < 0 ? B.queue : nullptr, &sh.q_n[0], true,
                         &owner, &cur,
                         w.key != state && w.epoch == ls.epoch ? ((r.x + 1) & sh.hmask) : r.x);
        if (h == kNoIdx || owner) continue;
      }
      while (mine.cost < cur.cost || (mine.cost == cur.cost && mine.arg < cur.arg)) {
        HVal got = cas_hval(&B.table[h].val, cur, mine);
        if (got.cost == cur.cost && got.arg == cur.arg) break;
        cur = got;
      }
    }
  }
#else
  for (uint32_t e = tid; e < n_cand; e += THREADS) {
    const uint4 c = KD_OPT_CAND_CACHED ? __ldcg(B.cand + e) : __ldcs(B.cand + e);
    unsigned long long nk = (static_cast<unsigned long long>(c.y) << 32) | c.x;
    if (!(nk < cstar_key)) continue;  // faster-decoder.cc:211 / simple-decoder.cc:170, final cutoff
    if (SIMPLE) {
      // SimpleDecoder prunes on (cost + w) + ac but stores cost + float(w + ac)
      // (simple-decoder.cc:168 vs simple-decoder.h:96)
      const int2 iw = gld(P.e_iw + c.z);
      const float ac = -(ROW_SMEM ? s_row[iw.x - 1] : __ldcg(row_g + iw.x - 1));
      const double stored = B.a_cost[c.w] + static_cast<double>(__fadd_rn(__int_as_float(iw.y), ac));
      min_stored = fmin(min_stored, stored);
      nk = dkey(stored);
    }
    insert_arc(P, B, sh, ls.epoch, c.z, nk, c.w);
  }
#endif
  double closure_cutoff = cstar;
  if (SIMPLE) {
    // ProcessNonemitting's cutoff: best stored cost + beam (simple-decoder.cc:196-204)
    const double smin = block_min<THREADS>(min_stored, sh, 1);
    closure_cutoff = smin + static_cast<double>(P.beam);
  }
  __syncthreads();
  if (tid == 0) {
    const long long t_end = clock64();
    ls.cyc_cutoff += sh.t_mark - t_begin;
    ls.cyc_expand += t_end - sh.t_mark;
    ls.cyc_scan += t_scan_end - sh.t_mark;
  }
  return closure_cutoff;
}

// ------------------------------------------------------------------ kernels

// Dynamic shared memory of kd_advance_kernel without the log-prob row, bytes.
template <int THREADS>
__host__ __device__ constexpr size_t advance_smem_fixed() {
  return THREADS * kTileTokens * (sizeof(double) + 4) +  // t_cost, t_beg
         (THREADS * kTileTokens + 4) * 4 +               // t_ex
         THREADS * kTileTokens * 4 +                     // t_tab
         THREADS * kTileTokens * 4 +                     // t_tok
         (kOrderBins + 2) * 2 + 12;                      // bin_start (+ pad to 16)
}

// ROW_SMEM is a template parameter of the kernel (not a run-time branch): the
// kernel is instruction-cache sensitive (7 lanes per SM run different phases of a
// ~3500-instruction body), so only the variant in use is instantiated per launch.
// Registers per thread such that MIN_BLOCKS lanes of THREADS threads fit an SM: the register
// file is four 16 K-register partitions (one per scheduler), a warp's registers come out
// of one of them, in units of 8 per thread.
constexpr int advance_max_regs(int threads, int min_blocks) {
  const int warps_per_partition = (min_blocks * (threads / 32) + 3) / 4;
  const int r = (16384 / warps_per_partition / 32) / 8 * 8;
  return r > 255 ? 255 : r;
}

// The start token (faster-decoder.cc:46-52): cost 0 at Start(), no arc, no predecessor.
// One thread; the caller closes it over the epsilon arcs and commits.
__device__ __forceinline__ void lane_start_token(const Params &P, const LaneBuf &B, Shared &sh,
                                                 uint32_t epoch) {
  const int4 st = gld(P.st + 2 * static_cast<size_t>(P.start));
  HVal v;
  v.cost = dkey(0.0);
  v.arg = (static_cast<unsigned long long>(kNoArc) << 32) | kNoPrev;
  bool owner;
  HVal unused;
  table_arrive(P, B, sh, epoch, P.start, v, st.w > 0 ? B.queue : nullptr, &sh.q_n[0], true, &owner,
               &unused);
}

__device__ __forceinline__ void lane_state_reset(LaneState &ls) {
  // (word by word, in place: a local copy of the struct would live on the stack)
  const uint32_t epoch = ls.epoch;  // entries written before this InitDecoding must stay stale
  uint32_t *w = reinterpret_cast<uint32_t *>(&ls);
#pragma unroll 1
  for (int i = 0; i < static_cast<int>(sizeof(LaneState) / 4); ++i) w[i] = 0u;
  ls.epoch = epoch;
  ls.best_cost = __longlong_as_double(0x7FF0000000000000ll);
  ls.best_idx = -1;
  ls.best_state = -1;
}

// Writes one arc of the best path (faster-decoder.cc:393-402: graph = arc weight,
// acoustic = float(cost - prev cost) - graph).
__device__ __forceinline__ void write_path_arc(const Params &P, uint32_t arc, double c, double pc,
                                               long long pos, int32_t *il, int32_t *ol, float *gw,
                                               float *aw) {
  int32_t ilab, olab;
  float graph;
  if (arc & kEpsFlag) {
    const int4 a = gld(P.n_arc + (arc & ~kEpsFlag));
    ilab = 0;
    olab = a.x;
    graph = __int_as_float(a.y);
  } else {
    const int2 iw = gld(P.e_iw + arc);
    const int2 no = gld(P.e_no + arc);
    ilab = iw.x;
    olab = no.y;
    graph = __int_as_float(iw.y);
  }
  const float tot = static_cast<float>(c - pc);
  il[pos] = ilab;
  ol[pos] = olab;
  gw[pos] = graph;
  aw[pos] = tot - graph;
}

// ReachedFinal (faster-decoder.cc:347-354): does a live token sit in a final state?
// All threads return the answer.  SimpleDecoder::PruneToks (simple-decoder.cc:251-279)
// runs after every frame: only tokens with cost < best + beam exist for ReachedFinal /
// GetBestPath.  The search applies the same test when it expands the next frame; here it
// is applied to the view.
template <int THREADS>
__device__ int lane_reached_final(const Params &P, const LaneBuf &B, Shared &sh,
                                  const LaneState &L) {
  const int tid = threadIdx.x;
  const int n = L.n_tok;
  const uint32_t base = L.tok_base;
  const double inf = __longlong_as_double(0x7FF0000000000000ll);
  const bool pruned_view = P.simple && L.frames_decoded > 0;
  const double limit = pruned_view ? L.best_cost + static_cast<double>(P.beam) : inf;
  if (tid == 0) sh.any_final = 0;
  __syncthreads();
  int any = 0;
  for (int i = tid; i < n; i += THREADS) {
    const int s = B.a_state[base + i];
    if (s < 0) continue;  // a hole, not a token
    const double c = B.a_cost[base + i];
    if (pruned_view && !(c < limit)) continue;
    const float f = __ldg(P.fin + s);
    if (c != inf && f != __int_as_float(0x7F800000)) any = 1;
  }
  if (any) atomicOr(&sh.any_final, 1);
  __syncthreads();
  return sh.any_final;
}

// ReachedFinal + best-token selection + the one walk of the backpointer chain
// (faster-decoder.cc:347-402).  Ties on the selection cost go to the lowest state id.
// The token indices met on the walk are parked in the lane's candidate buffer (idle
// between frames), so the arcs can then be written out by all threads at once.
template <int THREADS>
__device__ __noinline__ void lane_select_best(const Params &P, const LaneBuf &B, Shared &sh,
                                              LaneState &L) {
  const int tid = threadIdx.x;
  const int n = L.n_tok;
  const uint32_t base = L.tok_base;
  const double inf = __longlong_as_double(0x7FF0000000000000ll);
  const bool pruned_view = P.simple && L.frames_decoded > 0;
  const double limit = pruned_view ? L.best_cost + static_cast<double>(P.beam) : inf;
  const int is_final = lane_reached_final<THREADS>(P, B, sh, L);
  if (tid == 0) sh.best_tok = kNoIdx;
  double bv = inf;
  int bs = -1;  // state id is the tie-break key; token index recovered below
  for (int i = tid; i < n; i += THREADS) {
    const int s = B.a_state[base + i];
    if (s < 0) continue;
    const double c = B.a_cost[base + i];
    if (pruned_view && !(c < limit)) continue;
    const double v = is_final ? c + static_cast<double>(__ldg(P.fin + s)) : c;
    const bool take = is_final ? (v != inf) : true;
    if (take && (bs < 0 || v < bv || (v == bv && s < bs))) {
      bv = v;
      bs = s;
    }
  }
  double rv;
  int rs;
  // block_min_arg treats idx -1 as "none" (largest unsigned)
  block_min_arg<THREADS>(bs < 0 ? inf : bv, bs, sh, &rv, &rs);
  if (rs >= 0) {
    for (int i = tid; i < n; i += THREADS)
      if (B.a_state[base + i] == rs) sh.best_tok = base + i;
  }
  __syncthreads();
  if (tid == 0) {
    L.bp_final = is_final;
    L.bp_parked = 0;
    if (rs < 0 || sh.best_tok == kNoIdx) {
      L.bp_ok = 0;
      L.bp_len = 0;
      L.bp_stored = 0;
      L.bp_best_tok = kNoIdx;
      L.bp_best_state = -1;
      L.bp_final_w = 0.f;
      L.bp_value = inf;
    } else {
      L.bp_value = rv;
      uint32_t *path = reinterpret_cast<uint32_t *>(B.cand);
      const long long path_cap = 4ll * P.ccap;
      long long len = 0;
      uint32_t t = sh.best_tok;
      while (true) {
        const unsigned long long link = B.a_link[t];
        const uint32_t arc = static_cast<uint32_t>(link >> 32);
        if (arc == kNoArc) break;
        if (len < path_cap) path[len] = t;
        ++len;
        t = static_cast<uint32_t>(link);
      }
      L.bp_ok = 1;
      L.bp_stored = len <= path_cap ? 1 : 0;
      L.bp_len = len;
      L.bp_best_tok = sh.best_tok;
      L.bp_best_state = rs;
      L.bp_final_w = __ldg(P.fin + rs);
    }
  }
  __syncthreads();
}

// Writes the arcs of the selected path, in time order, into four arrays starting at
// il/ol/gw/aw[first].  With the token indices of the path at hand every arc is independent;
// otherwise (path longer than the candidate buffer) thread 0 walks the chain again.
__device__ __forceinline__ void lane_write_path(const Params &P, const LaneBuf &B,
                                                const LaneState &L, long long first, int32_t *il,
                                                int32_t *ol, float *gw, float *aw) {
  const long long len = L.bp_len;
  const long long last = first + len - 1;
  if (L.bp_stored) {
    const uint32_t *path = reinterpret_cast<const uint32_t *>(B.cand);
    for (long long i = threadIdx.x; i < len; i += blockDim.x) {
      const uint32_t t = path[i];
      const unsigned long long link = B.a_link[t];
      const uint32_t prev = static_cast<uint32_t>(link);
      write_path_arc(P, static_cast<uint32_t>(link >> 32), B.a_cost[t], B.a_cost[prev], last - i,
                     il, ol, gw, aw);
    }
    return;
  }
  if (threadIdx.x != 0) return;
  long long pos = last;
  uint32_t t = L.bp_best_tok;
  double c = B.a_cost[t];
  unsigned long long link = B.a_link[t];
  while (true) {
    const uint32_t arc = static_cast<uint32_t>(link >> 32);
    if (arc == kNoArc) break;
    const uint32_t prev = static_cast<uint32_t>(link);
    // the pointer chase is the critical path: its next load goes out before the
    // loads that only feed this step's output
    const unsigned long long link_next = B.a_link[prev];
    const double pc = B.a_cost[prev];
    write_path_arc(P, arc, c, pc, pos, il, ol, gw, aw);
    --pos;
    t = prev;
    c = pc;
    link = link_next;
  }
}

// kItemFinalize: GetBestPath inside the search launch.  The lane's CTA selects the best
// token, walks its backpointers and parks the path's arcs in the lane's epsilon-worklist
// buffer (idle once the last frame is committed) as four arrays of `path_cap` words
// [ilabel | olabel | graph | acoustic]; the host then fetches them with plain copies and
// no further kernel.  A path that does not fit stays unparked (the host falls back to
// kd_best_fill_kernel).
template <int THREADS>
__device__ __noinline__ void lane_finalize(const Params &P, const LaneBuf &B, Shared &sh,
                                           LaneState &L, int32_t path_cap) {
  lane_select_best<THREADS>(P, B, sh, L);
  if (!L.bp_ok) return;
  const long long room = path_cap < static_cast<long long>(P.qcap / 2) ? path_cap : P.qcap / 2;
  if (L.bp_len > room || path_cap <= 0) return;
  int32_t *out = reinterpret_cast<int32_t *>(B.queue);
  lane_write_path(P, B, L, 0, out, out + path_cap, reinterpret_cast<float *>(out + 2 * path_cap),
                  reinterpret_cast<float *>(out + 3 * path_cap));
  __syncthreads();
  if (threadIdx.x == 0) L.bp_parked = path_cap;
}

template <int THREADS, int MIN_BLOCKS, bool ROW_SMEM, bool SIMPLE = false>
__global__ void __launch_bounds__(THREADS) __maxnreg__(advance_max_regs(THREADS, MIN_BLOCKS))
    kd_advance_kernel(const __grid_constant__ Params P) {
  constexpr int TT = THREADS * kTileTokens;
  __shared__ Shared sh;
  __shared__ LaneState ls;
  __shared__ LaneBuf sB;  // per-lane base pointers live in shared memory, not registers
  __shared__ uint32_t s_list[kListSmem > 0 ? kListSmem : 1];
  const LaneBuf &B = sB;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  double *t_cost = reinterpret_cast<double *>(dyn_smem);
  uint32_t *t_beg = reinterpret_cast<uint32_t *>(t_cost + TT);
  uint32_t *t_ex = t_beg + TT;  // TT + 1 entries (+ pad to 4)
  int32_t *t_tab = reinterpret_cast<int32_t *>(t_ex + TT + 4);
  uint32_t *t_tok = reinterpret_cast<uint32_t *>(t_tab + TT);
  uint16_t *bin_start = reinterpret_cast<uint16_t *>(t_tok + TT);  // kOrderBins + 2 entries
  float *s_row = reinterpret_cast<float *>(bin_start + kOrderBins + 8);  // 16-byte aligned
  // the label order (one uint16 per column) follows the row
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&sh.row_bar, 1);
    sh.row_parity = 0;
    sh.row_pending = 0;
  }
  __syncthreads();

  while (true) {
    if (tid == 0) sh.item = atomicAdd(P.work_counter, 1);
    __syncthreads();
    const int item = sh.item;
    // a lane that stopped on an error may have left the bulk copy of its next row in flight:
    // it is drained before the row buffer gets a new owner (or the CTA exits)
    if (sh.row_pending != 0) {
      mbar_wait(&sh.row_bar, sh.row_parity);
      __syncthreads();
      if (tid == 0) {
        sh.row_parity ^= 1u;
        sh.row_pending = 0;
      }
      __syncthreads();
    }
    if (item >= P.n_items) return;
    const AdvanceItem it = P.items[item];
    // InitDecoding in the same launch: the first pass of the loop below has no emitting
    // phase -- the start token is closed over the epsilon arcs under FLT_MAX
    // (faster-decoder.cc:53; SimpleDecoder: under 0 + beam, simple-decoder.cc:29-41)
    bool init_pass = (it.flags & kItemInit) != 0;
    if (tid == 0) {
      sB = lane_buffers(P, it.lane);
      sB.s_list = s_list;
      ls = P.lanes[it.lane];
      if (init_pass) lane_state_reset(ls);
      sh.status = ls.status;
      sh.list_n = 0;
      sh.q_n[0] = 0;
      sh.rows_ready = P.progress != nullptr ? 0 : 0x7FFFFFFF;
      sh.yield = 0;
      sh.load_first = 0;
      sh.cand_n = 0;
      table_region_reset(P, sh);
    }
    __syncthreads();
    bool whole_table = false;  // the frame is a second attempt (its table region filled up)
    while (true) {
      double cstar, good_cut, mid_cut;
      int n_in = 0, frame = 0;
      if (init_pass) {
        if (tid == 0) lane_start_token(P, B, sh, ls.epoch);
        cstar = SIMPLE ? static_cast<double>(P.beam) : 3.4028234663852886e+38 /* FLT_MAX */;
        good_cut = mid_cut = 0.0;
        __syncthreads();
      } else {
        if (!(ls.frames_decoded < it.target && sh.status == 0)) break;
        frame = ls.frames_decoded;
        if (frame - it.offset >= sh.rows_ready) {
          // the row has not been seen to arrive yet: poll the copy stream's progress word
          __syncthreads();  // every thread has read rows_ready before thread 0 rewrites it
          if (tid == 0) {
            const volatile int32_t *pr = P.progress;
            volatile int32_t *fl = P.yield_flag;
            int32_t ready = *pr;
            const long long t0 = clock64();
            while (ready <= frame - it.offset) {
              // Nothing for ~1 s: the copies are not coming while this kernel runs
              // (launches are serialised -- CUDA_LAUNCH_BLOCKING, a profiler -- or the host
              // is staging pageable memory).  The lane yields: its state is saved as it is
              // and the host launches again once the copies are enqueued.
              if (clock64() - t0 > kYieldCycles || *fl != 0) {
                sh.yield = 1;
                *fl = 1;  // the other lanes need not wait as long
                break;
              }
              __nanosleep(256);
              ready = *pr;
            }
            __threadfence();
            sh.rows_ready = ready;
            ls.cyc_wait += clock64() - t0;
          }
          __syncthreads();
          if (sh.yield != 0) break;
        }
        const float *row_g = it.logp + static_cast<size_t>(frame - it.offset) * P.cols;
        // the next frame's row, if it is known to be there and fit for a bulk copy
        const float *next_row_g = nullptr;
        if (ROW_SMEM && frame + 1 < it.target && frame + 1 - it.offset < sh.rows_ready &&
            row_bulk_ok(row_g + P.cols, P.cols))
          next_row_g = row_g + P.cols;
        n_in = ls.n_live;
        uint16_t *lab_order = reinterpret_cast<uint16_t *>(s_row + (ROW_SMEM ? P.cols : 0));
        cstar = lane_expand_emitting<THREADS, ROW_SMEM, SIMPLE>(P, B, sh, ls, row_g, next_row_g,
                                                                s_row, t_cost, t_ex, t_beg, t_tab,
                                                                t_tok, lab_order, bin_start, whole_table);
        // min(new_weight) = cstar - adaptive_beam is not kept; cstar - beam is at least as large
        good_cut = cstar - 0.75 * static_cast<double>(P.beam);
        mid_cut = cstar - 0.25 * static_cast<double>(P.beam);
      }
      if (lane_closure_and_commit<THREADS, SIMPLE>(P, B, sh, ls, cstar, good_cut, mid_cut)) {
        // The frame's table region filled up: the same frame once more, with the whole table.
        // The next frame's row may be on its way into the row buffer: it is waited for, and
        // this frame's row is loaded again.
        if (sh.row_pending != 0) {
          mbar_wait(&sh.row_bar, sh.row_parity);
          __syncthreads();
          if (tid == 0) {
            sh.row_parity ^= 1u;
            sh.row_pending = 0;
          }
          __syncthreads();
        }
        whole_table = true;
        continue;
      }
      whole_table = false;
      if (tid == 0) {
        if (init_pass) {
          ls.st_sweeps = 0;
          ls.st_eps_arcs = 0;
          ls.cyc_closure = ls.cyc_commit = 0;
          ls.st_claimed = 0;
        } else {
          ls.frames_decoded = frame + 1;
          ls.st_frames += 1;
          ls.st_tokens_in += n_in;
          ls.st_tokens_out += ls.n_live;
          ls.st_emit_arcs += sh.acc_emit;
          ls.st_expanded += sh.acc_expanded;
          ls.st_items += sh.acc_items;
          if (ls.n_live > ls.st_max_tokens) ls.st_max_tokens = ls.n_live;
        }
      }
      init_pass = false;
      __syncthreads();
    }
    if ((it.flags & kItemFinalize) != 0 && sh.status == 0 && sh.yield == 0 &&
        ls.frames_decoded >= it.target)
      lane_finalize<THREADS>(P, B, sh, ls, it.path_cap);
    if (tid == 0) {
      ls.status = sh.status;
      P.lanes[it.lane] = ls;
    }
    __syncthreads();
  }
}

// InitDecoding alone (faster-decoder.cc:42-56): start token with cost 0, epsilon
// closure under cutoff FLT_MAX, zero frames decoded.  items[i].lane = lane.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) kd_init_kernel(Params P) {
  __shared__ Shared sh;
  __shared__ LaneState ls;
  const int tid = threadIdx.x;
  const int item = blockIdx.x;
  if (item >= P.n_items) return;
  const int lane = P.items[item].lane;
  __shared__ uint32_t s_list[kListSmem > 0 ? kListSmem : 1];
  LaneBuf B = lane_buffers(P, lane);
  B.s_list = s_list;
  if (tid == 0) {
    ls = P.lanes[lane];
    lane_state_reset(ls);
    sh.status = 0;
    sh.list_n = 0;
    sh.q_n[0] = 0;
    sh.load_first = 0;
    sh.cand_n = 0;
    table_region_reset(P, sh);
  }
  __syncthreads();
  if (tid == 0) lane_start_token(P, B, sh, ls.epoch);
  __syncthreads();
  if (P.simple) {
    // SimpleDecoder::InitDecoding: closure under cutoff 0 + beam (simple-decoder.cc:29-41,196-204)
    lane_closure_and_commit<THREADS, true>(P, B, sh, ls, static_cast<double>(P.beam), 0.0, 0.0);
  } else {
    lane_closure_and_commit<THREADS, false>(P, B, sh, ls, 3.4028234663852886e+38 /* FLT_MAX */,
                                            0.0, 0.0);
  }
  if (tid == 0) {
    ls.status = sh.status;
    ls.st_sweeps = 0;
    ls.st_eps_arcs = 0;
    ls.cyc_closure = ls.cyc_commit = 0;
    ls.st_claimed = 0;
    P.lanes[lane] = ls;
  }
}

// GetBestPath step 1 as a kernel of its own (lanes that were not finalized by their last
// search launch: partial results while streaming).  One CTA per lane.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) kd_best_select_kernel(const __grid_constant__ Params P) {
  __shared__ Shared sh;
  const int lane = P.items[blockIdx.x].lane;
  const LaneBuf B = lane_buffers(P, lane);
  lane_select_best<THREADS>(P, B, sh, P.lanes[lane]);
}

// ReachedFinal alone: no backpointer walk.  out[b] = flag of lane items[b].lane.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) kd_reached_final_kernel(Params P, int32_t *out) {
  __shared__ Shared sh;
  const int lane = P.items[blockIdx.x].lane;
  const LaneBuf B = lane_buffers(P, lane);
  const int f = lane_reached_final<THREADS>(P, B, sh, P.lanes[lane]);
  if (threadIdx.x == 0) out[blockIdx.x] = f;
}

// GetBestPath step 2: the best path of lane items[b].lane in time order at out_off[b].
__global__ void kd_best_fill_kernel(Params P, const long long *out_off, int32_t *il, int32_t *ol,
                                    float *gw, float *aw) {
  const int b = blockIdx.x;
  if (b >= P.n_items) return;
  const int lane = P.items[b].lane;
  const LaneBuf B = lane_buffers(P, lane);
  const LaneState *L = P.lanes + lane;
  if (!L->bp_ok) return;
  lane_write_path(P, B, *L, out_off[b], il, ol, gw, aw);
}

}  // namespace kd

#endif  // KD_KERNELS_CUH_
