// kaldi-decoder_b200/csrc/kd_kernels.cuh
//
// sm_100a kernels of the token-passing Viterbi beam search.  One CTA owns one
// utterance lane for a whole AdvanceDecoding call (all its frames), so the
// frame loop needs no grid-wide synchronisation: every phase boundary is a
// __syncthreads().  Reference semantics implemented:
//
//   GetCutoff           faster-decoder.cc:244-336   -> lane_cutoff()
//   ProcessEmitting     faster-decoder.cc:155-241   -> lane_expand_emitting()
//   ProcessNonemitting  faster-decoder.cc:59-119    -> lane_closure()
//   token list / Token  faster-decoder.h:110-156,
//                       hash-list-inl.h:127-173     -> per-lane open-addressing
//                       table (key = state, 128-bit value = ordered fp64 cost |
//                       arc | backpointer) + frame-major backpointer arena
//   InitDecoding        faster-decoder.cc:42-56     -> kd_init_kernel
//   ReachedFinal /      faster-decoder.cc:347-354,
//   GetBestPath         356-424                     -> kd_best_select_kernel,
//                                                      kd_best_fill_kernel
//
// Costs are fp64 sums of fp32 addends, added in the reference's order
// ((w + cost) + ac, faster-decoder.cc:210), so they are bit-identical to the
// CPU's.  The reference's running next-frame cutoff is order dependent; its
// final value C* = min(new_weight) + adaptive_beam is not.  The kernel keeps a
// running cutoff in shared memory only as a filter (anything admitted with
// cost >= C* is ignored later), which makes the surviving token set exactly
// {new_weight < C*}, independent of thread scheduling.  Equal-cost arrivals at
// a state are resolved towards the lowest emitting-arc index (deterministic);
// in the epsilon closure the incumbent stays (as in the reference).
#ifndef KD_KERNELS_CUH_
#define KD_KERNELS_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

namespace kd {

constexpr int kStatusHashOverflow = 1;
constexpr int kStatusArenaOverflow = 2;
constexpr int kStatusQueueOverflow = 4;
constexpr int kStatusPathOverflow = 8;

constexpr uint32_t kEpsFlag = 0x80000000u;
constexpr uint32_t kNoArc = 0x7FFFFFFFu;   // the start token's "arc"
constexpr uint32_t kNoPrev = 0xFFFFFFFFu;
constexpr int32_t kEmptyKey = -1;
constexpr unsigned long long kEmptyCost = 0xFFFFFFFFFFFFFFFFull;
constexpr unsigned long long kEmptyArg = 0xFFFFFFFFFFFFFFFFull;
constexpr uint32_t kNoIdx = 0xFFFFFFFFu;

struct __align__(16) HVal {
  unsigned long long cost;  // order-preserving image of the fp64 cost
  unsigned long long arg;   // (arc << 32) | prev
};

// One recombination-table entry = one 32-byte sector: a probe, the 128-bit
// value CAS, the commit numbering and the wipe of a state all touch the same
// sector (the first layout used three arrays = three sectors per state and
// thrashed L2: profiles/r1_v0_ncu_summary.txt).
struct __align__(32) Entry {
  HVal val;       // 16-byte aligned: target of the 128-bit CAS
  int32_t key;    // state id, kEmptyKey when free
  uint32_t idx;   // commit pass: index of the token in the next block
  uint32_t pad[2];
};

// Everything the device keeps per lane between calls.
struct __align__(16) LaneState {
  int32_t n_tok;           // tokens alive (the reference's toks_ list length)
  int32_t frames_decoded;  // num_frames_decoded_; -1 before InitDecoding
  int32_t status;          // kStatus* bits; non-zero = lane unusable until init
  int32_t best_idx;        // index (in the current token block) of a best token
  uint32_t tok_base;       // arena index of the current token block
  uint32_t arena_used;     // arena records in use
  double best_cost;        // min cost over the current tokens (+inf if none)
  // counters (kd_stats)
  long long st_frames, st_tokens_in, st_expanded, st_emit_arcs, st_eps_arcs,
      st_tokens_out, st_max_tokens, st_sweeps;
  // SM cycles spent per phase (clock64 of thread 0), for the phase breakdown
  long long cyc_cutoff, cyc_expand, cyc_closure, cyc_commit;
  // best-path selection results
  int32_t bp_ok, bp_final, bp_best_state;
  uint32_t bp_best_tok;    // arena index
  long long bp_len;
  float bp_final_w;
  int32_t pad0;
};

struct AdvanceItem {
  int32_t lane;
  int32_t rows;
  int32_t offset;
  int32_t target;       // frames_decoded to reach
  const float *logp;    // device pointer, row-major rows x cols
};

struct Params {
  // graph (device)
  const int4 *st;      // [S]  {emit_begin, emit_count, eps_begin, eps_count}
  const int4 *e_arc;   // [Ee] {ilabel, weight bits, nextstate, olabel}
  const int4 *n_arc;   // [En] {olabel, weight bits, nextstate, 0}
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
  uint32_t *list;
  uint32_t *queue;  // 2 * qcap per lane
  uint32_t hcap, hmask, lcap, qcap;
  int32_t hshift;
  int32_t cols;
  int32_t row_in_smem;
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

__device__ __forceinline__ HVal ld_hval(const HVal *p) {
  ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2 *>(p));
  HVal r;
  r.cost = v.x;
  r.arg = v.y;
  return r;
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

struct Shared {
  double cstar;                // C* of the frame being processed
  double red_d[32];
  int red_i[32];
  uint32_t hist[256];
  uint32_t cut_fkey;  // running next-frame cutoff, rounded UP to float (a filter only)
  uint32_t acc_emit, acc_eps, acc_expanded;  // per-frame counters
  uint32_t list_n;
  uint32_t q_n[2];
  uint32_t out_n;
  uint32_t chunk;
  uint32_t sel_bin, sel_k;
  long long t_mark;
  int status;
  int item;
};

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
  uint32_t *list;
  uint32_t *queue;
};

__device__ __forceinline__ LaneBuf lane_buffers(const Params &P, int lane) {
  LaneBuf b;
  size_t L = static_cast<size_t>(lane);
  b.a_cost = P.a_cost + L * P.arena_cap;
  b.a_link = P.a_link + L * P.arena_cap;
  b.a_state = P.a_state + L * P.arena_cap;
  b.table = P.table + L * P.hcap;
  b.list = P.list + L * P.lcap;
  b.queue = P.queue + L * 2 * P.qcap;
  return b;
}

// Finds the table slot of `state`, claiming an empty one if needed (then the
// slot is appended to this frame's slot list).  Returns kNoIdx on overflow.
__device__ __forceinline__ uint32_t table_slot(const Params &P, const LaneBuf &B, Shared &sh,
                                               int32_t state) {
  // groups of 4 consecutive states share a 128-byte line; groups are scattered
  uint32_t h = ((((static_cast<uint32_t>(state) >> 2) * 0x9E3779B1u) >> P.hshift) << 2) |
               (static_cast<uint32_t>(state) & 3u);
  for (uint32_t probe = 0; probe < P.hcap; ++probe) {
    int32_t k = __ldcg(&B.table[h].key);
    if (k == state) return h;
    if (k == kEmptyKey) {
      int32_t old = atomicCAS(&B.table[h].key, kEmptyKey, state);
      if (old == kEmptyKey) {
        uint32_t pos = atomicAdd(&sh.list_n, 1u);
        if (pos < P.lcap) {
          B.list[pos] = h;
        } else {
          atomicOr(&sh.status, kStatusHashOverflow);
        }
        return h;
      }
      if (old == state) return h;
    }
    h = (h + 1) & P.hmask;
  }
  atomicOr(&sh.status, kStatusHashOverflow);
  return kNoIdx;
}

// Emitting-phase recombination: keep the lexicographic minimum of (cost, arg).
__device__ __forceinline__ void table_min(HVal *slot, HVal mine) {
  HVal cur = ld_hval(slot);
  while (mine.cost < cur.cost || (mine.cost == cur.cost && mine.arg < cur.arg)) {
    HVal got = cas_hval(slot, cur, mine);
    if (got.cost == cur.cost && got.arg == cur.arg) return;
    cur = got;
  }
}

// GetCutoff's order statistic: the k-th smallest (0-based) of float(cost[i]).
template <int THREADS>
__device__ float select_kth(const double *cost, int n, uint32_t k, Shared &sh) {
  uint32_t prefix = 0, mask = 0;
  for (int pass = 3; pass >= 0; --pass) {
    const int shift = pass * 8;
    for (int i = threadIdx.x; i < 256; i += THREADS) sh.hist[i] = 0;
    __syncthreads();
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

// faster-decoder.cc:244-336.  n tokens, best = min cost.  All threads return
// the same (weight_cutoff, adaptive_beam).
template <int THREADS>
__device__ void lane_cutoff(const Params &P, const double *cost, int n, double best, Shared &sh,
                            double *weight_cutoff, float *adaptive_beam) {
  const double inf = __longlong_as_double(0x7FF0000000000000ll);
  if (P.max_active == 0x7FFFFFFF && P.min_active == 0) {
    *adaptive_beam = P.beam;
    *weight_cutoff = best + static_cast<double>(P.beam);
    return;
  }
  const double beam_cutoff = best + static_cast<double>(P.beam);
  double max_cut = inf, min_cut = inf;
  if (n > P.max_active)
    max_cut = static_cast<double>(select_kth<THREADS>(cost, n, P.max_active, sh));
  if (max_cut < beam_cutoff) {
    *adaptive_beam = static_cast<float>(max_cut - best + static_cast<double>(P.beam_delta));
    *weight_cutoff = max_cut;
    return;
  }
  if (n > P.min_active) {
    if (P.min_active == 0)
      min_cut = best;
    else
      min_cut = static_cast<double>(select_kth<THREADS>(cost, n, P.min_active, sh));
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
__device__ __forceinline__ void eps_arrival(const Params &P, const LaneBuf &B, Shared &sh,
                                            int32_t dst, unsigned long long cost_key,
                                            uint32_t arc, uint32_t src_slot,
                                            unsigned long long cstar_key, uint32_t *q_next,
                                            uint32_t *q_next_n) {
  uint32_t h = table_slot(P, B, sh, dst);
  if (h == kNoIdx) return;
  HVal mine;
  mine.cost = cost_key;
  mine.arg = (static_cast<unsigned long long>(arc | kEpsFlag) << 32) | src_slot;
  HVal cur = ld_hval(&B.table[h].val);
  while (true) {
    bool cur_is_eps = (cur.arg >> 63) != 0;
    bool replace = mine.cost < cur.cost || (!cur_is_eps && !(cur.cost < cstar_key));
    if (!replace) return;
    HVal got = cas_hval(&B.table[h].val, cur, mine);
    if (got.cost == cur.cost && got.arg == cur.arg) break;
    cur = got;
  }
  uint32_t pos = atomicAdd(q_next_n, 1u);
  if (pos < P.qcap) {
    q_next[pos] = h;
  } else {
    atomicOr(&sh.status, kStatusQueueOverflow);
  }
}

__device__ __forceinline__ void expand_eps(const Params &P, const LaneBuf &B, Shared &sh,
                                           uint32_t slot, unsigned long long cstar_key,
                                           double cstar, uint32_t *q_next, uint32_t *q_next_n,
                                           uint32_t *eps_count) {
  HVal v = ld_hval(&B.table[slot].val);
  bool is_eps = (v.arg >> 63) != 0;
  // a token iff cost < C*, or it came from an epsilon arc (then cost <= C*)
  if (v.cost == kEmptyCost || !(v.cost < cstar_key || is_eps)) return;
  int32_t state = __ldcg(&B.table[slot].key);
  int4 st = __ldg(P.st + state);
  if (st.w == 0) return;
  double cost = dunkey(v.cost);
  *eps_count += static_cast<uint32_t>(st.w);
  for (int a = st.z; a < st.z + st.w; ++a) {
    int4 arc = __ldg(P.n_arc + a);
    double nc = cost + static_cast<double>(__int_as_float(arc.y));
    if (nc > cstar) continue;  // faster-decoder.cc:92
    eps_arrival(P, B, sh, arc.z, dkey(nc), static_cast<uint32_t>(a), slot, cstar_key, q_next,
                q_next_n);
  }
}

// Epsilon closure (faster-decoder.cc:59-119) followed by the commit of the
// frame: live table entries become the next token block in the arena, the
// table is wiped, and the block's min cost is recorded for the next GetCutoff.
template <int THREADS>
__device__ void lane_closure_and_commit(const Params &P, const LaneBuf &B, Shared &sh,
                                        LaneState &ls, double cstar) {
  const int tid = threadIdx.x;
  const unsigned long long cstar_key = dkey(cstar);
  const double inf = __longlong_as_double(0x7FF0000000000000ll);
  const long long t_begin = clock64();
  // ---- closure: sweep 0 expands every token, later sweeps the improved ones
  if (tid == 0) {
    sh.q_n[0] = 0;
    sh.q_n[1] = 0;
    sh.out_n = 0;
    sh.acc_eps = 0;
  }
  __syncthreads();
  uint32_t eps_count = 0;
  const uint32_t m0 = min(sh.list_n, P.lcap);
  __syncthreads();  // everyone holds m0 before the closure starts growing the list
  uint32_t *q0 = B.queue, *q1 = B.queue + P.qcap;
  for (uint32_t p = tid; p < m0; p += THREADS)
    expand_eps(P, B, sh, B.list[p], cstar_key, cstar, q0, &sh.q_n[0], &eps_count);
  __syncthreads();
  int cur = 0;
  long long sweeps = 1;
  while (true) {
    uint32_t qn = min(sh.q_n[cur], P.qcap);
    if (qn == 0 || sh.status != 0) break;
    __syncthreads();
    if (tid == 0) sh.q_n[cur ^ 1] = 0;
    __syncthreads();
    uint32_t *qc = cur ? q1 : q0, *qx = cur ? q0 : q1;
    for (uint32_t p = tid; p < qn; p += THREADS)
      expand_eps(P, B, sh, qc[p], cstar_key, cstar, qx, &sh.q_n[cur ^ 1], &eps_count);
    __syncthreads();
    cur ^= 1;
    ++sweeps;
  }
  __syncthreads();
  const long long t_mid = clock64();
  // ---- commit, pass 1: number the live entries
  const uint32_t m = min(sh.list_n, P.lcap);
  for (uint32_t p0 = 0; p0 < m; p0 += THREADS) {  // uniform trip count: full-warp ballots
    const uint32_t p = p0 + tid;
    uint32_t h = 0;
    bool live = false;
    if (p < m) {
      h = B.list[p];
      HVal v = ld_hval(&B.table[h].val);
      live = v.cost != kEmptyCost && (v.cost < cstar_key || (v.arg >> 63) != 0);
    }
    // warp-aggregated numbering: one shared-memory atomic per warp
    const uint32_t live_mask = __ballot_sync(0xFFFFFFFFu, live);
    uint32_t wbase = 0;
    if ((tid & 31) == 0 && live_mask) wbase = atomicAdd(&sh.out_n, __popc(live_mask));
    wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
    if (p < m) {
      uint32_t idx = kNoIdx;
      if (live) idx = wbase + __popc(live_mask & ((1u << (tid & 31)) - 1u));
      B.table[h].idx = idx;
    }
  }
  __syncthreads();
  const uint32_t n_new = sh.out_n;
  const uint32_t new_base = ls.arena_used;
  if (static_cast<long long>(new_base) + n_new > P.arena_cap) {
    if (tid == 0) sh.status |= kStatusArenaOverflow;
  }
  __syncthreads();
  const bool write_ok = (sh.status & kStatusArenaOverflow) == 0;
  // ---- commit, pass 2: write records, wipe the table
  double my_min = inf;
  int my_arg = -1;
  for (uint32_t p = tid; p < m; p += THREADS) {
    uint32_t h = B.list[p];
    uint32_t idx = B.table[h].idx;
    if (idx != kNoIdx && write_ok) {
      HVal v = ld_hval(&B.table[h].val);
      uint32_t arc = static_cast<uint32_t>(v.arg >> 32);
      uint32_t prev = static_cast<uint32_t>(v.arg);
      if (arc & kEpsFlag) prev = new_base + B.table[prev].idx;
      double c = dunkey(v.cost);
      B.a_cost[new_base + idx] = c;
      B.a_link[new_base + idx] = (static_cast<unsigned long long>(arc) << 32) | prev;
      B.a_state[new_base + idx] = __ldcg(&B.table[h].key);
      if (c < my_min) {
        my_min = c;
        my_arg = static_cast<int>(idx);
      }
    }
  }
  __syncthreads();  // every idx/key read above precedes the wipe below
  for (uint32_t p = tid; p < m; p += THREADS) {
    uint32_t h = B.list[p];
    ulonglong2 e;
    e.x = kEmptyCost;
    e.y = kEmptyArg;
    *reinterpret_cast<ulonglong2 *>(&B.table[h].val) = e;
    B.table[h].key = kEmptyKey;
  }
  double bmin;
  int barg;
  block_min_arg<THREADS>(my_min, my_arg, sh, &bmin, &barg);
  // accumulate counters
  eps_count = __reduce_add_sync(0xFFFFFFFFu, eps_count);
  if ((tid & 31) == 0 && eps_count) atomicAdd(&sh.acc_eps, eps_count);
  __syncthreads();
  if (tid == 0) {
    if (write_ok) {
      ls.tok_base = new_base;
      ls.n_tok = static_cast<int32_t>(n_new);
      ls.arena_used = new_base + n_new;
      ls.best_cost = bmin;
      ls.best_idx = barg;
    } else {
      ls.n_tok = 0;
      ls.best_cost = inf;
      ls.best_idx = -1;
    }
    ls.st_sweeps += sweeps;
    ls.st_eps_arcs += sh.acc_eps;
    ls.cyc_closure += t_mid - t_begin;
    ls.cyc_commit += clock64() - t_mid;
    sh.list_n = 0;
  }
  __syncthreads();
}

// Admitted emitting arcs are not recombined where they are found: ~97% of the
// arcs a frame visits fail the pruning test, so a warp iteration over 32 arcs
// admits about one, and recombining it in place makes 31 lanes wait for one
// lane's table round trips (profiles/r1_v1_ncu_summary.txt).  Instead each
// warp parks admitted arcs in a shared-memory queue (one native 32-bit
// shared-memory atomicAdd per admitted arc, nothing per rejected arc) and
// recombines 32 of them at a time, one per lane, so the table latencies overlap.
constexpr uint32_t kQueueCap = 192;  // 31 left over + at most 4 x 32 parked between checks

struct WarpQueue {
  unsigned long long nk[kQueueCap];   // ordered fp64 cost
  unsigned long long arg[kQueueCap];  // (arc << 32) | source token
  int32_t dst[kQueueCap];             // destination state
  uint32_t n;
  uint32_t pad[3];
};

__device__ __forceinline__ void queue_insert_one(const Params &P, const LaneBuf &B, Shared &sh,
                                                 const WarpQueue &q, uint32_t e) {
  const unsigned long long nk = q.nk[e];
  // the running cutoff may have tightened since the arc was parked
  const double cut_now =
      static_cast<double>(funkey(*reinterpret_cast<volatile uint32_t *>(&sh.cut_fkey)));
  if (!(dunkey(nk) < cut_now)) return;
  uint32_t h = table_slot(P, B, sh, q.dst[e]);
  if (h == kNoIdx) return;
  HVal mine;
  mine.cost = nk;
  mine.arg = q.arg[e];
  table_min(&B.table[h].val, mine);
}

// Warp-convergent: recombines full groups of 32 parked arcs.
__device__ __forceinline__ void queue_drain_full(const Params &P, const LaneBuf &B, Shared &sh,
                                                 WarpQueue &q) {
  uint32_t n = *reinterpret_cast<volatile uint32_t *>(&q.n);
  if (n < 32) return;  // q.n is only written by this warp: the value is warp-uniform
  __syncwarp();
  do {
    queue_insert_one(P, B, sh, q, n - 32 + (threadIdx.x & 31));
    n -= 32;
  } while (n >= 32);
  __syncwarp();
  if ((threadIdx.x & 31) == 0) q.n = n;
  __syncwarp();
}

// Four emitting arcs per thread (faster-decoder.cc:208-229): new_weight =
// (w + cost) + ac for all four first -- straight-line code the compiler can
// interleave -- then the rare admitted ones are parked.  The shared running
// cutoff is kept as a float rounded UP (one native 32-bit shared-memory
// atomicMin); it only filters.  The exact C* comes from the per-thread fp64
// minimum `my_min` reduced at the end of the frame: the arc with the globally
// smallest new_weight always passes the filter.  Slots with bit u of `vmask`
// clear hold a dummy arc (ilabel 1) and are ignored.
template <bool ROW_SMEM>
__device__ __forceinline__ void emit4(Shared &sh, WarpQueue &q, const float *row,
                                      const int4 (&ar)[4], uint32_t vmask, uint32_t a0,
                                      uint32_t a_stride, double tcost, uint32_t tok_abs,
                                      double ab, double &my_min) {
  const double cut_d =
      static_cast<double>(funkey(*reinterpret_cast<volatile uint32_t *>(&sh.cut_fkey)));
  double nw[4];
  uint32_t adm = 0;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float lp = ROW_SMEM ? row[ar[u].x - 1] : __ldg(row + ar[u].x - 1);
    nw[u] = (static_cast<double>(__int_as_float(ar[u].y)) + tcost) + static_cast<double>(-lp);
    if (nw[u] < cut_d) adm |= 1u << u;  // faster-decoder.cc:211 (filter; exact test at commit)
  }
  adm &= vmask;
  if (adm == 0) return;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    if (adm & (1u << u)) {
      const uint32_t e = atomicAdd(&q.n, 1u);
      q.nk[e] = dkey(nw[u]);
      q.arg[e] = (static_cast<unsigned long long>(a0 + a_stride * u) << 32) | tok_abs;
      q.dst[e] = ar[u].z;
      if (nw[u] < my_min) {
        my_min = nw[u];
        const uint32_t fk = fkey(__double2float_ru(nw[u] + ab));  // faster-decoder.cc:215-217
        if (fk < *reinterpret_cast<volatile uint32_t *>(&sh.cut_fkey)) atomicMin(&sh.cut_fkey, fk);
      }
    }
  }
}

// faster-decoder.cc:155-241 for one lane-frame.  Returns C*.
//
// Work mapping: a warp takes 32 tokens at a time.  Tokens with few emitting
// arcs (<= kSmallDeg, the bulk of a lexicon trie) are expanded by their own
// thread; tokens with many arcs (trie roots, H states) are expanded by the
// whole warp, 32 consecutive 16-byte arcs per load instruction.
constexpr uint32_t kSmallDeg = 8;

template <int THREADS, bool ROW_SMEM>
__device__ double lane_expand_emitting(const Params &P, const LaneBuf &B, Shared &sh,
                                       LaneState &ls, const float *row_g, float *s_row,
                                       WarpQueue *queues) {
  const int tid = threadIdx.x, lane = tid & 31;
  const double inf = __longlong_as_double(0x7FF0000000000000ll);
  const int n = ls.n_tok;
  const uint32_t base = ls.tok_base;
  const double *cost = B.a_cost + base;
  const int32_t *state = B.a_state + base;
  WarpQueue &q = queues[tid >> 5];
  const long long t_begin = clock64();

  // the log-prob row of this frame -> shared memory (decodable-ctc.cc:22-29)
  const float *row = row_g;
  if (ROW_SMEM) {
    for (int i = tid; i < P.cols; i += THREADS) s_row[i] = __ldg(row_g + i);
    row = s_row;
  }
  if (tid == 0) {
    sh.cut_fkey = fkey(__int_as_float(0x7F800000));
    sh.chunk = 0;
    sh.acc_emit = sh.acc_expanded = 0;
  }
  double wc;
  float abf;
  lane_cutoff<THREADS>(P, cost, n, ls.best_cost, sh, &wc, &abf);
  const double ab = static_cast<double>(abf);
  __syncthreads();

  // seed the running cutoff from the best token's arcs (faster-decoder.cc:176-189)
  double seed = inf;
  if (n > 0 && ls.best_cost < wc) {
    int4 st = __ldg(P.st + state[ls.best_idx]);
    for (int a = tid; a < st.y; a += THREADS) {
      int4 arc = __ldg(P.e_arc + st.x + a);
      const float lp = ROW_SMEM ? row[arc.x - 1] : __ldg(row + arc.x - 1);
      double nw = (static_cast<double>(__int_as_float(arc.y)) + ls.best_cost) +
                  static_cast<double>(-lp);
      seed = fmin(seed, nw);
    }
  }
  {
    double smin;
    int dummy;
    block_min_arg<THREADS>(seed, 0, sh, &smin, &dummy);
    if (tid == 0) sh.cut_fkey = fkey(__double2float_ru(smin + ab));
    __syncthreads();
  }
  if (tid == 0) sh.t_mark = clock64();

  uint32_t n_expanded = 0, n_arcs = 0;
  double my_min = inf;
  if (lane == 0) q.n = 0;
  __syncwarp();
  while (true) {
    uint32_t c = 0;
    if (lane == 0) c = atomicAdd(&sh.chunk, 1u);
    c = __shfl_sync(0xFFFFFFFFu, c, 0);
    const uint32_t i0 = c * 32u;
    if (i0 >= static_cast<uint32_t>(n)) break;
    const uint32_t i = i0 + lane;
    double tc = inf;
    uint32_t cnt = 0, beg = 0;
    if (i < static_cast<uint32_t>(n)) {
      tc = cost[i];
      if (tc < wc) {  // faster-decoder.cc:202
        int4 st = __ldg(P.st + state[i]);
        beg = static_cast<uint32_t>(st.x);
        cnt = static_cast<uint32_t>(st.y);
        ++n_expanded;
        n_arcs += cnt;
      }
    }
    const bool big = cnt > kSmallDeg;
    // (i) small tokens: each thread walks its own arcs, 4 at a time
    const uint32_t scnt = big ? 0u : cnt;
#pragma unroll 1
    for (uint32_t k0 = 0; k0 < kSmallDeg; k0 += 4) {
      if (!__any_sync(0xFFFFFFFFu, k0 < scnt)) break;
      int4 ar[4];
      uint32_t vmask = 0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        ar[u] = make_int4(1, 0, 0, 0);
        if (k0 + u < scnt) {
          ar[u] = __ldg(P.e_arc + beg + k0 + u);
          vmask |= 1u << u;
        }
      }
      emit4<ROW_SMEM>(sh, q, row, ar, vmask, beg + k0, 1u, tc, base + i, ab, my_min);
      queue_drain_full(P, B, sh, q);
    }
    // (ii) big tokens: the warp walks the arc range together, 4 x 32 arcs per step
    uint32_t bm = __ballot_sync(0xFFFFFFFFu, big);
    while (bm) {
      const int src = __ffs(bm) - 1;
      bm &= bm - 1;
      const uint32_t b = __shfl_sync(0xFFFFFFFFu, beg, src);
      const uint32_t cn = __shfl_sync(0xFFFFFFFFu, cnt, src);
      const double cst = __shfl_sync(0xFFFFFFFFu, tc, src);
      const uint32_t tok_abs = base + i0 + src;
#pragma unroll 1
      for (uint32_t j0 = 0; j0 < cn; j0 += 128) {
        int4 ar[4];
        uint32_t vmask = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t j = j0 + 32u * u + lane;
          ar[u] = make_int4(1, 0, 0, 0);
          if (j < cn) {
            ar[u] = __ldg(P.e_arc + b + j);
            vmask |= 1u << u;
          }
        }
        emit4<ROW_SMEM>(sh, q, row, ar, vmask, b + j0 + lane, 32u, cst, tok_abs, ab, my_min);
        queue_drain_full(P, B, sh, q);
      }
    }
  }
  // recombine what is still parked
  __syncwarp();
  if (lane < *reinterpret_cast<volatile uint32_t *>(&q.n)) queue_insert_one(P, B, sh, q, lane);
  n_expanded = __reduce_add_sync(0xFFFFFFFFu, n_expanded);
  n_arcs = __reduce_add_sync(0xFFFFFFFFu, n_arcs);
  if (lane == 0 && n_expanded) atomicAdd(&sh.acc_expanded, n_expanded);
  if (lane == 0 && n_arcs) atomicAdd(&sh.acc_emit, n_arcs);
  // exact C* = min(new_weight) + adaptive_beam (faster-decoder.cc:240), including the seed
  double bmin;
  int dummy2;
  block_min_arg<THREADS>(fmin(my_min, seed), 0, sh, &bmin, &dummy2);
  if (tid == 0) {
    const long long t_end = clock64();
    ls.cyc_cutoff += sh.t_mark - t_begin;
    ls.cyc_expand += t_end - sh.t_mark;
  }
  return bmin + ab;
}

// ------------------------------------------------------------------ kernels

template <int THREADS, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) kd_advance_kernel(Params P) {
  __shared__ Shared sh;
  __shared__ LaneState ls;
  __shared__ LaneBuf sB;  // per-lane base pointers live in shared memory, not registers
  const LaneBuf &B = sB;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  WarpQueue *queues = reinterpret_cast<WarpQueue *>(dyn_smem);
  float *s_row = reinterpret_cast<float *>(dyn_smem + (THREADS / 32) * sizeof(WarpQueue));
  const int tid = threadIdx.x;

  while (true) {
    if (tid == 0) sh.item = atomicAdd(P.work_counter, 1);
    __syncthreads();
    const int item = sh.item;
    if (item >= P.n_items) return;
    const AdvanceItem it = P.items[item];
    if (tid == 0) {
      sB = lane_buffers(P, it.lane);
      ls = P.lanes[it.lane];
      sh.status = ls.status;
      sh.list_n = 0;
    }
    __syncthreads();
    while (ls.frames_decoded < it.target && sh.status == 0) {
      const int frame = ls.frames_decoded;
      const float *row_g = it.logp + static_cast<size_t>(frame - it.offset) * P.cols;
      const int n_in = ls.n_tok;
      double cstar;
      if (P.row_in_smem)
        cstar = lane_expand_emitting<THREADS, true>(P, B, sh, ls, row_g, s_row, queues);
      else
        cstar = lane_expand_emitting<THREADS, false>(P, B, sh, ls, row_g, s_row, queues);
      lane_closure_and_commit<THREADS>(P, B, sh, ls, cstar);
      if (tid == 0) {
        ls.frames_decoded = frame + 1;
        ls.st_frames += 1;
        ls.st_tokens_in += n_in;
        ls.st_tokens_out += ls.n_tok;
        ls.st_emit_arcs += sh.acc_emit;
        ls.st_expanded += sh.acc_expanded;
        if (ls.n_tok > ls.st_max_tokens) ls.st_max_tokens = ls.n_tok;
      }
      __syncthreads();
    }
    if (tid == 0) {
      ls.status = sh.status;
      P.lanes[it.lane] = ls;
    }
    __syncthreads();
  }
}

// InitDecoding (faster-decoder.cc:42-56): start token with cost 0, epsilon
// closure under cutoff FLT_MAX, zero frames decoded.  items[i].lane = lane.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) kd_init_kernel(Params P) {
  __shared__ Shared sh;
  __shared__ LaneState ls;
  const int tid = threadIdx.x;
  const int item = blockIdx.x;
  if (item >= P.n_items) return;
  const int lane = P.items[item].lane;
  const LaneBuf B = lane_buffers(P, lane);
  if (tid == 0) {
    LaneState z;
    memset(&z, 0, sizeof(z));
    z.frames_decoded = 0;
    z.best_cost = __longlong_as_double(0x7FF0000000000000ll);
    z.best_idx = -1;
    ls = z;
    sh.status = 0;
    sh.list_n = 0;
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t h = table_slot(P, B, sh, P.start);
    HVal v;
    v.cost = dkey(0.0);
    v.arg = (static_cast<unsigned long long>(kNoArc) << 32) | kNoPrev;
    *reinterpret_cast<ulonglong2 *>(&B.table[h].val) = make_ulonglong2(v.cost, v.arg);
  }
  __syncthreads();
  lane_closure_and_commit<THREADS>(P, B, sh, ls, 3.4028234663852886e+38 /* FLT_MAX */);
  if (tid == 0) {
    ls.status = sh.status;
    ls.st_sweeps = 0;
    ls.st_eps_arcs = 0;
    ls.cyc_closure = ls.cyc_commit = 0;
    P.lanes[lane] = ls;
  }
}

// ReachedFinal + best-token selection + path length (faster-decoder.cc:347-402).
// Ties on the selection cost go to the lowest state id.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) kd_best_select_kernel(Params P) {
  __shared__ Shared sh;
  __shared__ int s_any_final;
  const int tid = threadIdx.x;
  const int lane = P.items[blockIdx.x].lane;
  const LaneBuf B = lane_buffers(P, lane);
  LaneState *L = P.lanes + lane;
  const int n = L->n_tok;
  const uint32_t base = L->tok_base;
  const double inf = __longlong_as_double(0x7FF0000000000000ll);
  if (tid == 0) s_any_final = 0;
  __syncthreads();
  int any = 0;
  for (int i = tid; i < n; i += THREADS) {
    float f = __ldg(P.fin + B.a_state[base + i]);
    if (B.a_cost[base + i] != inf && f != __int_as_float(0x7F800000)) any = 1;
  }
  if (any) atomicOr(&s_any_final, 1);
  __syncthreads();
  const int is_final = s_any_final;
  double bv = inf;
  int bs = -1;  // state id is the tie-break key; token index recovered below
  for (int i = tid; i < n; i += THREADS) {
    int s = B.a_state[base + i];
    double c = B.a_cost[base + i];
    double v = is_final ? c + static_cast<double>(__ldg(P.fin + s)) : c;
    bool take = is_final ? (v != inf) : true;
    if (take && (bs < 0 || v < bv || (v == bv && s < bs))) {
      bv = v;
      bs = s;
    }
  }
  double rv;
  int rs;
  // block_min_arg treats idx -1 as "none" (largest unsigned)
  block_min_arg<THREADS>(bs < 0 ? inf : bv, bs, sh, &rv, &rs);
  // a token with cost +inf and no competitor: the reduction cannot tell it
  // from "none"; such tokens never exist (arrivals need cost < cutoff).
  __shared__ uint32_t s_best_tok;
  if (tid == 0) s_best_tok = kNoIdx;
  __syncthreads();
  if (rs >= 0) {
    for (int i = tid; i < n; i += THREADS)
      if (B.a_state[base + i] == rs) s_best_tok = base + i;
  }
  __syncthreads();
  if (tid == 0) {
    L->bp_final = is_final;
    if (rs < 0 || s_best_tok == kNoIdx) {
      L->bp_ok = 0;
      L->bp_len = 0;
      L->bp_best_tok = kNoIdx;
      L->bp_best_state = -1;
      L->bp_final_w = 0.f;
    } else {
      long long len = 0;
      uint32_t t = s_best_tok;
      while (true) {
        unsigned long long link = B.a_link[t];
        uint32_t arc = static_cast<uint32_t>(link >> 32);
        if (arc == kNoArc) break;
        ++len;
        t = static_cast<uint32_t>(link);
      }
      L->bp_ok = 1;
      L->bp_len = len;
      L->bp_best_tok = s_best_tok;
      L->bp_best_state = rs;
      L->bp_final_w = __ldg(P.fin + rs);
    }
  }
}

// Writes the best path of lane items[b].lane in time order at out_off[b]
// (faster-decoder.cc:393-402: graph = arc weight, acoustic = float(cost -
// prev cost) - graph).
__global__ void kd_best_fill_kernel(Params P, const long long *out_off, int32_t *il, int32_t *ol,
                                    float *gw, float *aw) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.n_items) return;
  const int lane = P.items[b].lane;
  const LaneBuf B = lane_buffers(P, lane);
  const LaneState *L = P.lanes + lane;
  if (!L->bp_ok) return;
  long long pos = out_off[b] + L->bp_len - 1;
  uint32_t t = L->bp_best_tok;
  double c = B.a_cost[t];
  while (true) {
    unsigned long long link = B.a_link[t];
    uint32_t arc = static_cast<uint32_t>(link >> 32);
    if (arc == kNoArc) break;
    uint32_t prev = static_cast<uint32_t>(link);
    double pc = B.a_cost[prev];
    int4 a;
    int32_t ilab, olab;
    if (arc & kEpsFlag) {
      a = __ldg(P.n_arc + (arc & ~kEpsFlag));
      ilab = 0;
      olab = a.x;
    } else {
      a = __ldg(P.e_arc + arc);
      ilab = a.x;
      olab = a.w;
    }
    float tot = static_cast<float>(c - pc);
    float graph = __int_as_float(a.y);
    il[pos] = ilab;
    ol[pos] = olab;
    gw[pos] = graph;
    aw[pos] = tot - graph;
    --pos;
    t = prev;
    c = pc;
  }
}

}  // namespace kd

#endif  // KD_KERNELS_CUH_
