// Lean exact search for the BATCHED path: byte-string left operands (linear,
// epsilon-free, unit weights) against a frozen transducer whose arc weights are
// finite and non-negative.  One G-lane group (G = 32 or 16) per string.
//
// Same observable behaviour as csp_warp.cuh (the reference's pop order, ids in
// first-touch order, its relax/tie rule and final pick; compose-shortest-path.zig
// :55-61, :70-89, :91-144, :159-179) — what changes is the cost per pop:
//
//   * TABLE.  Either a DENSE direct-indexed table  e = (p * S + s) * 2 + filter
//     of 16-byte records {dist, id, prev} (no hashing, no probing, neighbouring
//     transducer states share DRAM sectors) when (len+1) * S * 2 records fit the
//     per-string budget, or the open-addressing HASH table of 32-byte slots.
//   * BACK-POINTER = predecessor id only.  The reference stores (prev, ilabel,
//     olabel, weight) (:44-49); for a linear left operand the arc is a function
//     of (prev tuple, tuple, dist[prev], dist[tuple]): it is the FIRST arc in
//     expansion order from prev's transducer state to the tuple's state, of the
//     tuple's kind (match / input-epsilon), whose relaxation reaches dist[tuple]
//     — later arcs of the same expansion never replace it (they would need a
//     strictly smaller distance, or on a tie smaller (ilabel, olabel), impossible
//     in frozen arc order).  The back-track recomputes it for the P path arcs
//     instead of storing it for all N tuples.
//   * FOLD, DONE AT UPLOAD.  Arcs of one transducer state with the same (ilabel,
//     nextstate) always relax the same compose tuple from the same source, and
//     relaxing them one by one in arc order ends in the same table state as one
//     relaxation with the smallest weight by the first of them (DESIGN.md
//     §exactness-fold; fl(c + w) is monotone in w, so min over arcs of the new
//     distance is the new distance of the min weight).  The search records
//     `sarc` (device_types.cuh) carry that: {ilabel, nextstate | dup, wmin} — one
//     16-byte load per lane per pop, no run-time grouping.  (Run-time grouping
//     was measured first: REDUX.MIN over per-target masks is serialised per mask
//     by the compiler — 46 % of all executed instructions; MATCH.ANY + a shuffle
//     gather still cost 38 % of the instructions and the longest stall.)
//   * READY SET.  Bitmap over discovery ids, 32 ids per word, in HBM; the line
//     (G words = 32*G ids) that holds the current minimum lives in shared memory
//     (the WINDOW) and a one-bit-per-line summary sits above it.  The reference's
//     pop order sweeps ids almost monotonically (measured: 99.5 % of pops stay in
//     the same 1024-id line on the headline workload), so a pop is one LDS +
//     ballot, and an insert is one shared or global atomic OR — no loops.
//   * FUTURE SET (tuples whose tentative distance is above the current level):
//     nothing is recorded while the first level runs (only the smallest pushed
//     distance, for the early stop).  If a second level is really needed, one
//     scan over all tuples builds a radix heap over the IEEE bit pattern of the
//     distances (64 buckets by the highest differing bit; entries are ids,
//     validated against the tuple's current distance when visited).
//   * LOCKSTEP GROUPS.  The kernel is one loop whose iteration is "one step of
//     my string" (fetch / one pop / finish) with a warp-wide reconvergence at the
//     top, so two 16-lane groups in a warp issue their pops together instead of
//     drifting into different loops.
//   * REGISTERS.  Everything that is constant for the launch (arena offsets,
//     capacities) is read from the kernel parameters in the constant bank at the
//     point of use, and the cold state (radix-heap cursors, best final) lives in
//     shared memory, so that the hot loop fits 64 registers (8 blocks per SM).
//   * EAGER SEMANTICS (BASELINE config 5, SearchParams::eager).  The reference's eager pair — compose()
//     (compose.zig:29-198: lattice states numbered in FIFO discovery order) then shortestPath()
//     (shortest-path.zig:18-139: ties broken by the smaller predecessor STATE id, first arc in arc order, best
//     final = smallest (total, state id)) — gives a result that, for non-negative weights, is a function of the
//     distance field and the FIFO numbering only: back[v] = the smallest-numbered tight predecessor (every
//     reachable state pops once and relaxes v; a tie is taken iff the popped state's number is smaller, also on
//     settled targets, :75-78).  So the kernel (1) runs the exact search above for the distances, then (2) a
//     BFS phase that pops lattice states in FIFO order (numbers them) and gives every target its FIRST tight
//     relaxer — which is the smallest-numbered one — as back-pointer, (3) back-tracks with the eager rule.
#pragma once
#include "csp_kernels.cuh"
#include "csp_warp.cuh"   // kChunkIds, kNoChunk, kMaxFastTuples, bucket_of

// Resident 128-thread blocks per SM the lean kernel is compiled for (register cap = 65536 / (128 * this)).
#ifndef FSTB_LEAN_MINBLOCKS
#define FSTB_LEAN_MINBLOCKS 8
#endif

namespace fstb200 {

// 16-byte record of the dense table; all-ones bytes = never touched (id == kNone).
struct __align__(16) DenseEnt { double dist; uint32_t id; uint32_t prev; };
static_assert(sizeof(DenseEnt) == 16, "DenseEnt");
// 32-byte slot of the hash table (one DRAM sector); key all-ones = empty.  The record half is 16-byte aligned.
struct __align__(32) LeanSlot { unsigned long long key; unsigned long long spare; double dist; uint32_t id; uint32_t prev; };
static_assert(sizeof(LeanSlot) == 32, "LeanSlot");

constexpr uint32_t kBfsFlag = 0x40000000u;   // record id field: numbered by the BFS phase (ids stay below 2^22)
constexpr uint32_t kLeanColdWords = 12;   // smem words of cold per-group state (see LeanCold)
enum LeanCold : uint32_t { kcChunkNext = 0, kcFreeHead = 1, kcOccLo = 2, kcOccHi = 3, kcHaveBest = 4, kcBestId = 5,
                           kcBestFwLo = 6, kcBestFwHi = 7, kcBestTotLo = 8, kcBestTotHi = 9 };

struct LeanLayout {
  uint64_t off_tab, off_keyof, off_l0, off_chunks, total;
  uint64_t tab_bytes, l0_bytes;
  uint32_t n1;           // summary words (one bit per window line)
  uint32_t smem_words;   // u32 words of shared memory per group
};
// `tab_entries`: dense = (max_len + 1) * S * 2 records of 16 B; hash = slots of 32 B.
__host__ __device__ inline LeanLayout lean_layout(int G, bool dense, uint64_t tab_entries, uint32_t tuple_cap, uint32_t chunk_cap,
                                                  bool crec = false) {
  LeanLayout L;
  auto al = [](uint64_t x) { return (x + 127) & ~127ull; };
  const uint32_t ids_per_line = 32u * (uint32_t)G;
  tuple_cap += 8;   // slack: the fast kernel checks the capacity once per step (a step creates at most 8 tuples)
  const uint32_t lines = (tuple_cap + ids_per_line - 1) / ids_per_line;
  L.n1 = (lines + 31) / 32;
  L.tab_bytes = al(tab_entries * (dense ? (crec ? 8ull : 16ull) : 32ull));
  L.l0_bytes = al((uint64_t)lines * G * 4);
  L.off_tab = 0;
  L.off_keyof = L.tab_bytes;
  L.off_l0 = L.off_keyof + al((uint64_t)tuple_cap * (dense ? 4 : 8));   // id -> compact key (dense) / 64-bit key (hash)
  L.off_chunks = L.off_l0 + L.l0_bytes;
  L.total = (L.off_chunks + (uint64_t)chunk_cap * 128 + 255) & ~255ull;
  // radix buckets (64 x uint2) + window + summary + cold state
  L.smem_words = (128 + (uint32_t)G + L.n1 + kLeanColdWords + 3) & ~3u;
  return L;
}

// Arena initialisation (layout change): table bytes 0xFF (dense: id == kNone; hash: key == empty), bitmap zero.
__global__ void lean_arena_init_kernel(uint8_t* arena, uint64_t stride, uint32_t n_arenas, uint64_t off_l0, uint64_t tab_bytes,
                                       uint64_t l0_bytes) {
  const uint64_t tw = tab_bytes / 16, bw = l0_bytes / 16, per = tw + bw, total = per * n_arenas;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t ar = i / per, w = i % per;
    uint8_t* base = arena + ar * stride;
    if (w < tw) reinterpret_cast<uint4*>(base)[w] = make_uint4(~0u, ~0u, ~0u, ~0u);
    else reinterpret_cast<uint4*>(base + off_l0)[w - tw] = make_uint4(0, 0, 0, 0);
  }
}

// Per-group context: two pointers.  Every other address is `base + constant-bank offset`.
struct LeanCtx {
  uint8_t* base;    // this group's arena in HBM
  uint32_t* sm;     // this group's shared memory: [0,128) buckets | [128,128+G) window | l1[n1] | cold[kLeanColdWords]
};
template <int G> struct LgG { static constexpr int v = G == 32 ? 5 : (G == 16 ? 4 : (G == 8 ? 3 : 2)); };

#define LEAN_KEYOF(p, c) ((c).base + (p).off_keyof)
#define LEAN_L0(p, c) (reinterpret_cast<uint32_t*>((c).base + (p).off_l0))
#define LEAN_CHUNKS(p, c) (reinterpret_cast<uint32_t*>((c).base + (p).off_chunks))
#define LEAN_BUCKET(c) (reinterpret_cast<uint2*>((c).sm))
#define LEAN_WIN(c) ((c).sm + 128)
#define LEAN_L1(c, G) ((c).sm + 128 + (G))
#define LEAN_COLD(p, c, G) ((c).sm + 128 + (G) + (p).n1)

// Collectives of the hot path.  HOT = true: executed by ALL 32 lanes of the warp with the full mask (one
// VOTE/SHFL instruction, uniform mask), each group extracting its own lanes; every lane of the warp must
// reach every HOT collective of an iteration (groups that have nothing to do pass `false`).  HOT = false:
// the group's own member mask — valid inside code only one group executes (cold paths), but the compiler
// serialises a warp instruction per distinct mask (measured: ~12 instructions per vote).
template <int G, bool HOT>
__device__ __forceinline__ unsigned lean_ballot(const Group<G>& g, bool pred) {
  if (HOT && G < 32) return (__ballot_sync(0xFFFFFFFFu, pred) >> g.base) & Group<G>::kBits;
  return g.ballot(pred);
}
template <int G, bool HOT, class T>
__device__ __forceinline__ T lean_shfl(const Group<G>& g, T v, int src) {
  if (HOT && G < 32) return __shfl_sync(0xFFFFFFFFu, v, (int)g.base + src);
  return g.shfl(v, src);
}
template <int G, bool HOT>
__device__ __forceinline__ void lean_sync(const Group<G>& g) {
  if (HOT && G < 32) __syncwarp(); else g.sync();
}

// Hot per-string state (registers).
struct LeanState {
  uint32_t n_tuples;
  uint32_t wline;                // line held in the window, kNone = none
  unsigned long long last;       // bit pattern of the current level distance
  unsigned long long future_min; // per lane: smallest distance pushed beyond the current level
  unsigned long long relax_calls;   // PER-LANE partial count (summed over the warp when the kernel ends)
  bool low_pending;              // a ready id below the window was inserted
  bool overflow;                 // tuple capacity exhausted -> retry with a larger arena
  bool heap_overflow;            // radix-heap pool exhausted -> retry with a deeper pool
  bool sorted;                   // the radix heap exists (a second level was needed)
  bool bfs_started;              // the id -> key array holds BFS numbers (a full table reset is needed after an abort)
  bool stuck;                    // safety valve tripped (see the kernel loop)
  bool wide;                     // PER-LANE: a distance did not fit the compact record (table kind 2) -> retry with 16-byte records
  uint32_t occ;                  // eager BFS phase: records in the table = search tuples + tuples first met by the BFS
};

// Tuple key of the lean path as a (P, SF) pair: P = string position, SF = (transducer state << 1) | filter
// (filter is 0 or 1 for an epsilon-free left operand; states < 2^31 is checked at upload).
//   hash table:  64-bit key (P << 32) | SF.
//   dense table: index lean_dense_pos(P, SF) (position-major or diagonal rows); id -> key array holds the COMPACT 32-bit form
//                (P << key_sbits) | SF  (fits: the dense table has < 2^30 records).
// Table kind (template parameter DENSE of everything below): 0 = hash table of 32-byte slots, 1 = dense table of
// 16-byte records {dist f64, id, prev}, 2 = dense table of COMPACT 8-byte records  dist:20 | prev:22 | id:22  (most
// significant first: the relax rule "take iff (new dist, popped id) < (dist, prev)" is ONE unsigned compare of the top 42 bits)  for
// transducers whose weights are all small non-negative integers (every distance is then an integer, exact in
// both forms): half the table bytes per string, so more strings fit HBM and a DRAM sector holds four records.
// All-ones = never touched in every kind.  A distance that does not fit 20 bits marks the string for a retry with
// 16-byte records (LeanState::wide).  Kind 3 = kind 2 for the EAGER kernels (one id bit holds the BFS flag).
constexpr uint32_t kCrecNone = 0x3FFFFFu;          // 22-bit id / prev: none
constexpr double kCrecMaxDist = 1048574.0;         // 20-bit distance; 0xFFFFF = +inf (a lattice state the search never reached, eager BFS)
constexpr uint32_t kCrecBfsBit = 0x200000u;        // table kind 3 (eager): bit 21 of the id field = numbered by the BFS phase, ids < 2^21 - 1

template <int DENSE>
__device__ __forceinline__ void lean_keyof_store(const SearchParams& p, const LeanCtx& c, uint32_t id, uint32_t P, uint32_t SF) {
  if (DENSE) reinterpret_cast<uint32_t*>(LEAN_KEYOF(p, c))[id] = (P << p.key_sbits) | SF;
  else reinterpret_cast<unsigned long long*>(LEAN_KEYOF(p, c))[id] = ((unsigned long long)P << 32) | SF;
}
template <int DENSE>
__device__ __forceinline__ void lean_keyof_load(const SearchParams& p, const LeanCtx& c, uint32_t id, uint32_t& P, uint32_t& SF) {
  if (DENSE) {
    const uint32_t k = reinterpret_cast<const uint32_t*>(LEAN_KEYOF(p, c))[id];
    P = k >> p.key_sbits; SF = k & ((1u << p.key_sbits) - 1u);
  } else {
    const unsigned long long k = reinterpret_cast<const unsigned long long*>(LEAN_KEYOF(p, c))[id];
    P = (uint32_t)(k >> 32); SF = (uint32_t)k;
  }
}

// ── table access ──
// Dense index: P * 2S + ((state << 1) | filter) — the two filter variants of a (position, state) pair share a
// sector.  (A filter-major index, which makes the match targets of one expansion contiguous, was measured:
// L1 hit rate 55 % -> 43 %, DRAM reads x2.8; the variants are touched close together in time.)
__device__ __forceinline__ uint32_t lean_dense_pos(const SearchParams& p, uint32_t P, uint32_t SF) {
  return SF * p.pos_h + ((SF & 1u) ? p.pos_k : 0u) + p.pos_c + P * p.pos_m2;   // see SearchParams::pos_h
}
// Hash of a tuple key for the open-addressing table: two odd multipliers and an xorshift-multiply finisher (the slot is
// taken from the HIGH bits).  Keys are structured (small positions, clustered state numbers).  Against the 64-bit
// finalizer used before (a quarter of the instructions): WeText-style config 4 451 k -> 472 k strings/s, ambiguous
// len 96 forced onto the hash table 987 k -> 966 k.
__device__ __forceinline__ uint32_t lean_hash(uint32_t P, uint32_t SF) {
  uint32_t h = SF * 0x9E3779B1u + P * 0x85EBCA6Bu;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 13;
  return h;
}
// Find the record of key (P, SF): position and contents; id == kNone <=> not present (hash: `pos` is then
// the empty slot that ended the probe — pass it to lean_claim before storing).
template <int DENSE>
__device__ __forceinline__ void lean_lookup(const SearchParams& p, const LeanCtx& c, uint32_t P, uint32_t SF, uint32_t& pos, double& dist,
                                            uint32_t& id, uint32_t& prev) {
  if (DENSE >= 2) {
    pos = lean_dense_pos(p, P, SF);
    const uint2 v = *reinterpret_cast<const uint2*>(c.base + (uint64_t)pos * 8);   // x = low word: prev:10 low bits | id:22 ; y = dist:20 | prev:12 high bits
    const uint32_t i = v.x & kCrecNone, pr = ((v.y & 0xFFFu) << 10) | (v.x >> 22);
    id = i == kCrecNone ? kNone : (DENSE == 3 ? ((i & (kCrecBfsBit - 1u)) | ((i & kCrecBfsBit) ? kBfsFlag : 0u)) : i);
    prev = pr == kCrecNone ? kNone : pr;
    dist = (i == kCrecNone || (v.y >> 12) == 0xFFFFFu) ? d_inf() : (double)(v.y >> 12);
  } else if (DENSE) {
    pos = lean_dense_pos(p, P, SF);
    const uint4 v = *reinterpret_cast<const uint4*>(c.base + (uint64_t)pos * 16);
    dist = __hiloint2double((int)v.y, (int)v.x); id = v.z; prev = v.w;
  } else {
    const unsigned long long K = ((unsigned long long)P << 32) | SF;
    const LeanSlot* tab = reinterpret_cast<const LeanSlot*>(c.base);
    const uint32_t cap = (uint32_t)p.tab_entries;
    uint32_t i = (uint32_t)(((unsigned long long)lean_hash(P, SF) * cap) >> 32);
    for (;;) {
      const unsigned long long k = tab[i].key;
      if (k == kEmptyKey) { pos = i; dist = d_inf(); id = kNone; prev = kNone; return; }
      if (k == K) {
        pos = i;
        const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(&tab[i]) + 16);
        dist = __hiloint2double((int)v.y, (int)v.x); id = v.z; prev = v.w;
        return;
      }
      if (++i == cap) i = 0;
    }
  }
}
// Hash table: claim a slot for the new key, starting at the empty position the probe found (other
// lanes of the group insert other keys concurrently).  Dense table: nothing to do.
template <int DENSE>
__device__ __forceinline__ uint32_t lean_claim(const SearchParams& p, const LeanCtx& c, uint32_t P, uint32_t SF, uint32_t pos) {
  if (DENSE) return pos;
  const unsigned long long K = ((unsigned long long)P << 32) | SF;
  LeanSlot* tab = reinterpret_cast<LeanSlot*>(c.base);
  const uint32_t cap = (uint32_t)p.tab_entries;
  for (;;) {
    if (atomicCAS(&tab[pos].key, kEmptyKey, K) == kEmptyKey) return pos;
    if (++pos == cap) pos = 0;
  }
}
// Write a record.  The hash variant rewrites the key half too: plain stores keep this SM's L1 copy of the
// sector consistent with what later plain-load probes must see (the claiming CAS acts on L2 only).
template <int DENSE>
__device__ __forceinline__ void lean_store(const LeanCtx& c, uint32_t pos, uint32_t P, uint32_t SF, double dist, uint32_t id, uint32_t prev) {
  if (DENSE >= 2) {
    const uint32_t d = d_isinf(dist) ? 0xFFFFFu : __double2uint_rn(fmin(dist, kCrecMaxDist)), pr = prev & kCrecNone;   // kNone -> kCrecNone
    const uint32_t i = DENSE == 3 ? ((id & (kCrecBfsBit - 1u)) | ((id & kBfsFlag) ? kCrecBfsBit : 0u)) : id;
    *reinterpret_cast<uint2*>(c.base + (uint64_t)pos * 8) = make_uint2((pr << 22) | (i & kCrecNone), (d << 12) | ((pr >> 10) & 0xFFFu));
    return;
  }
  const uint4 v = make_uint4((uint32_t)__double2loint(dist), (uint32_t)__double2hiint(dist), id, prev);
  if (DENSE) {
    *reinterpret_cast<uint4*>(c.base + (uint64_t)pos * 16) = v;
  } else {
    uint4* sl = reinterpret_cast<uint4*>(c.base + (uint64_t)pos * 32);
    sl[0] = make_uint4(SF, P, 0u, 0u);
    sl[1] = v;
  }
}
template <int DENSE>
__device__ __forceinline__ double lean_dist_of_id(const SearchParams& p, const LeanCtx& c, uint32_t id) {
  uint32_t P, SF, pos, i2, pr; double d;
  lean_keyof_load<DENSE>(p, c, id, P, SF);
  lean_lookup<DENSE>(p, c, P, SF, pos, d, i2, pr);
  return d;
}

// ── ready set ──
// Insert ids (collective; one atomic OR per inserting lane).
template <int G, bool HOT>
__device__ __forceinline__ void lean_ready_insert(const SearchParams& p, const Group<G>& g, const LeanCtx& c, LeanState& st, bool active,
                                                  uint32_t id) {
  bool low = false;
  if (active) {
    const uint32_t line = id >> (5 + LgG<G>::v);
    const uint32_t bit = 1u << (id & 31u);
    if (line == st.wline) {
      atomicOr(&LEAN_WIN(c)[(id >> 5) & (G - 1)], bit);
    } else {
      atomicOr(&LEAN_L0(p, c)[id >> 5], bit);
      atomicOr(&LEAN_L1(c, G)[line >> 5], 1u << (line & 31u));
      low = st.wline != kNone && line < st.wline;
    }
  }
  if (lean_ballot<G, HOT>(g, low)) st.low_pending = true;   // the vote also orders the atomics before the next window read
  lean_sync<G, HOT>(g);
}
// Write the window back (it holds ids above a newly inserted smaller one).
template <int G>
__device__ __forceinline__ void lean_window_evict(const SearchParams& p, const Group<G>& g, const LeanCtx& c, LeanState& st) {
  if (st.wline != kNone) {
    const uint32_t w = LEAN_WIN(c)[g.lane];
    if (w) __stcg(&LEAN_L0(p, c)[st.wline * G + g.lane], w);
    if (g.any(w != 0)) { if (g.lane == 0) LEAN_L1(c, G)[st.wline >> 5] |= 1u << (st.wline & 31u); }
    LEAN_WIN(c)[g.lane] = 0;
    st.wline = kNone;
  }
  g.sync();
}
// Load the lowest non-empty line into the window.  False: the ready set is empty.
template <int G, int DENSE>
__device__ __forceinline__ bool lean_window_next(const SearchParams& p, const Group<G>& g, const LeanCtx& c, LeanState& st) {
  const uint32_t words = min(p.n1, (st.n_tuples >> (10 + LgG<G>::v)) + 1u);
  uint32_t* l1 = LEAN_L1(c, G);
  for (uint32_t base = 0; base < words; base += G) {
    const uint32_t i = base + g.lane;
    const uint32_t v = i < words ? l1[i] : 0u;
    const unsigned bal = g.ballot(v != 0);
    if (bal) {
      const int src = __ffs(bal) - 1;
      const uint32_t vv = g.shfl(v, src);
      const uint32_t line = ((base + src) << 5) + (__ffs(vv) - 1);
      if ((int)g.lane == src) l1[i] = vv & (vv - 1);
      st.wline = line;
      uint32_t* wp = &LEAN_L0(p, c)[line * G + g.lane];
      const uint32_t w = __ldcg(wp);
      if (w) __stcg(wp, 0u);
      LEAN_WIN(c)[g.lane] = w;
      // the keys of this line are read one pop at a time: pull them into L1 now (32*G ids, 4 or 8 B each)
      constexpr uint32_t kKeyBytes = DENSE ? 4u : 8u;
      const char* kp = reinterpret_cast<const char*>(LEAN_KEYOF(p, c)) + (uint64_t)line * 32 * G * kKeyBytes + (size_t)g.lane * 128;
#pragma unroll
      for (uint32_t r = 0; r < kKeyBytes / 4; r++) asm volatile("prefetch.global.L1 [%0];" ::"l"(kp + (size_t)r * G * 128));
      g.sync();
      return true;
    }
  }
  st.wline = kNone;
  return false;
}

// ── future set (radix heap; its cursors live in shared memory, lane 0 writes, everyone reads) ──
template <int G>
__device__ __forceinline__ unsigned long long lean_occupied(const SearchParams& p, const LeanCtx& c) {
  const uint32_t* cold = LEAN_COLD(p, c, G);
  return ((unsigned long long)cold[kcOccHi] << 32) | cold[kcOccLo];
}
template <int G>
__device__ __forceinline__ void lean_set_occupied(const SearchParams& p, const Group<G>& g, const LeanCtx& c, unsigned long long v) {
  uint32_t* cold = LEAN_COLD(p, c, G);
  if (g.lane == 0) { cold[kcOccLo] = (uint32_t)v; cold[kcOccHi] = (uint32_t)(v >> 32); }
}
// Take a chunk from the free list or the pool (group-uniform result).
template <int G>
__device__ __forceinline__ uint32_t lean_chunk_alloc(const SearchParams& p, const Group<G>& g, const LeanCtx& c, LeanState& st) {
  uint32_t* cold = LEAN_COLD(p, c, G);
  const uint32_t fh = cold[kcFreeHead], cn = cold[kcChunkNext];
  uint32_t ch;
  g.sync();
  if (fh != kNoChunk) {
    ch = fh;
    const uint32_t nx = LEAN_CHUNKS(p, c)[(uint64_t)ch * 32];
    if (g.lane == 0) cold[kcFreeHead] = nx;
  } else if (cn < p.heap_cap) {
    ch = cn;
    if (g.lane == 0) cold[kcChunkNext] = cn + 1;
  } else {
    st.heap_overflow = true; ch = 0;
  }
  g.sync();
  return ch;
}
template <int G>
__device__ __forceinline__ void lean_chunk_free(const SearchParams& p, const Group<G>& g, const LeanCtx& c, uint32_t ch) {
  uint32_t* cold = LEAN_COLD(p, c, G);
  const uint32_t fh = cold[kcFreeHead];
  g.sync();
  if (g.lane == 0) { LEAN_CHUNKS(p, c)[(uint64_t)ch * 32] = fh; cold[kcFreeHead] = ch; }
  g.sync();
}
// Radix-heap chunk (128 bytes): word 0 = next chunk, words 1..10 = ids, words 12..31 = the distance keys the ids were
// pushed with (8 bytes each).  Keeping the key beside the id lets a level advance classify the entries of a bucket
// without a table lookup per entry; an entry is checked against the tuple's current distance only when it is about
// to enter the ready set (a stale entry — the tuple was lowered again later — is dropped there).
constexpr uint32_t kLeanChunkIds = 10;
__device__ __forceinline__ unsigned long long* lean_chunk_keys(uint32_t* chunks, uint32_t ch) {
  return reinterpret_cast<unsigned long long*>(chunks + (uint64_t)ch * 32 + 12);
}
// Append (id, key) entries to radix buckets (collective).  `b` in 1..64 for active lanes.
template <int G>
__device__ __forceinline__ void lean_bucket_push(const SearchParams& p, const Group<G>& g, const LeanCtx& c, LeanState& st, bool active,
                                                 uint32_t id, uint32_t b, unsigned long long key) {
  unsigned m = g.ballot(active);
  const unsigned lt = g.lt_mask();
  uint32_t* chunks = LEAN_CHUNKS(p, c);
  unsigned long long occ = lean_occupied<G>(p, c);
  while (m) {
    const int first = __ffs(m) - 1;
    const uint32_t bb = g.shfl(b, first);
    const bool mine = active && b == bb;
    const unsigned same = g.ballot(mine);
    m &= ~same;
    const uint32_t k = __popc(same), rank = __popc(same & lt);
    const bool empty = !((occ >> (bb - 1)) & 1ull);
    const uint2 hb = LEAN_BUCKET(c)[bb - 1];
    uint32_t head = empty ? kNoChunk : hb.x, cnt = empty ? kLeanChunkIds : hb.y;
    const uint32_t space = kLeanChunkIds - cnt;
    if (mine && rank < space) { chunks[(uint64_t)head * 32 + 1 + cnt + rank] = id; lean_chunk_keys(chunks, head)[cnt + rank] = key; }
    uint32_t left = k > space ? k - space : 0, done = k - left;
    if (left == 0) cnt += k;
    while (left > 0) {
      const uint32_t ch = lean_chunk_alloc<G>(p, g, c, st);
      if (st.heap_overflow) return;
      const uint32_t take = left < kLeanChunkIds ? left : kLeanChunkIds;
      if (g.lane == 0) chunks[(uint64_t)ch * 32] = head;
      if (mine && rank >= done && rank < done + take) { chunks[(uint64_t)ch * 32 + 1 + (rank - done)] = id; lean_chunk_keys(chunks, ch)[rank - done] = key; }
      head = ch; cnt = take; done += take; left -= take;
    }
    g.sync();
    if (g.lane == 0) LEAN_BUCKET(c)[bb - 1] = make_uint2(head, cnt);
    occ |= 1ull << (bb - 1);
    g.sync();
  }
  lean_set_occupied<G>(p, g, c, occ);
  g.sync();
}
template <int G>
__device__ __forceinline__ unsigned long long lean_group_min_u64(const Group<G>& g, unsigned long long v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(g.mask, v, o, G); v = t < v ? t : v; }
  return v;
}

// Build the radix heap: every tuple whose distance is above the level just finished is unsettled and
// belongs to the future set (one scan, once per string, only if a second level is needed).
template <int G, int DENSE>
__device__ __forceinline__ void lean_build_radix(const SearchParams& p, const Group<G>& g, const LeanCtx& c, LeanState& st) {
  st.sorted = true;
  const uint32_t n = st.n_tuples;
  for (uint32_t base = 0; base < n && !st.heap_overflow; base += G) {
    const uint32_t j = base + g.lane;
    bool valid = false; unsigned long long k = 0;
    if (j < n) {
      k = (unsigned long long)__double_as_longlong(lean_dist_of_id<DENSE>(p, c, j));
      valid = k > st.last;
    }
    if (g.any(valid)) lean_bucket_push<G>(p, g, c, st, valid, j, valid ? bucket_of(k, st.last) : 1u, k);
  }
}

// Advance to the next distance level; false = the search is finished (no valid entry left, or no
// remaining tuple can change the result).
template <int G, int DENSE>
__device__ __forceinline__ bool lean_advance_level(const SearchParams& p, const Group<G>& g, const LeanCtx& c, LeanState& st) {
  const uint32_t* cold = LEAN_COLD(p, c, G);
  const bool have_best = cold[kcHaveBest] != 0;
  const double best_total = __hiloint2double((int)cold[kcBestTotHi], (int)cold[kcBestTotLo]);
  if (!st.sorted) {
    const unsigned long long fm = lean_group_min_u64<G>(g, st.future_min);
    if (fm == ~0ull) return false;   // nothing was ever pushed beyond the levels already done
    // lower bound of every remaining distance: cannot reach or tie the best total -> done
    if (!p.exhaustive && have_best && __longlong_as_double((long long)fm) > best_total) return false;
    lean_build_radix<G, DENSE>(p, g, c, st);
    if (st.heap_overflow) return false;
  }
  uint32_t* chunks = LEAN_CHUNKS(p, c);
  for (;;) {
    unsigned long long occ = lean_occupied<G>(p, c);
    if (occ == 0) return false;
    const uint32_t b0 = __ffsll((long long)occ);   // bucket number 1..64
    const uint2 hb = LEAN_BUCKET(c)[b0 - 1];
    // pass 1: smallest key in the bucket (the keys stored with the ids: no table access; a stale entry can only make
    // the level a distance at which nothing turns out to be ready, and its key is never below a valid one's)
    unsigned long long m = ~0ull;
    {
      uint32_t ch = hb.x, cnt = hb.y;
      while (ch != kNoChunk) {
        const uint32_t next = chunks[(uint64_t)ch * 32];
        const unsigned long long* kp = lean_chunk_keys(chunks, ch);
        for (uint32_t o = g.lane; o < cnt; o += G) { const unsigned long long k = kp[o]; if (k < m) m = k; }
        ch = next; cnt = kLeanChunkIds;
      }
      m = lean_group_min_u64<G>(g, m);
    }
    g.sync();
    lean_set_occupied<G>(p, g, c, occ & ~(1ull << (b0 - 1)));
    g.sync();
    if (!p.exhaustive && have_best && __longlong_as_double((long long)m) > best_total) return false;
    // pass 2: entries at the new level key m enter the ready set if they are current (the tuple still has that
    // distance); the others move to the bucket of their key relative to m
    st.last = m;
    uint32_t ch = hb.x, cnt = hb.y;
    while (ch != kNoChunk && !st.heap_overflow) {
      const uint32_t* cp = chunks + (uint64_t)ch * 32;
      const unsigned long long* kp = lean_chunk_keys(chunks, ch);
      const uint32_t next = cp[0];
      for (uint32_t o = 0; o < cnt && !st.heap_overflow; o += G) {
        bool have = false, ready = false; uint32_t id = 0; unsigned long long k = 0;
        if (o + g.lane < cnt) {
          have = true; id = cp[1 + o + g.lane]; k = kp[o + g.lane];
          if (k == m) ready = (unsigned long long)__double_as_longlong(lean_dist_of_id<DENSE>(p, c, id)) == m;
        }
        lean_ready_insert<G, false>(p, g, c, st, ready, id);
        const bool tb = have && k != m;
        if (g.any(tb)) lean_bucket_push<G>(p, g, c, st, tb, id, tb ? bucket_of(k, m) : 1u, k);
      }
      lean_chunk_free<G>(p, g, c, ch);   // recycle
      ch = next; cnt = kLeanChunkIds;
    }
    return !st.heap_overflow;
  }
}

// ── one relaxation per lane ──
// `active` lanes hold DISTINCT targets (the static search records fold parallel arcs, see
// device_types.cuh `sarc`); `my_cnt` = this lane's share of the reference's relax calls (all arcs of the
// expansion, folded ones included).
// `first`: lanes that precede the other lanes in the reference's expansion order (match arcs
// :182-202 before input-epsilon arcs :254-278); only used to number newly discovered tuples.
// HOT: see lean_ballot — every lane of the warp calls, groups without work pass active = false.
template <int G, int DENSE, bool HOT, bool EAGER = true>
__device__ __forceinline__ void lean_relax(const SearchParams& p, const Group<G>& g, const LeanCtx& c, LeanState& st, uint32_t cur_id,
                                           uint32_t my_cnt, bool active, uint32_t P, uint32_t SF, double wmin, unsigned first,
                                           bool bfs = false, double bfs_dist = 0.0) {
  // NOTE: the search step and the BFS step of the eager semantics share every collective CALL SITE below —
  // groups of one warp can be in different modes, and a warp-wide collective only matches itself.
  st.relax_calls += my_cnt;
  uint32_t pos = 0, old_id = 0, old_prev = kNone; double old_dist = d_inf();
  if (active) lean_lookup<DENSE>(p, c, P, SF, pos, old_dist, old_id, old_prev);
  // smallest new distance over the parallel arcs (fl(c + w) is monotone in w)
  const double nd = (bfs ? bfs_dist : __longlong_as_double((long long)st.last)) + wmin;
  if (DENSE >= 2 && active && !bfs && nd > kCrecMaxDist) st.wide = true;   // (BFS phase: distances are final, +inf = unreached)
  const bool untouched = active && old_id == kNone;
  // search: a target is new when it has no record; BFS: when it has no BFS number yet (compose.zig:77-91)
  const bool is_new = bfs ? (active && (untouched || !(old_id & kBfsFlag))) : untouched;
  const unsigned newmask = lean_ballot<G, HOT>(g, is_new);
  const uint32_t n_new = __popc(newmask);
  // BFS: the table holds the search's tuples AND the ones the BFS met first, while n_tuples restarts at 1 — bound
  // the records (occ >= n_tuples), or a hash table sized for tuple_cap fills up and probing never ends
  uint32_t n_unt = 0;
  if (EAGER) n_unt = __popc(lean_ballot<G, HOT>(g, bfs && untouched));
  if ((EAGER && bfs ? st.occ + n_unt : st.n_tuples + n_new) > p.tuple_cap) { st.overflow = true; active = false; }   // nothing was written yet
  bool lowered = false;
  uint32_t my_id = old_id;
  if (active) {
    if (is_new) {
      const unsigned lt = g.lt_mask();
      const bool in_first = (first >> g.lane) & 1u;
      const uint32_t rank = in_first ? __popc(newmask & first & lt) : (__popc(newmask & first) + __popc(newmask & ~first & lt));
      my_id = st.n_tuples + rank;   // discovery order == reference expansion order (:80-87)
      lean_keyof_store<DENSE>(p, c, my_id, P, SF);
      // hash table: the empty slot the probe found is claimed with a CAS only if another lane of the group inserts in
      // this step too (only this group writes to this arena; a lone insert cannot race: lean_store writes the key)
      if (untouched && (DENSE || ((EAGER && bfs) ? n_unt : n_new) > 1u)) pos = lean_claim<DENSE>(p, c, P, SF, pos);
    }
    if (bfs) {
      // eager semantics: the first tight relaxer in FIFO order is the smallest-numbered tight predecessor
      // (shortest-path.zig:75-78); a target the search never reached (dist +inf) is never tight
      const bool tight = !d_isinf(old_dist) && nd == old_dist;
      if (is_new) lean_store<DENSE>(c, pos, P, SF, untouched ? d_inf() : old_dist, my_id | kBfsFlag, tight ? cur_id : kNone);
      else if (old_prev == kNone && tight) lean_store<DENSE>(c, pos, P, SF, old_dist, old_id, cur_id);
    } else {
      lowered = is_new || nd < old_dist;                                                         // :109-114, :137-142
      const bool take = lowered || (nd == old_dist && (old_prev == kNone || cur_id < old_prev)); // :115-126
      if (take) lean_store<DENSE>(c, pos, P, SF, nd, my_id, cur_id);
    }
  }
  if (!st.overflow) { st.n_tuples += n_new; if (EAGER) st.occ += n_unt; }
  // queue: ready set if at the current level, else future set (the BFS step inserts nothing)
  const unsigned long long k = (unsigned long long)__double_as_longlong(nd);
  lean_ready_insert<G, HOT>(p, g, c, st, lowered && k == st.last, my_id);
  const bool to_future = lowered && k != st.last;
  if (to_future && k < st.future_min) st.future_min = k;
  if (st.sorted && !bfs) {   // cold: only this group is in here
    if (g.any(to_future)) lean_bucket_push<G>(p, g, c, st, to_future, my_id, to_future ? bucket_of(k, st.last) : 1u, k);
  }
}

// Recover the arc of the step prev -> tuple (see file header).  Returns false if none fits (internal error).
__device__ inline bool lean_recover_arc(const DevFstView& F, const LhsBytes& lhs, uint32_t pu, uint32_t su, uint32_t pv, uint32_t sv,
                                        double du, double dv, PoolArc& out) {
  const uint4 rec = __ldg(&F.state_rec[su]);
  uint32_t lo, hi, il = 0;
  const bool match = pv == pu + 1;
  if (match) {
    il = (uint32_t)__ldg(lhs.s + pu) + 1u;
    uint32_t l = rec.y, h = rec.z;   // labels >= 1 start after the epsilon prefix
    while (l < h) { const uint32_t m = l + (h - l) / 2; if (__ldg(F.ilabel + m) < il) l = m + 1; else h = m; }
    lo = l; hi = rec.z;
  } else {
    lo = rec.x; hi = rec.y;
  }
  for (uint32_t arc = lo; arc < hi; arc++) {
    if (match && __ldg(F.ilabel + arc) != il) break;
    const uint4 pl = __ldg(&F.payload[arc]);
    if (pl.y != sv) continue;
    const double w2 = __hiloint2double((int)pl.w, (int)pl.z);
    const double ew = match ? 0.0 + w2 : w2;   // :189-198 a1.weight (x) a2.weight with a1.weight == One; :258-268 a2.weight
    if (du + ew == dv) { out.ilabel = il; out.olabel = pl.x; out.weight = ew; return true; }
  }
  return false;
}

// Start a string: initial tuple id 0, dist One, ready at level 0 (compose-shortest-path.zig:146-153).
// Precondition (invariant between strings): table untouched-state, ready bitmap, window and summary zero.
template <int G, int DENSE>
__device__ __forceinline__ void lean_begin(const SearchParams& p, const Group<G>& g, const LeanCtx& c, LeanState& st) {
  st.n_tuples = 1; st.wline = 0; st.relax_calls = 0; st.last = 0; st.future_min = ~0ull;
  st.low_pending = false; st.overflow = false; st.heap_overflow = false; st.sorted = false; st.bfs_started = false; st.stuck = false; st.occ = 0; st.wide = false;
  uint32_t* cold = LEAN_COLD(p, c, G);
  if (g.lane == 0) {
    const uint32_t SF = p.fst.start << 1;
    uint32_t pos, id, prev; double d;
    lean_lookup<DENSE>(p, c, 0u, SF, pos, d, id, prev);
    pos = lean_claim<DENSE>(p, c, 0u, SF, pos);
    lean_store<DENSE>(c, pos, 0u, SF, 0.0, 0u, kNone);
    lean_keyof_store<DENSE>(p, c, 0u, 0u, SF);
    cold[kcChunkNext] = 0; cold[kcFreeHead] = kNoChunk; cold[kcOccLo] = 0; cold[kcOccHi] = 0; cold[kcHaveBest] = 0;
  }
  LEAN_WIN(c)[g.lane] = g.lane == 0 ? 1u : 0u;   // window = line 0, bit 0
  g.sync();
}

// One step: pop the smallest ready id and expand it, or (rarely) switch the window / advance the level.
// Called by ALL lanes of the warp in every iteration in which any of its groups is running (`running` =
// this lane's group is); the hot collectives are warp-wide, the rare paths are group-local branches.
// False: this group's search is over (queue empty, early stop, or an overflow).
// `mode`: 0 = this lane's group is not running, 1 = search step, 2 = BFS step of the eager semantics (pops are
// the lattice states in FIFO order: st.wline is the cursor).
template <int G, int DENSE, bool SLAB, bool EAGER>
__device__ __forceinline__ bool lean_step(const SearchParams& p, const Group<G>& g, const LeanCtx& c, LeanState& st, const LhsBytes& lhs,
                                          uint32_t mode) {
  constexpr bool HOT = true;
  const DevFstView& F = p.fst;
  const bool running = mode != 0, bfs = EAGER && mode == 2;
  bool live = running && !(st.overflow || st.heap_overflow);
  bool over = running && !live;
  // ── pop the smallest ready id (:159-163; the set only holds unsettled tuples at the level distance) ──
  if (live && !bfs && st.low_pending) { lean_window_evict<G>(p, g, c, st); st.low_pending = false; }
  const uint32_t w = (live && !bfs) ? LEAN_WIN(c)[g.lane] : 0u;
  const unsigned bal = lean_ballot<G, HOT>(g, w != 0);
  if (live && !bfs && bal == 0) {   // rare: next window line, or next distance level; the pop happens in the next iteration
    if (!lean_window_next<G, DENSE>(p, g, c, st)) {
      if (!lean_advance_level<G, DENSE>(p, g, c, st)) over = true;
    }
    live = false;
  }
  if (live && bfs && st.wline >= st.n_tuples) { over = true; live = false; }   // FIFO drained: the lattice is complete
  const int src = (live && !bfs) ? __ffs(bal) - 1 : 0;
  const uint32_t ww = lean_shfl<G, HOT>(g, w, src);
  uint32_t cur_id = 0, s1 = 0, sf = 0;
  double bfs_dist = d_inf();
  if (live) {
    if (bfs) {
      cur_id = st.wline++;
    } else {
      const uint32_t bit = __ffs(ww) - 1;
      cur_id = ((st.wline * G + src) << 5) + bit;
      // clear the popped bit atomically: commutes with the atomic ORs of this iteration's inserts (other bits of
      // the same word), so no barrier is needed in between; the vote + barrier of lean_ready_insert orders all of
      // them before the next iteration reads the window
      if ((int)g.lane == src) atomicAnd(&LEAN_WIN(c)[src], ~(1u << bit));
    }
    lean_keyof_load<DENSE>(p, c, cur_id, s1, sf);
    if (bfs) { uint32_t ps, i2, pr; lean_lookup<DENSE>(p, c, s1, sf, ps, bfs_dist, i2, pr); }
  }
  const uint32_t s2 = sf >> 1;

  // eager final pick (shortest-path.zig:88-104): smallest (total, state number) over reached final states; the BFS
  // pops in increasing number, so only a strictly smaller total replaces the best
  if (live && bfs && s1 == lhs.len && !d_isinf(bfs_dist)) {
    const double fw2 = F.final_w[s2];
    if (!d_isinf(fw2)) {
      uint32_t* cold = LEAN_COLD(p, c, G);
      const double final_w = 0.0 + fw2;
      const double total = bfs_dist + final_w;
      const bool have_best = cold[kcHaveBest] != 0;
      const double best_total = __hiloint2double((int)cold[kcBestTotHi], (int)cold[kcBestTotLo]);
      g.sync();
      if (!have_best || total < best_total) {
        if (g.lane == 0) {
          cold[kcHaveBest] = 1; cold[kcBestId] = cur_id;
          cold[kcBestFwLo] = (uint32_t)__double2loint(final_w); cold[kcBestFwHi] = (uint32_t)__double2hiint(final_w);
          cold[kcBestTotLo] = (uint32_t)__double2loint(total); cold[kcBestTotHi] = (uint32_t)__double2hiint(total);
        }
      }
      g.sync();
    }
  }
  // final check (:165-179): only the last state of the string acceptor is final, weight One
  if (live && !bfs && s1 == lhs.len) {
    const double fw2 = F.final_w[s2];
    if (!d_isinf(fw2)) {
      uint32_t* cold = LEAN_COLD(p, c, G);
      const double final_w = 0.0 + fw2;
      const double total = __longlong_as_double((long long)st.last) + final_w;
      const bool have_best = cold[kcHaveBest] != 0;
      const double best_total = __hiloint2double((int)cold[kcBestTotHi], (int)cold[kcBestTotLo]);
      const uint32_t best_id = cold[kcBestId];
      g.sync();
      if (!have_best || total < best_total || (total == best_total && cur_id < best_id)) {
        if (g.lane == 0) {
          cold[kcHaveBest] = 1; cold[kcBestId] = cur_id;
          cold[kcBestFwLo] = (uint32_t)__double2loint(final_w); cold[kcBestFwHi] = (uint32_t)__double2hiint(final_w);
          cold[kcBestTotLo] = (uint32_t)__double2loint(total); cold[kcBestTotHi] = (uint32_t)__double2hiint(total);
        }
      }
      g.sync();
    }
  }
  // ── expansion (:182-202 match arcs, then :254-278 input-epsilon arcs; filter is 0 or 1 here) ──
  // one lane per arc of the state, in frozen order (the ilabel-0 prefix first); idle lanes hold ilabel 0xFFFFFFFF
  uint4 sa = make_uint4(0xFFFFFFFFu, 0x80000000u, 0, 0);
  uint4 rec = make_uint4(0, 0, 0, 0);
  uint32_t x = 0xFFFFFFFDu;
  bool big = false, csr = false;   // csr: the lanes hold CSR search records (full ilabel, one relax call each)
  // 8 lanes per string read the LEADER slab: one record per (ilabel, nextstate) group, labels first then the
  // input-epsilon records (lane order == expansion order), x = ilabel | folded arc count << 16
  constexpr bool LEADERS = SLAB && (G == 8 || G == 4);
  if (live) {
    x = s1 < lhs.len ? (uint32_t)__ldg(lhs.s + s1) + 1u : 0xFFFFFFFDu;
    if (LEADERS) {
      sa = __ldg(G == 8 ? &F.wslab[(uint64_t)s2 * kWaveSlots + g.lane] : &F.wslab4[(uint64_t)s2 * 4 + g.lane]);
      big = sa.x == kWaveBig;
      if (big && F.bigidx) {
        // a state wider than the slab: its label index gives the arcs with the string's label and the epsilon prefix;
        // if those fit the group (nearly always: a trie node has one child per label) they are fetched from the CSR
        // search records, match arcs first, and take the same single relax step as any other state
        const uint2* bi = F.bigidx + (uint64_t)sa.z * 257u;
        const uint2 e = __ldg(bi);
        const uint2 m = x <= 256u ? __ldg(bi + x) : make_uint2(0u, 0u);
        if (m.y + e.y <= (uint32_t)G) {
          sa = make_uint4(0xFFFFFFFFu, 0x80000000u, 0, 0);
          if (g.lane < m.y) sa = __ldg(&F.sarc[m.x + g.lane]);
          else if (g.lane < m.y + e.y) sa = __ldg(&F.sarc[e.x + (g.lane - m.y)]);
          big = false; csr = true;
        }
      }
    } else if (SLAB) {
      sa = __ldg(&F.slab[(uint64_t)s2 * G + g.lane]);
      big = sa.x == 0xFFFFFFFEu;
    } else {
      rec = __ldg(&F.state_rec[s2]);   // {arc_begin, eps_end, arc_end}
      const uint32_t deg = rec.z - rec.x;
      big = deg > (uint32_t)G;
      if (!big && g.lane < deg) sa = __ldg(&F.sarc[rec.x + g.lane]);
    }
  }
  const uint32_t lab = (LEADERS && !csr) ? (sa.x & 0xFFFFu) : sa.x;
  const bool is_match = lab == x;                  // x is never 0xFFFFFFFF / 0xFFFFFFFE / 0xFFFF (labels are byte + 1)
  const bool is_eps = lab == 0u;                   // ilabel 0 only occurs in the epsilon records
  const unsigned first = LEADERS ? Group<G>::kBits : lean_ballot<G, HOT>(g, is_match);
  const uint32_t my_cnt = (is_match || is_eps) ? ((LEADERS && !csr) ? (sa.x >> 16) : 1u) : 0u;
  lean_relax<G, DENSE, HOT, EAGER>(p, g, c, st, cur_id, my_cnt, (is_match || is_eps) && !(sa.y >> 31), is_match ? s1 + 1u : s1,
                            (sa.y << 1) | (is_match ? 0u : 1u), __hiloint2double((int)sa.w, (int)sa.z), first, bfs, bfs_dist);
  if (live && big) {
    // a state wider than the group: binary-searched match range, then the epsilon prefix, G arcs per step
    if (SLAB) rec = __ldg(&F.state_rec[s2]);
    uint32_t lo = 0, hi = 0;
    if (s1 < lhs.len) equal_range(g, F.ilabel, rec.x, rec.z, x, lo, hi);
    for (uint32_t cb = lo; cb < hi && !st.overflow && !st.heap_overflow; cb += G) {
      const bool cand = cb + g.lane < hi;
      uint4 sb = make_uint4(0, 0x80000000u, 0, 0);
      if (cand) sb = __ldg(&F.sarc[cb + g.lane]);
      lean_relax<G, DENSE, false, EAGER>(p, g, c, st, cur_id, cand ? 1u : 0u, cand && !(sb.y >> 31), s1 + 1u, sb.y << 1,
                                  __hiloint2double((int)sb.w, (int)sb.z), Group<G>::kBits, bfs, bfs_dist);
      g.sync();
    }
    for (uint32_t cb = rec.x; cb < rec.y && !st.overflow && !st.heap_overflow; cb += G) {
      const bool cand = cb + g.lane < rec.y;
      uint4 sb = make_uint4(0, 0x80000000u, 0, 0);
      if (cand) sb = __ldg(&F.sarc[cb + g.lane]);
      lean_relax<G, DENSE, false, EAGER>(p, g, c, st, cur_id, cand ? 1u : 0u, cand && !(sb.y >> 31), s1, (sb.y << 1) | 1u,
                                  __hiloint2double((int)sb.w, (int)sb.z), Group<G>::kBits, bfs, bfs_dist);
      g.sync();
    }
  }
  return !over;
}

// Start the BFS phase of the eager semantics after the search: the initial tuple is lattice state 0
// (compose.zig:57-61); st.wline becomes the FIFO cursor, st.n_tuples the number of states numbered so far.
template <int G, int DENSE>
__device__ __forceinline__ void lean_bfs_begin(const SearchParams& p, const Group<G>& g, const LeanCtx& c, LeanState& st) {
  uint32_t* cold = LEAN_COLD(p, c, G);
  g.sync();
  if (g.lane == 0) {
    const uint32_t SF = p.fst.start << 1;
    uint32_t pos, id, prev; double d;
    lean_lookup<DENSE>(p, c, 0u, SF, pos, d, id, prev);
    lean_store<DENSE>(c, pos, 0u, SF, 0.0, 0u | kBfsFlag, kNone);
    lean_keyof_store<DENSE>(p, c, 0u, 0u, SF);
    cold[kcHaveBest] = 0;
  }
  st.occ = st.n_tuples;
  st.n_tuples = 1; st.wline = 0; st.relax_calls = 0; st.bfs_started = true;
  g.sync();
}

// Eager lattice as CSR: the result of compose() (compose.zig:29-198) for this string, written after
// the BFS phase has numbered every lattice state.  State u (BFS number) = tuple key_of[u]; its arcs in the reference's
// order: match arcs (:95-109: ilabel of the string, olabel and weight One (x) w of the transducer arc, in frozen order),
// then the transducer's input-epsilon arcs (:138-149: ilabel 0); final weight One (x) fw2 on states at the end of the
// string whose transducer state is final (:69-74).  Lanes take consecutive states, a group prefix sum places the arcs.
template <int G, int DENSE>
__device__ __forceinline__ void lean_emit_lattice(const SearchParams& p, const Group<G>& g, const LeanCtx& c, const LeanState& st,
                                                  const LhsBytes& lhs, uint32_t idx) {
  const DevFstView& F = p.fst;
  const bool ok = st.bfs_started && !(st.overflow || st.heap_overflow || st.stuck) && !(DENSE >= 2 && g.any(st.wide));
  if (!ok) return;   // the string is retried (or failed): the pass that completes it emits
  const uint32_t n = st.n_tuples;
  unsigned long long total = st.relax_calls;   // per-lane partial count of the BFS phase == lattice arcs
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) total += __shfl_xor_sync(g.mask, total, o, G);
  unsigned long long sbase = 0, abase = 0;
  if (g.lane == 0) {
    sbase = atomicAdd(p.lat_cursors + 0, (unsigned long long)n);
    abase = atomicAdd(p.lat_cursors + 1, total);
    p.lat_state_base[idx] = sbase; p.lat_arc_base[idx] = abase; p.lat_n_states[idx] = n; p.lat_n_arcs[idx] = total;
  }
  sbase = g.shfl(sbase, 0); abase = g.shfl(abase, 0);
  if (sbase + n > p.lat_state_cap || abase + total > p.lat_arc_cap) return;   // sizes only: the host grows the arrays
  unsigned long long cursor = 0;
  for (uint32_t base = 0; base < n; base += G) {
    const uint32_t u = base + g.lane;
    uint32_t P = 0, SF = 0, lo = 0, hi = 0, e0 = 0, e1 = 0, x = 0;
    if (u < n) {
      lean_keyof_load<DENSE>(p, c, u, P, SF);
      const uint4 rec = __ldg(&F.state_rec[SF >> 1]);
      e0 = rec.x; e1 = rec.y;
      if (P < lhs.len) {
        x = (uint32_t)__ldg(lhs.s + P) + 1u;
        uint32_t l = rec.y, h = rec.z;
        while (l < h) { const uint32_t m = l + (h - l) / 2; if (__ldg(F.ilabel + m) < x) l = m + 1; else h = m; }
        lo = l; h = rec.z;
        uint32_t l2 = l;
        while (l2 < h) { const uint32_t m = l2 + (h - l2) / 2; if (__ldg(F.ilabel + m) <= x) l2 = m + 1; else h = m; }
        hi = l2;
      }
    }
    const uint32_t deg = (hi - lo) + (e1 - e0);
    uint32_t incl = deg;
#pragma unroll
    for (int o = 1; o < G; o <<= 1) { const uint32_t t = __shfl_up_sync(g.mask, incl, o, G); if ((int)g.lane >= o) incl += t; }
    const uint32_t step_total = g.shfl(incl, G - 1);
    if (u < n) {
      const unsigned long long first = cursor + (incl - deg);
      p.lat_arc_begin[sbase + u] = (uint32_t)first;
      double fin = d_inf();
      if (P == lhs.len) { const double fw2 = F.final_w[SF >> 1]; if (!d_isinf(fw2)) fin = 0.0 + fw2; }
      p.lat_final[sbase + u] = fin;
      unsigned long long w = abase + first;
      for (uint32_t a = lo; a < hi; a++, w++) {
        const uint4 pl = __ldg(&F.payload[a]);
        uint32_t pos, id, prev; double d;
        lean_lookup<DENSE>(p, c, P + 1u, pl.y << 1, pos, d, id, prev);
        p.lat_il[w] = x; p.lat_ol[w] = pl.x; p.lat_w[w] = 0.0 + __hiloint2double((int)pl.w, (int)pl.z); p.lat_next[w] = id & ~kBfsFlag;
      }
      for (uint32_t a = e0; a < e1; a++, w++) {
        const uint4 pl = __ldg(&F.payload[a]);
        uint32_t pos, id, prev; double d;
        lean_lookup<DENSE>(p, c, P, (pl.y << 1) | 1u, pos, d, id, prev);
        p.lat_il[w] = 0u; p.lat_ol[w] = pl.x; p.lat_w[w] = __hiloint2double((int)pl.w, (int)pl.z); p.lat_next[w] = id & ~kBfsFlag;
      }
    }
    cursor += step_total;
  }
}

// End of a string: back-track, emit the reversed path into the pool, restore the arena invariants.
template <int G, int DENSE>
__device__ __forceinline__ int32_t lean_finish(const SearchParams& p, const Group<G>& g, const LeanCtx& c, const LeanState& st,
                                               const LhsBytes& lhs, uint32_t* out_path_len, uint64_t* out_pool_off, double* out_final_w) {
  const DevFstView& F = p.fst;
  uint32_t* cold = LEAN_COLD(p, c, G);
  int32_t status = kStPath;
  uint32_t plen = 0;
  unsigned long long poff = 0;
  uint32_t* scratch = LEAN_CHUNKS(p, c);   // the future set is dead now
  const uint32_t scratch_cap = p.heap_cap * 32u;
  g.sync();
  const bool have_best = cold[kcHaveBest] != 0;
  const uint32_t best_id = cold[kcBestId];
  const double best_fw = __hiloint2double((int)cold[kcBestFwHi], (int)cold[kcBestFwLo]);
  const bool wide = DENSE >= 2 && g.any(st.wide);
  if (wide) {
    status = kStRetryWide;
  } else if (st.stuck) {
    status = kStInternal;
  } else if (st.overflow) {
    status = kStRetry;
  } else if (st.heap_overflow) {
    status = kStRetryHeap;
  } else if (!have_best) {
    status = kStNoPath;                                               // :368-370
  } else {
    if (g.lane == 0) {                                                // :372-380 back-track (ids only)
      uint32_t cur = best_id;
      const bool eager = st.bfs_started;
      while (eager || cur != 0) {
        uint32_t P, SF, pos, id, prev; double d;
        lean_keyof_load<DENSE>(p, c, cur, P, SF);
        lean_lookup<DENSE>(p, c, P, SF, pos, d, id, prev);
        if (prev == kNone) {
          // lazy: a missing back-pointer before the initial tuple -> empty FST (:375-377); eager: the chain ends at
          // the first state without back-pointer, which must be the start state (shortest-path.zig:113-123)
          if (!eager || cur != 0) status = kStNoPath;
          break;
        }
        if (plen >= st.n_tuples) { status = kStCycle; break; }        // hazard H1 (reference: out of memory)
        if (plen >= scratch_cap) { status = kStRetryHeap; break; }
        scratch[plen++] = cur;
        cur = prev;
      }
      if (status == kStPath && plen > 0) {
        poff = atomicAdd(p.pool_cursor, (unsigned long long)plen);
        if (poff + plen > p.pool_cap) status = kStRetry;
      }
    }
    g.sync();
    status = g.shfl(status, 0); plen = g.shfl(plen, 0); poff = g.shfl(poff, 0);
    if (status == kStPath) {
      bool bad = false;
      for (uint32_t i = g.lane; i < plen; i += G) {
        uint32_t pv, sfv, pu, sfu, pos, id, prev, id2, prev2; double dv, du;
        lean_keyof_load<DENSE>(p, c, scratch[i], pv, sfv);
        lean_lookup<DENSE>(p, c, pv, sfv, pos, dv, id, prev);
        lean_keyof_load<DENSE>(p, c, prev, pu, sfu);
        lean_lookup<DENSE>(p, c, pu, sfu, pos, du, id2, prev2);
        PoolArc pa; pa.ilabel = 0; pa.olabel = 0; pa.weight = 0.0;
        if (!lean_recover_arc(F, lhs, pu, sfu >> 1, pv, sfv >> 1, du, dv, pa)) bad = true;
        p.pool[poff + i] = pa;
      }
      if (g.any(bad)) { status = kStInternal; plen = 0; }
    } else {
      plen = 0;
    }
  }
  // Restore the arena invariants for the next string: table untouched-state, bitmaps zero.
  g.sync();
  const uint32_t n = st.n_tuples;
  const bool aborted = st.overflow || st.heap_overflow || st.stuck || wide;
  const bool mixed = st.bfs_started && aborted;   // id -> key array is part search ids, part BFS numbers
  if (mixed) {
    uint4* t = reinterpret_cast<uint4*>(c.base);
    const uint64_t vecs = DENSE >= 2 ? (p.tab_entries + 1) / 2 : p.tab_entries * (DENSE ? 1ull : 2ull);
    for (uint64_t i = g.lane; i < vecs; i += G) t[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
  } else if (DENSE) {
    if ((uint64_t)n * 4 < p.tab_entries) {
      for (uint32_t i = g.lane; i < n; i += G) {
        uint32_t P, SF;
        lean_keyof_load<1>(p, c, i, P, SF);
        if (DENSE >= 2) *reinterpret_cast<uint2*>(c.base + (uint64_t)lean_dense_pos(p, P, SF) * 8) = make_uint2(~0u, ~0u);
        else *reinterpret_cast<uint4*>(c.base + (uint64_t)lean_dense_pos(p, P, SF) * 16) = make_uint4(~0u, ~0u, ~0u, ~0u);
      }
    } else {
      uint4* t = reinterpret_cast<uint4*>(c.base);
      const uint64_t vecs = DENSE >= 2 ? (p.tab_entries + 1) / 2 : p.tab_entries;   // the table is padded to 128 bytes
      for (uint64_t i = g.lane; i < vecs; i += G) t[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
    }
  } else {
    // two phases: resolve every tuple's slot first (probing needs intact chains), then clear
    unsigned long long* key_of = reinterpret_cast<unsigned long long*>(LEAN_KEYOF(p, c));
    uint32_t* slot_tmp = reinterpret_cast<uint32_t*>(key_of);
    for (uint32_t base = 0; base < n; base += G) {
      const uint32_t i = base + g.lane;
      uint32_t sl = 0, id, prev; double d;
      if (i < n) { const unsigned long long K = key_of[i]; lean_lookup<false>(p, c, (uint32_t)(K >> 32), (uint32_t)K, sl, d, id, prev); }
      g.sync();                        // all keys of this stripe are read before any is overwritten
      if (i < n) slot_tmp[i] = sl;     // aliases key_of[i/2]: only stripes already resolved
      g.sync();
    }
    LeanSlot* tab = reinterpret_cast<LeanSlot*>(c.base);
    for (uint32_t i = g.lane; i < n; i += G) tab[slot_tmp[i]].key = kEmptyKey;
  }
  if (aborted) {
    // aborted searches can leave ready bits behind
    for (uint32_t i = g.lane; i < ((p.tuple_cap + 32u * G - 1) / (32u * G)) * G; i += G) LEAN_L0(p, c)[i] = 0;
    for (uint32_t i = g.lane; i < p.n1; i += G) LEAN_L1(c, G)[i] = 0;
    LEAN_WIN(c)[g.lane] = 0;
  }
  g.sync();
  *out_path_len = plen; *out_pool_off = poff; *out_final_w = (status == kStPath) ? best_fw : d_inf();
  return status;
}

// Persistent batch kernel.  One loop; an iteration is one step of this group's string (fetch the next
// string / one pop / finish), with a warp-wide reconvergence at the top so that the 16-lane groups of a
// warp stay in lockstep.
template <int G, int DENSE, bool SLAB, bool EAGER>
__global__ void __launch_bounds__(128, FSTB_LEAN_MINBLOCKS) csp_batch_lean_kernel(const __grid_constant__ SearchParams p) {
  extern __shared__ __align__(16) uint32_t smem_all[];
  const Group<G> g;
  const uint32_t gib = threadIdx.x / G;
  const uint32_t gslot = blockIdx.x * (blockDim.x / G) + gib;
  LeanCtx c;
  c.base = p.arena + (uint64_t)gslot * p.arena_stride;
  c.sm = smem_all + (size_t)gib * p.smem_words;
  for (uint32_t i = 128 + g.lane; i < p.smem_words; i += G) c.sm[i] = 0;   // window, summary and cold state start empty
  g.sync();
  enum { kFetch = 0, kRun = 1, kFinish = 2, kDone = 3, kBfs = 4 };
  uint32_t phase = kFetch, idx = 0;
  uint32_t steps = 0;   // safety valve: steps of the current string (tuple_cap <= 4 M, so the bound fits 32 bits)
  LeanState st;
  st.n_tuples = 0; st.wline = kNone; st.relax_calls = 0; st.last = 0; st.future_min = ~0ull;
  st.low_pending = false; st.overflow = false; st.heap_overflow = false; st.sorted = false; st.bfs_started = false; st.stuck = false; st.occ = 0; st.wide = false;
  LhsBytes lhs; lhs.s = nullptr; lhs.len = 0;
  unsigned long long relax_total = 0, tuple_total = 0;
  for (;;) {
    if (phase == kFetch) {
      uint32_t item = 0;
      if (g.lane == 0) item = atomicAdd(p.queue_head, 1u);
      item = g.shfl(item, 0);
      if (item >= p.n_items) {
        phase = kDone;
      } else {
        idx = p.order ? p.order[item] : item;
        lhs.s = p.bytes + p.offsets[idx]; lhs.len = (uint32_t)(p.offsets[idx + 1] - p.offsets[idx]);
        const int32_t pre = p.skip ? p.skip[idx] : kStPath;
        if (p.fst.start == kNone || pre != kStPath) {
          if (g.lane == 0) { p.status[idx] = pre != kStPath ? pre : kStNoPath; p.path_len[idx] = 0; p.pool_off[idx] = 0; p.final_w[idx] = d_inf(); p.n_tuples[idx] = 0; }
        } else {
          lean_begin<G, DENSE>(p, g, c, st);
          phase = kRun; steps = 0;
        }
      }
    }
    // the step is warp-uniform: every lane enters it whenever a group of the warp is running
    const bool stepping = phase == kRun || (EAGER && phase == kBfs);
    const bool anyrun = G < 32 ? __any_sync(0xFFFFFFFFu, stepping) : stepping;
    if (!anyrun) {   // nobody runs: either everybody is done, or somebody finishes/fetches below
      if (G < 32 ? __all_sync(0xFFFFFFFFu, phase == kDone) : phase == kDone) break;
    } else {
      bool cont = lean_step<G, DENSE, SLAB, EAGER>(p, g, c, st, lhs, phase == kRun ? 1u : (phase == kBfs ? 2u : 0u));
      // every step pops a tuple or retires a window line / distance level: more than a few steps per tuple slot
      // means the engine is not making progress — fail the string (kStInternal) instead of spinning
      if (stepping && ++steps > 8u * p.tuple_cap + 4096u) { st.stuck = true; cont = false; }
      if (stepping && !cont) {
        if (EAGER && phase == kRun && !st.overflow && !st.heap_overflow && !st.stuck) { lean_bfs_begin<G, DENSE>(p, g, c, st); phase = kBfs; }
        else phase = kFinish;
      }
    }
    if (phase == kFinish) {
      uint32_t plen; uint64_t poff; double fw;
      if (EAGER && p.lat_state_base != nullptr) lean_emit_lattice<G, DENSE>(p, g, c, st, lhs, idx);
      const int32_t status = lean_finish<G, DENSE>(p, g, c, st, lhs, &plen, &poff, &fw);
      if (g.lane == 0) {
        p.status[idx] = status; p.path_len[idx] = plen; p.pool_off[idx] = poff; p.final_w[idx] = fw; p.n_tuples[idx] = st.n_tuples;
      }
      relax_total += st.relax_calls; tuple_total += st.n_tuples;
      phase = kFetch;
    }
  }
  // relax counts are per-lane partial sums: add up the warp
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) relax_total += __shfl_xor_sync(0xFFFFFFFFu, relax_total, o);
  if ((threadIdx.x & 31u) == 0 && relax_total) atomicAdd(p.relax_counter, relax_total);
  if (g.lane == 0 && tuple_total) atomicAdd(p.tuple_counter, tuple_total);
}

}  // namespace fstb200
