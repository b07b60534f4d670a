// Device-side data layout shared by the upload code, the kernels and the engine.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fstb200 {

constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr uint32_t kWaveSlots = 8;            // search records per transducer state in the leader slab (`wslab`)
constexpr uint32_t kWaveBig = 0xFFFFFFFEu;    // leader slab marker (every slot): the state has more records than the slab holds

// Frozen transducer in HBM (struct-of-arrays CSR, frozen arc order unchanged —
// reference src/fst.zig:16-40 and SURVEY App. D):
//   state_rec[s] = {arc_begin, eps_end, arc_end, 0}   absolute arc indices; the
//                  arcs [arc_begin, eps_end) are the state's ilabel==0 prefix
//                  (valid because frozen arcs are ilabel-sorted, src/arc.zig:46-54)
//   final_w[s]   = f64 final weight (+inf = not final)
//   ilabel[a]    = u32 search key array (binary-searched, src/fst.zig:112-136)
//   payload[a]   = {olabel, nextstate, weight(f64 as 2xu32)} one 16-byte vector load
//   sarc[a]      = SEARCH record of the lean batched kernel, one 16-byte load per lane per pop:
//                  {ilabel, nextstate | dup << 31, wmin(f64 as 2xu32)}.  Arcs of one state with the
//                  same (ilabel, nextstate) always hit the same compose tuple; the first of them in
//                  frozen order (dup == 0) carries wmin = the smallest weight of the group and relaxes
//                  for all of them, the others (dup == 1) only count as relax calls.  Static per
//                  transducer, computed once at upload (see csp_lean.cuh, "fold").
struct DevFstView {
  uint32_t num_states, num_arcs, start, max_degree;
  const uint4* state_rec;
  const double* final_w;
  const uint32_t* ilabel;
  const uint4* payload;
  const uint4* sarc;
  // optional fixed-stride copy of sarc: state s owns slab[s * slab_lanes .. +slab_lanes), unused lanes hold
  // ilabel 0xFFFFFFFF; a state wider than slab_lanes holds the marker ilabel 0xFFFFFFFE (use the CSR arrays).
  // Saves the state_rec hop of the per-pop dependent load chain.  slab_lanes == 0: not built.
  const uint4* slab;
  uint32_t slab_lanes, pad0;
  // optional leader slab (lean kernel with 8 lanes per string, wave kernel): state s owns wslab[s * 8 .. +8): its LEADER search records only (one per
  // (ilabel, nextstate) group), labels 1..256 in frozen order first, then the input-epsilon records, as
  // {ilabel | folded arc count << 16, nextstate, wmin(f64 as 2xu32)}; unused slots hold 0xFFFFFFFF in x; a state
  // with more than 8 records (or a group of more than 65535 arcs) holds the marker 0xFFFFFFFE in every slot.
  // Arcs with ilabel > 256 never match a byte string and are left out.  null: not built.
  const uint4* wslab;
  // the same with 4 records per state (lean kernel with 4 lanes per string: sparse transducers — most states of a
  // real grammar have one or two arcs); a state with more than 4 records holds the marker in every slot.  null: not built.
  const uint4* wslab4;
  // label index of the states the leader slab cannot hold (their marker record carries the index number in z):
  // bigidx[(n * 257 + label)] = {first arc, arc count} of the state's arcs with that ilabel (label 0 = the
  // input-epsilon prefix), so that an expansion fetches only the few arcs that can match.  null: not built.
  const uint2* bigidx;
  // integer copy of the leader slab for transducers whose weights are all integers in 0..4095 (fast kernel,
  // csp_fast.cuh): same records with z = weight << 12 (the distance field of the compact table record) and w = 0,
  // plus one all-idle row at index num_states that groups without a string read.  null: not built.
  const uint4* islab;
};

// General (non-linear) left operand as CSR in STORED arc order
// (reference src/mutable-fst.zig:183 `arcs`, no sorting).
struct DevLhsCsr {
  uint32_t num_states, num_arcs, start, pad;
  const uint32_t* arc_off;  // [S+1]
  const double* final_w;    // [S]
  const uint32_t* ilabel;   // [A]
  const uint32_t* olabel;   // [A]
  const double* weight;     // [A]
  const uint32_t* next;     // [A]
};

// One search-state record: a slot of the per-string open-addressing table.
// 32 bytes = one DRAM sector: a relaxation touches exactly one sector.
struct __align__(32) TupleSlot {
  unsigned long long key;  // (s1 << 34) | (s2 << 2) | filter ; all-ones = empty
  double dist;             // tentative / final distance (tropical)
  uint32_t id_flags;       // discovery-order id (bits 0..30) | settled (bit 31)
  uint32_t prev_id;        // back-pointer: predecessor id, kNone = no back-pointer
  uint32_t rhs_arc;        // back-pointer: transducer arc index or kNone
  uint32_t lhs_arc;        // back-pointer: left-operand arc index or kNone
};
static_assert(sizeof(TupleSlot) == 32, "TupleSlot must be one 32-byte sector");

constexpr unsigned long long kEmptyKey = ~0ull;
constexpr uint32_t kSettledBit = 0x80000000u;

// Per-string status written by the search kernel (internal; the C ABI maps
// kRetry to a retry pass and finally to FST_B200_TOO_LARGE).
// kStNotBytes: a path was found (path arrays are valid) but its output tape holds a label above 256, so it has no
// byte-string form (the reference's fst_print_output_string returns -1 for such a chain).
enum : int32_t { kStPath = 0, kStNoPath = 1, kStCycle = 2, kStTooLarge = 3, kStInternal = 4, kStNotBytes = 5, kStRetry = 100, kStRetryHeap = 101, kStRetryWide = 102 };

// Reversed path arc as written to the path pool by the search kernel.
struct __align__(16) PoolArc { uint32_t ilabel, olabel; double weight; };

struct SearchParams {
  DevFstView fst;
  // batch of byte strings (linear left operands, label = byte + 1)
  const uint8_t* bytes;
  const uint64_t* offsets;
  const uint32_t* order;   // optional: work item i -> string index (retry passes); null = identity
  const int32_t* skip;     // optional: per-string status of an earlier pipeline stage; anything but kStPath is passed
                           // through as this string's status and the string is not searched
  uint32_t n_items;
  // or: one general left operand (n_items == 1, bytes == nullptr)
  DevLhsCsr lhs;
  // per-group arenas
  uint8_t* arena;
  uint64_t arena_stride;
  uint32_t hash_cap;       // power of two
  uint32_t tuple_cap;
  uint32_t heap_cap;       // binary-heap entries (serial path) / 128-byte chunks (warp path)
  uint32_t bag_cap;        // warp path: unsorted future ids
  uint32_t exhaustive;
  uint32_t dense;          // lean path: 1 = direct-indexed table, 0 = hash table
  uint64_t tab_entries;    // lean path: dense records ((max_len+1) * S * 2) or hash slots
  // lean path: arena geometry precomputed on the host (read from the constant bank at the point of use)
  uint64_t off_keyof, off_l0, off_chunks;
  uint32_t n1;             // ready-bitmap summary words
  uint32_t smem_words;     // shared-memory words per group
  uint32_t dense_stride;   // 2 * S (bits of the compact id -> key form)
  // dense table index of tuple (P, SF = state << 1 | filter), an affine map (mod 2^32):
  //     pos = SF * pos_h + (filter ? pos_k : 0) + pos_c + P * pos_m2
  //   position-major (pos_h 1, pos_k 0, pos_m2 2S, pos_c 0):   pos = P * 2S + SF   — a row per string position
  //   diagonal(skew) (pos_h H = half row of max_len + 2 records padded to a whole number of 32-byte sectors, pos_k 1,
  //   pos_m2 1 - skew 2H, pos_c skew max_len 2H):
  //                                   pos = ((state - skew P + skew max_len) * 2 + filter) * H + P + filter
  //   — a row per DIAGONAL state - skew * position, the two filter variants in separate half rows.  On a chain-like
  //   transducer with input-epsilon arcs the reference's pop order runs along such diagonals (engine.cuh,
  //   DeviceFst::layout_skew): a chain of pops (position + k, state + k skew) and its match targets walk along a few
  //   half rows 8 bytes per step; the filter-1 half row (the epsilon target sits at the popped position, the match
  //   targets at the next one) is shifted by one record and the half rows are sector-padded, so that all streams
  //   enter a new 32-byte sector in the same step: one DRAM round trip every fourth pop of a chain instead of every pop.
  uint32_t pos_h, pos_m2, pos_c, pos_k;
  uint32_t key_sbits;      // dense: bits of ((s << 1) | filter) in the compact id -> key array
  uint32_t eager;          // lean path: 1 = eager semantics (compose() then shortestPath(), BASELINE config 5)
  // work queue + counters
  uint32_t* queue_head;
  unsigned long long* pool_cursor;
  unsigned long long* relax_counter;
  unsigned long long* tuple_counter;
  unsigned long long* wave_stats;   // optional [4]: chunk steps, tuples popped by chunks, single-pop steps, abandoned chunks
  // optional eager lattice output (csp_lean.cuh lean_emit_lattice; lat_state_base == null: off).
  // Every string reserves n_states / n_arcs slots of the flat arrays with the two cursors; a string that does not
  // fit writes only its sizes (the host grows the arrays and runs again).
  unsigned long long* lat_cursors;   // [2]: states, arcs
  uint64_t lat_state_cap, lat_arc_cap;
  uint64_t* lat_state_base;          // [n] first slot of the string in the per-state arrays
  uint64_t* lat_arc_base;            // [n] first slot of the string in the per-arc arrays
  uint32_t* lat_n_states;            // [n]
  uint64_t* lat_n_arcs;              // [n]
  uint32_t* lat_arc_begin;           // per state: first arc of the state, relative to the string's arc base
  double* lat_final;                 // per state: final weight (+inf = not final)
  uint32_t* lat_il; uint32_t* lat_ol; uint32_t* lat_next; double* lat_w;   // per arc
  // outputs (indexed by string index)
  int32_t* status;
  uint32_t* path_len;
  uint64_t* pool_off;
  double* final_w;
  uint32_t* n_tuples;
  PoolArc* pool;
  uint64_t pool_cap;
};

struct EmitParams {
  const int32_t* status;
  const uint32_t* path_len;
  const uint64_t* pool_off;
  const uint64_t* path_offsets;  // exclusive scan of path_len, [n+1]
  const PoolArc* pool;
  uint32_t n_strings;
  uint32_t* ilabels;
  uint32_t* olabels;
  double* weights;
  uint64_t path_capacity;
  // optional fused output-tape string (label-1, epsilons dropped)
  const uint64_t* out_offsets;   // exclusive scan of out_len
  uint8_t* out_bytes;
};

}  // namespace fstb200
