// Warp-per-string exact search for non-negative weights (the production path).
//
// With non-negative weights the reference's pops are monotone in distance, so its
// (dist, id) min-heap (compose-shortest-path.zig:55-61) can be replaced, without
// changing the pop sequence, by
//   * a READY SET: the unsettled tuples whose tentative distance equals the
//     current level distance, kept as a hierarchical bitmap indexed by discovery
//     id (level 0 in HBM/L1, upper levels in shared memory); "pop" = find-first-set
//     = the smallest id, exactly the heap's tie rule;
//   * a FUTURE SET of ids whose tentative distance is larger.  It starts as an
//     unsorted bag (one coalesced append per expansion).  Only if a second
//     distance level is really needed is it turned into a radix heap over the
//     IEEE-754 bit pattern of the distance (64 buckets by the highest bit in which
//     a key differs from the last popped key).  Entries are ids only; an entry is
//     valid iff the tuple's CURRENT distance still maps to the bucket it sits in (a
//     lowered tuple always has a second entry in the right place, so stale ones
//     are dropped — the reference skips them at pop time, :162).  A per-lane
//     running minimum of pushed distances is a lower bound of the next level; when
//     it already exceeds the best total the search stops without touching the bag.
// The reference's `settled` flag is not needed: a tuple at the current level is
// either in the ready set or already expanded, and in both cases a tie relaxation
// (:115-126) changes the back-pointer only; strictly better relaxations can only
// hit unsettled tuples.  See DESIGN.md §"exactness".
//
// All collectives use the full-warp mask known at compile time (single SASS
// instructions); there are no sub-warp groups on this path.
#pragma once
#include <type_traits>

#include "csp_kernels.cuh"

namespace fstb200 {

constexpr uint32_t kChunkIds = 31;             // ids per 128-byte chunk (word 0 = next chunk)
constexpr uint32_t kNoChunk = 0xFFFFFFFFu;
constexpr uint32_t kMaxFastTuples = 4u << 20;  // bitmap capacity 64*32*32*64
constexpr unsigned kFull = 0xFFFFFFFFu;

struct WarpLayout {
  uint64_t off_table, off_keyof, off_l0, off_bag, off_chunks, total;
  uint32_t n0, n1, n2, n3;       // words per bitmap level (l0: u64 in HBM; l1..l3: u32 in smem)
  uint32_t smem_words;           // u32 words of shared memory per warp (buckets + l1..l3)
};
__host__ __device__ inline WarpLayout warp_layout(uint32_t hash_cap, uint32_t tuple_cap, uint32_t chunk_cap, uint32_t bag_cap) {
  WarpLayout L;
  auto al = [](uint64_t x) { return (x + 127) & ~127ull; };
  L.n0 = (tuple_cap + 63) / 64; L.n1 = (L.n0 + 31) / 32; L.n2 = (L.n1 + 31) / 32; L.n3 = (L.n2 + 31) / 32;
  L.off_table = 0;
  L.off_keyof = al((uint64_t)hash_cap * sizeof(TupleSlot));
  L.off_l0 = L.off_keyof + al((uint64_t)tuple_cap * 8);
  L.off_bag = L.off_l0 + al((uint64_t)L.n0 * 8);
  L.off_chunks = L.off_bag + al((uint64_t)bag_cap * 4);
  L.total = (L.off_chunks + (uint64_t)chunk_cap * 128 + 255) & ~255ull;
  L.smem_words = (128 + L.n1 + L.n2 + L.n3 + 2 + 3) & ~3u;   // +2: ready_pop reads l3[0..1]; 16-byte multiple
  return L;
}

struct WarpArena {
  TupleSlot* table;
  unsigned long long* key_of;   // id -> tuple key
  unsigned long long* l0;       // ready bitmap level 0 (bit per id)
  uint32_t* bag;                // unsorted future ids
  uint32_t* chunks;             // radix-heap chunk pool; chunk c = chunks[c*32 .. c*32+31], word 0 = next
  uint2* bucket;                // smem [64] {head chunk, ids in head chunk}
  uint32_t* l1; uint32_t* l2; uint32_t* l3;   // smem bitmap levels
  uint32_t hash_cap, tuple_cap, chunk_cap, bag_cap, n0, n1, n2, n3;
};

// Arena initialisation: table keys empty, level-0 bitmap zero (run when the layout changes).
__global__ void warp_arena_init_kernel(uint8_t* arena, uint64_t stride, uint32_t n_arenas, uint32_t hash_cap, uint32_t tuple_cap,
                                       uint32_t chunk_cap, uint32_t bag_cap) {
  WarpLayout L = warp_layout(hash_cap, tuple_cap, chunk_cap, bag_cap);
  const uint64_t table_words = (uint64_t)hash_cap * 4;
  const uint64_t bitmap_words = L.n0;
  const uint64_t per = table_words + bitmap_words;
  const uint64_t total = per * n_arenas;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t ar = i / per, w = i % per;
    unsigned long long* base = reinterpret_cast<unsigned long long*>(arena + ar * stride);
    if (w < table_words) base[w] = ((w & 3) == 0) ? kEmptyKey : 0ull;
    else base[L.off_l0 / 8 + (w - table_words)] = 0ull;
  }
}

__device__ __forceinline__ uint32_t w_home(const WarpArena& a, unsigned long long key) {
  return (uint32_t)(((unsigned long long)hash_key(key) * a.hash_cap) >> 32);
}
// Find `key`: returns true and its slot contents, or false and the empty position.
__device__ __forceinline__ bool w_probe(const WarpArena& a, unsigned long long key, uint32_t& pos, TupleSlot& s) {
  uint32_t i = w_home(a, key);
  for (;;) {
    s = a.table[i];
    if (s.key == key) { pos = i; return true; }
    if (s.key == kEmptyKey) { pos = i; return false; }
    if (++i == a.hash_cap) i = 0;
  }
}
__device__ __forceinline__ uint32_t w_probe_pos(const WarpArena& a, unsigned long long key) {
  uint32_t i = w_home(a, key);
  for (;;) {
    unsigned long long k = a.table[i].key;
    if (k == key || k == kEmptyKey) return i;
    if (++i == a.hash_cap) i = 0;
  }
}
__device__ __forceinline__ uint32_t w_claim(const WarpArena& a, unsigned long long key, uint32_t pos) {
  for (;;) {
    unsigned long long old = atomicCAS(&a.table[pos].key, kEmptyKey, key);
    if (old == kEmptyKey) return pos;
    if (++pos == a.hash_cap) pos = 0;
  }
}
__device__ __forceinline__ uint32_t bucket_of(unsigned long long k, unsigned long long last) {
  return 64u - (uint32_t)__clzll((long long)(k ^ last));   // 1..64 for k != last
}
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
  for (int o = 16; o > 0; o >>= 1) { unsigned long long t = __shfl_xor_sync(kFull, v, o); v = t < v ? t : v; }
  return v;
}

// Uniform (warp-replicated) search state.
struct WarpState {
  uint32_t n_tuples;
  uint32_t chunk_next, free_head, bag_len;
  unsigned long long occupied;   // radix buckets in use
  unsigned long long last;       // key (bit pattern) of the current level distance
  unsigned long long relax_calls;
  bool overflow, sorted, lossy;
};

__device__ __forceinline__ bool ready_empty(const WarpArena& a) {
  uint32_t t = 0;
  for (uint32_t i = 0; i < a.n3; i++) t |= a.l3[i];
  return t == 0;
}

// Insert ids into the ready bitmap (warp-collective).
__device__ __forceinline__ void ready_insert(const WarpArena& a, unsigned lane, bool active, uint32_t id) {
  unsigned m = __ballot_sync(kFull, active);
  const uint32_t wi = id >> 6;
  while (m) {
    const int src = __ffs(m) - 1;
    const uint32_t w = __shfl_sync(kFull, wi, src);
    const bool mine = active && wi == w;
    m &= ~__ballot_sync(kFull, mine);
    const unsigned long long bit = mine ? (1ull << (id & 63u)) : 0ull;
    const uint32_t lo = __reduce_or_sync(kFull, (uint32_t)bit);
    const uint32_t hi = __reduce_or_sync(kFull, (uint32_t)(bit >> 32));
    if ((int)lane == src) {
      a.l0[w] |= ((unsigned long long)hi << 32) | lo;
      atomicOr(&a.l1[w >> 5], 1u << (w & 31u));
      atomicOr(&a.l2[w >> 10], 1u << ((w >> 5) & 31u));
      atomicOr(&a.l3[w >> 15], 1u << ((w >> 10) & 31u));
    }
  }
}

// Pop the smallest ready id (warp-collective; the set must be non-empty).
__device__ __forceinline__ uint32_t ready_pop(const WarpArena& a, unsigned lane) {
  uint32_t j3 = 0, t3 = a.l3[0];
  if (t3 == 0) { j3 = 1; t3 = a.l3[1]; }
  const uint32_t i2 = j3 * 32 + (__ffs(t3) - 1);
  const uint32_t w2 = a.l2[i2];
  const uint32_t i1 = i2 * 32 + (__ffs(w2) - 1);
  const uint32_t w1 = a.l1[i1];
  const uint32_t i0 = i1 * 32 + (__ffs(w1) - 1);
  unsigned long long w0 = a.l0[i0];
  const uint32_t id = i0 * 64 + (__ffsll((long long)w0) - 1);
  w0 &= w0 - 1;
  if (lane == 0) {
    a.l0[i0] = w0;
    if (w0 == 0) {
      const uint32_t n1 = w1 & ~(1u << (i0 & 31u));
      a.l1[i1] = n1;
      if (n1 == 0) {
        const uint32_t n2 = w2 & ~(1u << (i1 & 31u));
        a.l2[i2] = n2;
        if (n2 == 0) a.l3[j3] = t3 & ~(1u << (i2 & 31u));
      }
    }
  }
  __syncwarp();
  return id;
}

__device__ __forceinline__ uint32_t chunk_alloc(const WarpArena& a, WarpState& st) {
  uint32_t c;
  if (st.free_head != kNoChunk) {
    c = st.free_head;
    st.free_head = a.chunks[(uint64_t)c * 32];
  } else if (st.chunk_next < a.chunk_cap) {
    c = st.chunk_next++;
  } else {
    st.overflow = true; c = 0;
  }
  return c;
}

// Append ids to radix buckets (warp-collective).  `b` in 1..64 for active lanes.
__device__ inline void bucket_push(const WarpArena& a, WarpState& st, unsigned lane, bool active, uint32_t id, uint32_t b) {
  unsigned m = __ballot_sync(kFull, active);
  const unsigned lt = (1u << lane) - 1u;
  while (m) {
    const int first = __ffs(m) - 1;
    const uint32_t bb = __shfl_sync(kFull, b, first);
    const bool mine = active && b == bb;
    const unsigned same = __ballot_sync(kFull, mine);
    m &= ~same;
    const uint32_t k = __popc(same), rank = __popc(same & lt);
    const bool empty = !((st.occupied >> (bb - 1)) & 1ull);
    const uint2 hb = a.bucket[bb - 1];
    uint32_t head = empty ? kNoChunk : hb.x, cnt = empty ? kChunkIds : hb.y;
    const uint32_t space = kChunkIds - cnt;
    if (mine && rank < space) a.chunks[(uint64_t)head * 32 + 1 + cnt + rank] = id;
    uint32_t left = k > space ? k - space : 0, done = k - left;
    if (left == 0) cnt += k;
    while (left > 0) {
      const uint32_t c = chunk_alloc(a, st);
      if (st.overflow) return;
      const uint32_t take = left < kChunkIds ? left : kChunkIds;
      if (lane == 0) a.chunks[(uint64_t)c * 32] = head;
      if (mine && rank >= done && rank < done + take) a.chunks[(uint64_t)c * 32 + 1 + (rank - done)] = id;
      head = c; cnt = take; done += take; left -= take;
    }
    __syncwarp();
    if (lane == 0) a.bucket[bb - 1] = make_uint2(head, cnt);
    st.occupied |= 1ull << (bb - 1);
    __syncwarp();
  }
}

// Queue action for relaxed targets: ready set if at the current level, else future set.
__device__ __forceinline__ void queue_insert(const WarpArena& a, WarpState& st, unsigned lane, bool need, uint32_t id, double dist,
                                             unsigned long long& future_min) {
  const unsigned long long k = (unsigned long long)__double_as_longlong(dist);
  const bool to_ready = need && k == st.last;
  const bool to_future = need && k != st.last;
  ready_insert(a, lane, to_ready, id);
  const unsigned f = __ballot_sync(kFull, to_future);
  if (f) {
    if (to_future && k < future_min) future_min = k;
    if (st.sorted) {
      bucket_push(a, st, lane, to_future, id, to_future ? bucket_of(k, st.last) : 1u);
    } else if (!st.lossy) {
      const uint32_t cnt = __popc(f);
      if (st.bag_len + cnt <= a.bag_cap) {
        if (to_future) a.bag[st.bag_len + __popc(f & ((1u << lane) - 1u))] = id;
        st.bag_len += cnt;
      } else {
        st.lossy = true;   // the bag is abandoned; a rescan of all tuples rebuilds the future set if ever needed
      }
    }
  }
  __syncwarp();
}

// One relaxation candidate per lane; `first` = lanes of the expansion group that
// precedes the other active lanes in the reference's expansion order (match arcs
// before epsilon arcs, :182-278) — only used to number newly discovered tuples.
template <class Lhs>
__device__ inline void relax_warp(const SearchParams& p, const Lhs& lhs, const WarpArena& a, WarpState& st, unsigned lane,
                                  uint32_t cur_id, double cur_dist, bool active, const Cand& c, unsigned first,
                                  unsigned long long& future_min) {
  const unsigned act = __ballot_sync(kFull, active);
  if (act == 0) return;
  st.relax_calls += __popc(act);
  const unsigned long long mkey = active ? c.key : (0xFFFFFFFFFFFFFF00ull | lane);
  const unsigned peers = __match_any_sync(kFull, mkey);
  const bool leader = active && ((unsigned)(__ffs(peers) - 1) == lane);
  uint32_t pos = 0; bool found = false;
  TupleSlot s;
  s.key = c.key; s.dist = d_inf(); s.id_flags = 0; s.prev_id = kNone; s.rhs_arc = kNone; s.lhs_arc = kNone;
  if (leader) found = w_probe(a, c.key, pos, s);
  const unsigned newmask = __ballot_sync(kFull, leader && !found);
  const uint32_t n_new = __popc(newmask);
  if (st.n_tuples + n_new > a.tuple_cap) { st.overflow = true; return; }
  double old_dist = d_inf();
  if (leader) {
    if (!found) {
      pos = w_claim(a, c.key, pos);
      const unsigned lt = (1u << lane) - 1u;
      const bool in_first = (first >> lane) & 1u;
      const uint32_t rank = in_first ? __popc(newmask & first & lt) : (__popc(newmask & first) + __popc(newmask & ~first & lt));
      const uint32_t my_id = st.n_tuples + rank;   // discovery order == reference expansion order
      a.key_of[my_id] = c.key;
      s.key = c.key; s.dist = d_inf(); s.id_flags = my_id; s.prev_id = kNone; s.rhs_arc = kNone; s.lhs_arc = kNone;
    } else {
      old_dist = s.dist;
    }
  }
  st.n_tuples += n_new;
  const double nd = d_times(cur_dist, c.ew);
  bool changed = false;
  uint32_t s_il = 0, s_ol = 0;
  if (leader) {
    if (s.prev_id == cur_id) backptr_labels(p, lhs, s, s_il, s_ol);
    if (take_rule(nd, cur_id, c.il, c.ol, s.dist, s.prev_id, s_il, s_ol)) {
      s.dist = nd; s.prev_id = cur_id; s.rhs_arc = c.rhs_arc; s.lhs_arc = c.lhs_arc; s_il = c.il; s_ol = c.ol; changed = true;
    }
  }
  // fold further candidates of the same target in lane order (registers only)
  unsigned rest = leader ? (peers & ~(1u << lane)) : 0u;
  while (__any_sync(kFull, rest != 0)) {
    const int src = rest ? (__ffs(rest) - 1) : (int)lane;
    const double pnd = __shfl_sync(kFull, nd, src);
    const uint32_t pil = __shfl_sync(kFull, c.il, src), pol = __shfl_sync(kFull, c.ol, src);
    const uint32_t pl = __shfl_sync(kFull, c.lhs_arc, src), pr = __shfl_sync(kFull, c.rhs_arc, src);
    if (rest) {
      rest &= rest - 1;
      if (take_rule(pnd, cur_id, pil, pol, s.dist, s.prev_id, s_il, s_ol)) {
        s.dist = pnd; s.prev_id = cur_id; s.rhs_arc = pr; s.lhs_arc = pl; s_il = pil; s_ol = pol; changed = true;
      }
    }
  }
  if (leader && (changed || !found)) a.table[pos] = s;
  const bool need = leader && (!found || s.dist < old_dist);
  queue_insert(a, st, lane, need, s.id_flags, s.dist, future_min);
}

// Build the radix heap from the bag (or, if the bag was abandoned, from all tuples).
__device__ inline void build_radix(const WarpArena& a, WarpState& st, unsigned lane) {
  st.sorted = true;
  const uint32_t n = st.lossy ? st.n_tuples : st.bag_len;
  for (uint32_t base = 0; base < n && !st.overflow; base += 32) {
    const uint32_t j = base + lane;
    bool valid = false; uint32_t id = 0; unsigned long long k = 0;
    if (j < n) {
      id = st.lossy ? j : a.bag[j];
      const uint32_t pos = w_probe_pos(a, a.key_of[id]);
      k = (unsigned long long)__double_as_longlong(a.table[pos].dist);
      valid = k > st.last;
    }
    if (__any_sync(kFull, valid)) bucket_push(a, st, lane, valid, id, valid ? bucket_of(k, st.last) : 1u);
  }
  st.bag_len = 0; st.lossy = false;
}

// Advance to the next distance level.  Returns false when the search is finished
// (no valid entry left, or no remaining tuple can change the result).
__device__ inline bool advance_level(const SearchParams& p, const WarpArena& a, WarpState& st, unsigned lane, bool have_best,
                                     double best_total, unsigned long long future_min) {
  if (!st.sorted) {
    const unsigned long long fm = warp_min_u64(future_min);
    if (fm == ~0ull) return false;   // nothing was ever pushed beyond the levels already done
    // lower bound of every remaining distance: cannot reach or tie the best total -> done
    if (!p.exhaustive && have_best && __longlong_as_double((long long)fm) > best_total) return false;
    build_radix(a, st, lane);
    if (st.overflow) return false;
  }
  while (st.occupied) {
    const uint32_t b0 = __ffsll((long long)st.occupied);   // bucket number 1..64
    const uint2 hb = a.bucket[b0 - 1];
    // pass 1: smallest valid key in the bucket
    unsigned long long m = ~0ull;
    {
      uint32_t c = hb.x, cnt = hb.y;
      while (c != kNoChunk) {
        const uint32_t* ch = a.chunks + (uint64_t)c * 32;
        const uint32_t next = ch[0];
        if (lane < cnt) {
          const uint32_t id = ch[1 + lane];
          const uint32_t pos = w_probe_pos(a, a.key_of[id]);
          const unsigned long long k = (unsigned long long)__double_as_longlong(a.table[pos].dist);
          if (k > st.last && bucket_of(k, st.last) == b0 && k < m) m = k;
        }
        c = next; cnt = kChunkIds;
      }
      m = warp_min_u64(m);
    }
    st.occupied &= ~(1ull << (b0 - 1));
    __syncwarp();
    if (m == ~0ull) {   // only stale entries: recycle the chunks
      uint32_t c = hb.x;
      while (c != kNoChunk) {
        const uint32_t next = a.chunks[(uint64_t)c * 32];
        __syncwarp();
        if (lane == 0) a.chunks[(uint64_t)c * 32] = st.free_head;
        st.free_head = c;
        __syncwarp();
        c = next;
      }
      continue;
    }
    if (!p.exhaustive && have_best && __longlong_as_double((long long)m) > best_total) return false;
    // pass 2: redistribute relative to the new level key m
    const unsigned long long old_last = st.last;
    st.last = m;
    uint32_t c = hb.x, cnt = hb.y;
    while (c != kNoChunk && !st.overflow) {
      const uint32_t* ch = a.chunks + (uint64_t)c * 32;
      const uint32_t next = ch[0];
      bool valid = false; uint32_t id = 0; unsigned long long k = 0;
      if (lane < cnt) {
        id = ch[1 + lane];
        const uint32_t pos = w_probe_pos(a, a.key_of[id]);
        k = (unsigned long long)__double_as_longlong(a.table[pos].dist);
        valid = k > old_last && bucket_of(k, old_last) == b0;
      }
      ready_insert(a, lane, valid && k == m, id);
      const bool tb = valid && k != m;
      if (__any_sync(kFull, tb)) bucket_push(a, st, lane, tb, id, tb ? bucket_of(k, m) : 1u);
      __syncwarp();
      if (lane == 0) a.chunks[(uint64_t)c * 32] = st.free_head;   // recycle
      st.free_head = c;
      __syncwarp();
      c = next; cnt = kChunkIds;
    }
    return !st.overflow;
  }
  return false;
}

// Expansion of one popped tuple, general left operand: the reference's four groups
// in order (compose-shortest-path.zig:182-365), 32 arcs per step.
template <class Lhs>
__device__ inline void expand_general(const SearchParams& p, const Lhs& lhs, const WarpArena& a, WarpState& st, unsigned lane,
                                      uint32_t cur_id, double cur_dist, uint32_t s1, uint32_t s2, uint32_t filt,
                                      unsigned long long& future_min) {
  const DevFstView& F = p.fst;
  const Group<32> g;
  const uint4 rec = __ldg(&F.state_rec[s2]);
  const uint32_t n1 = lhs.n_arcs(s1);
  const uint32_t a1base = lhs.arc_base(s1);
  for (uint32_t k = 0; k < n1 && !st.overflow; k++) {
    uint32_t il1, ol1, nx1; double w1;
    lhs.arc(a1base + k, il1, ol1, w1, nx1);
    if (ol1 == 0) continue;
    uint32_t lo, hi;
    equal_range(g, F.ilabel, rec.x, rec.z, ol1, lo, hi);
    for (uint32_t cb = lo; cb < hi && !st.overflow; cb += 32) {
      const bool active = cb + lane < hi;
      Cand c; c.key = 0; c.ew = 0; c.il = il1; c.ol = 0; c.lhs_arc = a1base + k; c.rhs_arc = cb + lane;
      if (active) {
        const uint4 pl = __ldg(&F.payload[c.rhs_arc]);
        c.ol = pl.x;
        c.ew = d_times(w1, __hiloint2double((int)pl.w, (int)pl.z));
        c.key = pack_key(nx1, pl.y, 0);
      }
      relax_warp(p, lhs, a, st, lane, cur_id, cur_dist, active, c, kFull, future_min);
    }
  }
  if (filt != 1) {
    for (uint32_t kb = 0; kb < n1 && !st.overflow; kb += 32) {
      const uint32_t k = kb + lane;
      bool active = false;
      Cand c; c.key = 0; c.ew = 0; c.il = 0; c.ol = 0; c.lhs_arc = a1base + k; c.rhs_arc = kNone;
      if (k < n1) {
        uint32_t il1, ol1, nx1; double w1;
        lhs.arc(a1base + k, il1, ol1, w1, nx1);
        if (ol1 == 0) { active = true; c.il = il1; c.ew = w1; c.key = pack_key(nx1, s2, filt == 0 ? 2u : filt); }
      }
      relax_warp(p, lhs, a, st, lane, cur_id, cur_dist, active, c, kFull, future_min);
    }
  }
  if (filt != 2) {
    for (uint32_t cb = rec.x; cb < rec.y && !st.overflow; cb += 32) {
      const bool active = cb + lane < rec.y;
      Cand c; c.key = 0; c.ew = 0; c.il = 0; c.ol = 0; c.lhs_arc = kNone; c.rhs_arc = cb + lane;
      if (active) {
        const uint4 pl = __ldg(&F.payload[c.rhs_arc]);
        c.ol = pl.x;
        c.ew = __hiloint2double((int)pl.w, (int)pl.z);
        c.key = pack_key(s1, pl.y, filt == 0 ? 1u : filt);
      }
      relax_warp(p, lhs, a, st, lane, cur_id, cur_dist, active, c, kFull, future_min);
    }
  }
  if (filt == 0 && rec.y > rec.x) {
    for (uint32_t k = 0; k < n1 && !st.overflow; k++) {
      uint32_t il1, ol1, nx1; double w1;
      lhs.arc(a1base + k, il1, ol1, w1, nx1);
      if (ol1 != 0) continue;
      for (uint32_t cb = rec.x; cb < rec.y && !st.overflow; cb += 32) {
        const bool active = cb + lane < rec.y;
        Cand c; c.key = 0; c.ew = 0; c.il = il1; c.ol = 0; c.lhs_arc = a1base + k; c.rhs_arc = cb + lane;
        if (active) {
          const uint4 pl = __ldg(&F.payload[c.rhs_arc]);
          c.ol = pl.x;
          c.ew = d_times(w1, __hiloint2double((int)pl.w, (int)pl.z));
          c.key = pack_key(nx1, pl.y, 0);
        }
        relax_warp(p, lhs, a, st, lane, cur_id, cur_dist, active, c, kFull, future_min);
      }
    }
  }
}

// Expansion specialised for a byte-string left operand (one arc per state, no
// epsilons, unit weight) and a transducer state of <= 32 arcs: the match arcs and
// the input-epsilon arcs of the state are relaxed in ONE step.  The lanes hold the
// arcs in frozen order (epsilon prefix first), the discovery ids follow the
// reference's order (match arcs :182-202 before epsilon arcs :254-278).
__device__ __forceinline__ bool expand_bytes_small(const SearchParams& p, const LhsBytes& lhs, const WarpArena& a, WarpState& st,
                                                   unsigned lane, uint32_t cur_id, double cur_dist, uint32_t s1, uint32_t s2,
                                                   uint32_t filt, unsigned long long& future_min) {
  const DevFstView& F = p.fst;
  const uint4 rec = __ldg(&F.state_rec[s2]);
  const uint32_t deg = rec.z - rec.x;
  if (deg > 32) return false;
  const uint32_t arc = rec.x + lane;
  const bool valid = lane < deg;
  uint32_t il = 0xFFFFFFFFu;
  uint4 pl = make_uint4(0, 0, 0, 0);
  if (valid) { il = __ldg(F.ilabel + arc); pl = __ldg(&F.payload[arc]); }
  const uint32_t x = s1 < lhs.len ? (uint32_t)__ldg(lhs.s + s1) + 1u : 0xFFFFFFFEu;
  const bool is_match = valid && il == x;
  const bool is_eps = valid && arc < rec.y;       // filter is never 2 for an epsilon-free left operand
  const double w2 = __hiloint2double((int)pl.w, (int)pl.z);
  Cand c;
  c.ol = pl.x; c.rhs_arc = arc;
  c.il = is_match ? x : 0u;
  c.lhs_arc = is_match ? s1 : kNone;
  c.ew = is_match ? d_times(0.0, w2) : w2;
  c.key = is_match ? pack_key(s1 + 1u, pl.y, 0u) : pack_key(s1, pl.y, filt == 0 ? 1u : filt);
  const unsigned first = __ballot_sync(kFull, is_match);
  relax_warp(p, lhs, a, st, lane, cur_id, cur_dist, is_match || is_eps, c, first, future_min);
  return true;
}

template <class Lhs>
__device__ inline int32_t search_warp(const SearchParams& p, const Lhs& lhs, const WarpArena& a, uint32_t* out_path_len,
                                      uint64_t* out_pool_off, double* out_final_w, uint32_t* out_n_tuples,
                                      unsigned long long* out_relax) {
  const unsigned lane = threadIdx.x & 31u;
  WarpState st;
  st.n_tuples = 0; st.chunk_next = 0; st.free_head = kNoChunk; st.bag_len = 0; st.occupied = 0; st.last = 0; st.relax_calls = 0;
  st.overflow = false; st.sorted = false; st.lossy = false;
  unsigned long long future_min = ~0ull;
  *out_path_len = 0; *out_pool_off = 0; *out_final_w = d_inf(); *out_n_tuples = 0; *out_relax = 0;
  const DevFstView& F = p.fst;
  if (lhs.start() == kNone || F.start == kNone) return kStNoPath;

  // initial tuple: id 0, dist One, ready at level 0 (compose-shortest-path.zig:146-153)
  if (lane == 0) {
    const unsigned long long k0 = pack_key(lhs.start(), F.start, 0);
    const uint32_t pos = w_probe_pos(a, k0);
    TupleSlot s; s.key = k0; s.dist = 0.0; s.id_flags = 0; s.prev_id = kNone; s.rhs_arc = kNone; s.lhs_arc = kNone;
    a.table[pos] = s;
    a.key_of[0] = k0;
  }
  st.n_tuples = 1;
  ready_insert(a, lane, lane == 0, 0u);
  __syncwarp();

  bool have_best = false; uint32_t best_id = 0; double best_fw = d_inf(), best_total = d_inf();
  double cur_dist = 0.0;

  for (;;) {
    if (st.overflow) break;
    if (ready_empty(a)) {
      if (!advance_level(p, a, st, lane, have_best, best_total, future_min)) break;
      cur_dist = __longlong_as_double((long long)st.last);
      continue;
    }
    const uint32_t cur_id = ready_pop(a, lane);
    const unsigned long long ckey = a.key_of[cur_id];
    const uint32_t s1 = (uint32_t)(ckey >> 34), s2 = (uint32_t)(ckey >> 2), filt = (uint32_t)(ckey & 3u);

    // final check (compose-shortest-path.zig:165-179)
    const double fw1 = lhs.final_w(s1);
    if (!d_isinf(fw1)) {
      const double fw2 = F.final_w[s2];
      if (!d_isinf(fw2)) {
        const double final_w = d_times(fw1, fw2);
        const double total = d_times(cur_dist, final_w);
        if (!have_best || total < best_total || (total == best_total && cur_id < best_id)) {
          have_best = true; best_id = cur_id; best_fw = final_w; best_total = total;
        }
      }
    }
    bool done = false;
    if constexpr (std::is_same<Lhs, LhsBytes>::value)
      done = expand_bytes_small(p, lhs, a, st, lane, cur_id, cur_dist, s1, s2, filt, future_min);
    if (!done) expand_general(p, lhs, a, st, lane, cur_id, cur_dist, s1, s2, filt, future_min);
  }

  int32_t status = kStPath;
  uint32_t plen = 0;
  unsigned long long poff = 0;
  uint32_t* scratch = a.bag;   // future set is dead now; bag_cap >= tuple_cap by construction
  if (st.overflow) {
    status = kStRetry;
  } else if (!have_best) {
    status = kStNoPath;                                               // :368-370
  } else {
    if (lane == 0) {                                                  // :372-380 back-track
      uint32_t cur = best_id;
      while (cur != 0) {
        const uint32_t sl = w_probe_pos(a, a.key_of[cur]);
        const uint32_t prev = a.table[sl].prev_id;
        if (prev == kNone) { status = kStNoPath; break; }             // :375-377
        if (plen >= st.n_tuples) { status = kStCycle; break; }        // hazard H1 (reference: out of memory)
        scratch[plen++] = sl;
        cur = prev;
      }
      if (status == kStPath && plen > 0) {
        poff = atomicAdd(p.pool_cursor, (unsigned long long)plen);
        if (poff + plen > p.pool_cap) status = kStRetry;
      }
    }
    __syncwarp();
    status = __shfl_sync(kFull, status, 0); plen = __shfl_sync(kFull, plen, 0); poff = __shfl_sync(kFull, poff, 0);
    if (status == kStPath) {
      for (uint32_t i = lane; i < plen; i += 32) {
        const TupleSlot s = a.table[scratch[i]];
        uint32_t il = 0, ol = 0; double w1 = 0.0, w2 = 0.0, w;
        if (s.lhs_arc != kNone) { uint32_t o, n; lhs.arc(s.lhs_arc, il, o, w1, n); }
        if (s.rhs_arc != kNone) { const uint4 pl = __ldg(&F.payload[s.rhs_arc]); ol = pl.x; w2 = __hiloint2double((int)pl.w, (int)pl.z); }
        if (s.lhs_arc != kNone && s.rhs_arc != kNone) w = d_times(w1, w2);
        else if (s.lhs_arc != kNone) w = w1;
        else w = w2;
        PoolArc pa; pa.ilabel = il; pa.olabel = ol; pa.weight = w;
        p.pool[poff + i] = pa;
      }
    } else {
      plen = 0;
    }
  }
  // Restore the arena invariants for the next string: table empty, bitmaps zero.
  // Two phases: resolve every tuple's slot first (probing needs intact chains), then clear.
  __syncwarp();
  uint32_t* slot_tmp = reinterpret_cast<uint32_t*>(a.key_of);
  for (uint32_t base = 0; base < st.n_tuples; base += 32) {
    const uint32_t i = base + lane;
    uint32_t sl = 0;
    if (i < st.n_tuples) sl = w_probe_pos(a, a.key_of[i]);
    __syncwarp();                               // all keys of this stripe are read before any is overwritten
    if (i < st.n_tuples) slot_tmp[i] = sl;      // aliases key_of[i/2]: only stripes already resolved
    __syncwarp();
  }
  for (uint32_t i = lane; i < st.n_tuples; i += 32) a.table[slot_tmp[i]].key = kEmptyKey;
  for (uint32_t i = lane; i < (st.n_tuples + 63) / 64; i += 32) a.l0[i] = 0;
  {
    const uint32_t u0 = (st.n_tuples + 63) / 64, u1 = (u0 + 31) / 32, u2 = (u1 + 31) / 32;
    for (uint32_t i = lane; i < u1; i += 32) a.l1[i] = 0;
    for (uint32_t i = lane; i < u2; i += 32) a.l2[i] = 0;
    if (lane < 2) a.l3[lane] = 0;
  }
  __syncwarp();
  *out_path_len = plen; *out_pool_off = poff; *out_final_w = (status == kStPath) ? best_fw : d_inf();
  *out_n_tuples = st.n_tuples; *out_relax = st.relax_calls;
  return status;
}

__device__ inline WarpArena warp_arena_at(const SearchParams& p, uint32_t slot_idx, uint32_t* smem_warp) {
  const WarpLayout L = warp_layout(p.hash_cap, p.tuple_cap, p.heap_cap, p.bag_cap);
  uint8_t* base = p.arena + (uint64_t)slot_idx * p.arena_stride;
  WarpArena a;
  a.table = reinterpret_cast<TupleSlot*>(base + L.off_table);
  a.key_of = reinterpret_cast<unsigned long long*>(base + L.off_keyof);
  a.l0 = reinterpret_cast<unsigned long long*>(base + L.off_l0);
  a.bag = reinterpret_cast<uint32_t*>(base + L.off_bag);
  a.chunks = reinterpret_cast<uint32_t*>(base + L.off_chunks);
  a.bucket = reinterpret_cast<uint2*>(smem_warp);
  a.l1 = smem_warp + 128; a.l2 = a.l1 + L.n1; a.l3 = a.l2 + L.n2;
  a.hash_cap = p.hash_cap; a.tuple_cap = p.tuple_cap; a.chunk_cap = p.heap_cap; a.bag_cap = p.bag_cap;
  a.n0 = L.n0; a.n1 = L.n1; a.n2 = L.n2; a.n3 = L.n3;
  // shared-memory bitmap levels start empty
  const unsigned lane = threadIdx.x & 31u;
  for (uint32_t i = lane; i < L.smem_words - 128; i += 32) a.l1[i] = 0;
  __syncwarp();
  return a;
}

// Persistent batch kernel: every warp pulls strings from a global queue.
__global__ void __launch_bounds__(128) csp_batch_warp_kernel(SearchParams p) {
  extern __shared__ __align__(16) uint32_t smem_all[];
  const unsigned lane = threadIdx.x & 31u;
  const uint32_t wib = threadIdx.x >> 5;
  const uint32_t gslot = blockIdx.x * (blockDim.x >> 5) + wib;
  const WarpLayout L = warp_layout(p.hash_cap, p.tuple_cap, p.heap_cap, p.bag_cap);
  WarpArena a = warp_arena_at(p, gslot, smem_all + (size_t)wib * L.smem_words);
  unsigned long long relax_total = 0, tuple_total = 0;
  for (;;) {
    uint32_t item = 0;
    if (lane == 0) item = atomicAdd(p.queue_head, 1u);
    item = __shfl_sync(kFull, item, 0);
    if (item >= p.n_items) break;
    const uint32_t idx = p.order ? p.order[item] : item;
    LhsBytes lhs; lhs.s = p.bytes + p.offsets[idx]; lhs.len = (uint32_t)(p.offsets[idx + 1] - p.offsets[idx]);
    uint32_t plen = 0; uint64_t poff = 0; double fw = d_inf(); uint32_t nt = 0; unsigned long long nr = 0;
    const int32_t pre = p.skip ? p.skip[idx] : kStPath;   // an earlier pipeline stage failed: pass its status through
    const int32_t status = pre != kStPath ? pre : search_warp(p, lhs, a, &plen, &poff, &fw, &nt, &nr);
    if (lane == 0) {
      p.status[idx] = status; p.path_len[idx] = plen; p.pool_off[idx] = poff; p.final_w[idx] = fw; p.n_tuples[idx] = nt;
    }
    relax_total += nr; tuple_total += nt;
  }
  if (lane == 0) {
    if (relax_total) atomicAdd(p.relax_counter, relax_total);
    if (tuple_total) atomicAdd(p.tuple_counter, tuple_total);
  }
}

// One general left operand (the fst_compose_frozen_shortest_path drop-in).
__global__ void __launch_bounds__(32) csp_general_warp_kernel(SearchParams p) {
  extern __shared__ __align__(16) uint32_t smem_all[];
  const unsigned lane = threadIdx.x & 31u;
  WarpArena a = warp_arena_at(p, 0, smem_all);
  LhsCsr lhs; lhs.v = p.lhs;
  uint32_t plen; uint64_t poff; double fw; uint32_t nt; unsigned long long nr;
  const int32_t status = search_warp(p, lhs, a, &plen, &poff, &fw, &nt, &nr);
  if (lane == 0) {
    p.status[0] = status; p.path_len[0] = plen; p.pool_off[0] = poff; p.final_w[0] = fw; p.n_tuples[0] = nt;
    atomicAdd(p.relax_counter, nr); atomicAdd(p.tuple_counter, (unsigned long long)nt);
  }
}

}  // namespace fstb200
