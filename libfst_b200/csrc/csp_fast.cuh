// Fast exact search for non-negative weights (the common case).
//
// With non-negative weights the reference's pops are monotone in distance, so its
// (dist, id) min-heap (compose-shortest-path.zig:55-61) can be replaced, without
// changing the pop sequence, by
//   * a READY SET: the unsettled tuples whose tentative distance equals the
//     current level distance, kept as a hierarchical bitmap indexed by discovery
//     id; "pop" = find-first-set = the smallest id, exactly the heap's tie rule;
//   * FUTURE BUCKETS: a radix heap over the IEEE-754 bit pattern of the distance
//     (monotone keys; 64 buckets by the highest bit in which a key differs from the
//     last popped key).  Entries are ids only; an entry is valid iff the tuple's
//     CURRENT distance still maps to the bucket it sits in (a lowered tuple always
//     has a second entry in the right place, so stale ones are simply dropped —
//     the reference skips them at pop time, :162).
// The reference's `settled` flag is not needed: a tuple at the current level is
// either in the ready set or already expanded, and in both cases a tie relaxation
// (:115-126) changes the back-pointer only; strictly better relaxations can only
// hit unsettled tuples.  See DESIGN.md §"exactness of the monotone queue".
#pragma once
#include "csp_kernels.cuh"

namespace fstb200 {

constexpr uint32_t kChunkIds = 31;           // ids per 128-byte chunk (word 0 = next chunk)
constexpr uint32_t kNoChunk = 0xFFFFFFFFu;
constexpr uint32_t kMaxFastTuples = 8u << 20;  // 64^3 * 32 bitmap capacity

struct FastLayout {
  uint64_t off_table, off_keyof, off_l0, off_l1, off_l2, off_chunks, total;
  uint32_t n0, n1, n2;
};
__host__ __device__ inline FastLayout fast_layout(uint32_t hash_cap, uint32_t tuple_cap, uint32_t chunk_cap) {
  FastLayout L;
  auto al = [](uint64_t x) { return (x + 127) & ~127ull; };
  L.n0 = (tuple_cap + 63) / 64; L.n1 = (L.n0 + 63) / 64; L.n2 = (L.n1 + 63) / 64;
  L.off_table = 0;
  L.off_keyof = al((uint64_t)hash_cap * sizeof(TupleSlot));
  L.off_l0 = L.off_keyof + al((uint64_t)tuple_cap * 8);
  L.off_l1 = L.off_l0 + al((uint64_t)L.n0 * 8);
  L.off_l2 = L.off_l1 + al((uint64_t)L.n1 * 8);
  L.off_chunks = L.off_l2 + al((uint64_t)L.n2 * 8);
  L.total = (L.off_chunks + (uint64_t)chunk_cap * 128 + 255) & ~255ull;
  return L;
}

struct FastArena {
  TupleSlot* table;
  unsigned long long* key_of;   // id -> tuple key
  unsigned long long* l0; unsigned long long* l1; unsigned long long* l2;   // ready bitmap levels
  uint32_t* chunks;             // chunk c = chunks[c*32 .. c*32+31]; word 0 = next
  uint32_t hash_cap, tuple_cap, chunk_cap;
  uint32_t n0, n1, n2;
};
__device__ inline FastArena fast_arena_at(const SearchParams& p, uint32_t slot_idx) {
  FastLayout L = fast_layout(p.hash_cap, p.tuple_cap, p.heap_cap);
  uint8_t* base = p.arena + (uint64_t)slot_idx * p.arena_stride;
  FastArena a;
  a.table = reinterpret_cast<TupleSlot*>(base + L.off_table);
  a.key_of = reinterpret_cast<unsigned long long*>(base + L.off_keyof);
  a.l0 = reinterpret_cast<unsigned long long*>(base + L.off_l0);
  a.l1 = reinterpret_cast<unsigned long long*>(base + L.off_l1);
  a.l2 = reinterpret_cast<unsigned long long*>(base + L.off_l2);
  a.chunks = reinterpret_cast<uint32_t*>(base + L.off_chunks);
  a.hash_cap = p.hash_cap; a.tuple_cap = p.tuple_cap; a.chunk_cap = p.heap_cap;
  a.n0 = L.n0; a.n1 = L.n1; a.n2 = L.n2;
  return a;
}

// Arena initialisation: table keys empty, bitmap zero (run when the layout changes).
__global__ void fast_arena_init_kernel(uint8_t* arena, uint64_t stride, uint32_t n_arenas, uint32_t hash_cap, uint32_t tuple_cap,
                                       uint32_t chunk_cap) {
  FastLayout L = fast_layout(hash_cap, tuple_cap, chunk_cap);
  const uint64_t table_words = (uint64_t)hash_cap * 4;              // 8-byte words in the table
  const uint64_t bitmap_words = (L.off_chunks - L.off_l0) / 8;
  const uint64_t per = table_words + bitmap_words;
  const uint64_t total = per * n_arenas;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t ar = i / per, w = i % per;
    unsigned long long* base = reinterpret_cast<unsigned long long*>(arena + ar * stride);
    if (w < table_words) base[w] = ((w & 3) == 0) ? kEmptyKey : 0ull;
    else base[L.off_l0 / 8 + (w - table_words)] = 0ull;
  }
}

__device__ __forceinline__ uint32_t fast_home(const FastArena& a, unsigned long long key) {
  return (uint32_t)(((unsigned long long)hash_key(key) * a.hash_cap) >> 32);
}
__device__ __forceinline__ bool fast_probe(const FastArena& a, unsigned long long key, uint32_t& pos) {
  uint32_t i = fast_home(a, key);
  for (;;) {
    unsigned long long k = a.table[i].key;
    if (k == key) { pos = i; return true; }
    if (k == kEmptyKey) { pos = i; return false; }
    if (++i == a.hash_cap) i = 0;
  }
}
__device__ __forceinline__ uint32_t fast_claim(const FastArena& a, unsigned long long key, uint32_t pos) {
  for (;;) {
    unsigned long long old = atomicCAS(&a.table[pos].key, kEmptyKey, key);
    if (old == kEmptyKey) return pos;
    if (++pos == a.hash_cap) pos = 0;
  }
}

// ── radix-heap bucket index of key k relative to the last popped key ──
__device__ __forceinline__ uint32_t bucket_of(unsigned long long k, unsigned long long last) {
  return 64u - (uint32_t)__clzll((long long)(k ^ last));   // 1..64 for k != last (0 never stored)
}

// Per-group queue state.  Bucket heads live in shared memory (64 x {chunk, count}).
struct FastQueue {
  uint2* bucket;               // smem: [64] {head chunk, ids in head chunk}
  unsigned long long occupied; // bit b-1 set: bucket b non-empty            (uniform)
  uint32_t top;                // top level of the ready bitmap (<= 32 bits) (uniform)
  uint32_t chunk_next;         // bump allocator                             (uniform)
  uint32_t free_head;          // free-list of recycled chunks               (uniform)
  unsigned long long last;     // key of the current level                   (uniform)
};

template <int G>
__device__ inline uint32_t chunk_alloc(const Group<G>& g, const FastArena& a, FastQueue& q, bool& overflow) {
  // uniform: every lane computes the same result (free list head is read uniformly)
  uint32_t c;
  if (q.free_head != kNoChunk) {
    c = q.free_head;
    q.free_head = a.chunks[(uint64_t)c * 32];
  } else if (q.chunk_next < a.chunk_cap) {
    c = q.chunk_next++;
  } else {
    overflow = true; c = 0;
  }
  return c;
}

// Insert ids into the ready bitmap (collective; `active` lanes carry distinct or equal ids).
template <int G>
__device__ inline void ready_insert(const Group<G>& g, const FastArena& a, FastQueue& q, bool active, uint32_t id) {
  if (!g.any(active)) return;
  uint32_t wi = id >> 6;
  unsigned long long bit = 1ull << (id & 63u);
  unsigned peers = g.match_any(active ? (unsigned long long)wi : (0xFFFFFFFF00000000ull | g.lane));
  bool leader = active && ((unsigned)(__ffs(peers) - 1) == g.lane);
  unsigned wmask = peers << g.base;
  uint32_t lo = __reduce_or_sync(wmask, (uint32_t)bit);
  uint32_t hi = __reduce_or_sync(wmask, (uint32_t)(bit >> 32));
  bool up = false;
  if (leader) {
    unsigned long long old = a.l0[wi];
    a.l0[wi] = old | ((unsigned long long)hi << 32) | lo;
    up = (old == 0);
  }
  unsigned m = g.ballot(up);
  if (m) {
    uint32_t top = q.top;
    while (m) {
      int src = __ffs(m) - 1; m &= m - 1;
      uint32_t w = g.shfl(wi, src);
      if (g.lane == 0) {
        unsigned long long o1 = a.l1[w >> 6];
        a.l1[w >> 6] = o1 | (1ull << (w & 63u));
        if (o1 == 0) {
          uint32_t w1 = w >> 6;
          unsigned long long o2 = a.l2[w1 >> 6];
          a.l2[w1 >> 6] = o2 | (1ull << (w1 & 63u));
          if (o2 == 0) top |= 1u << (w1 >> 6);
        }
      }
    }
    q.top = g.shfl(top, 0);
  }
  g.sync();
}

// Pop the smallest ready id (collective; requires q.top != 0).
template <int G>
__device__ inline uint32_t ready_pop(const Group<G>& g, const FastArena& a, FastQueue& q) {
  uint32_t i2 = __ffs(q.top) - 1;
  unsigned long long w2 = a.l2[i2];
  uint32_t i1 = i2 * 64 + (__ffsll((long long)w2) - 1);
  unsigned long long w1 = a.l1[i1];
  uint32_t i0 = i1 * 64 + (__ffsll((long long)w1) - 1);
  unsigned long long w0 = a.l0[i0];
  uint32_t b0 = __ffsll((long long)w0) - 1;
  uint32_t id = i0 * 64 + b0;
  w0 &= w0 - 1;
  bool z0 = (w0 == 0);
  unsigned long long nw1 = w1 & ~(1ull << (i0 & 63u));
  bool z1 = z0 && nw1 == 0;
  unsigned long long nw2 = w2 & ~(1ull << (i1 & 63u));
  bool z2 = z1 && nw2 == 0;
  if (g.lane == 0) {
    a.l0[i0] = w0;
    if (z0) a.l1[i1] = nw1;
    if (z1) a.l2[i2] = nw2;
  }
  if (z2) q.top &= ~(1u << i2);
  g.sync();
  return id;
}

// Append ids to future buckets (collective).  `b` in 1..64.
template <int G>
__device__ inline void bucket_push(const Group<G>& g, const FastArena& a, FastQueue& q, bool active, uint32_t id, uint32_t b,
                                   bool& overflow) {
  unsigned m = g.ballot(active);
  while (m) {
    int first = __ffs(m) - 1;
    uint32_t bb = g.shfl(b, first);
    unsigned same = g.ballot(active && b == bb);
    m &= ~same;
    uint32_t k = __popc(same);
    uint32_t rank = __popc(same & g.lt_mask());
    uint2 st = q.bucket[bb - 1];
    bool empty = !((q.occupied >> (bb - 1)) & 1ull);
    uint32_t head = empty ? kNoChunk : st.x, cnt = empty ? kChunkIds : st.y;
    uint32_t space = kChunkIds - cnt;
    bool mine = active && b == bb;
    if (mine && rank < space) a.chunks[(uint64_t)head * 32 + 1 + cnt + rank] = id;
    uint32_t left = k > space ? k - space : 0;
    uint32_t done = k - left;
    while (left > 0) {   // k <= 32 > 31 possible only when space == 0 and k == 32: two chunks
      uint32_t c = chunk_alloc(g, a, q, overflow);
      if (overflow) return;
      uint32_t take = left < kChunkIds ? left : kChunkIds;
      if (g.lane == 0) a.chunks[(uint64_t)c * 32] = head;
      if (mine && rank >= done && rank < done + take) a.chunks[(uint64_t)c * 32 + 1 + (rank - done)] = id;
      head = c; cnt = 0; done += take; left -= take;
      cnt = take;
    }
    if (k <= space) cnt = cnt + k;
    g.sync();
    if (g.lane == 0) q.bucket[bb - 1] = make_uint2(head, cnt);
    q.occupied |= 1ull << (bb - 1);
    g.sync();
  }
}

struct FastState {
  uint32_t n_tuples;
  bool overflow;
  unsigned long long relax_calls;
};

// Relax up to G candidates of one expansion (monotone mode).  Same folding as
// relax_chunk in csp_kernels.cuh; queue actions only for new / strictly lowered targets.
template <int G, class Lhs>
__device__ inline void relax_chunk_fast(const Group<G>& g, const SearchParams& p, const Lhs& lhs, const FastArena& a, FastQueue& q,
                                        FastState& st, uint32_t cur_id, double cur_dist, bool active, const Cand& c) {
  unsigned act = g.ballot(active);
  if (act == 0) return;
  st.relax_calls += __popc(act);
  unsigned long long mkey = active ? c.key : (0xFFFFFFFFFFFFFF00ull | g.lane);
  unsigned peers = g.match_any(mkey);
  bool leader = active && ((unsigned)(__ffs(peers) - 1) == g.lane);
  uint32_t pos = 0; bool found = false;
  if (leader) found = fast_probe(a, c.key, pos);
  unsigned newmask = g.ballot(leader && !found);
  uint32_t n_new = __popc(newmask);
  if (st.n_tuples + n_new > a.tuple_cap) { st.overflow = true; return; }
  TupleSlot s;
  double old_dist = d_inf();
  if (leader) {
    if (!found) {
      pos = fast_claim(a, c.key, pos);
      uint32_t my_id = st.n_tuples + __popc(newmask & g.lt_mask());
      a.key_of[my_id] = c.key;
      s.key = c.key; s.dist = d_inf(); s.id_flags = my_id; s.prev_id = kNone; s.rhs_arc = kNone; s.lhs_arc = kNone;
    } else {
      s = a.table[pos];
      old_dist = s.dist;
    }
  }
  st.n_tuples += n_new;
  double nd = d_times(cur_dist, c.ew);
  bool changed = false;
  uint32_t s_il = 0, s_ol = 0;
  if (leader) {
    if (s.prev_id == cur_id) backptr_labels(p, lhs, s, s_il, s_ol);
    if (take_rule(nd, cur_id, c.il, c.ol, s.dist, s.prev_id, s_il, s_ol)) {
      s.dist = nd; s.prev_id = cur_id; s.rhs_arc = c.rhs_arc; s.lhs_arc = c.lhs_arc; s_il = c.il; s_ol = c.ol; changed = true;
    }
  }
  unsigned rest = leader ? (peers & ~(1u << g.lane)) : 0u;
  while (g.any(rest != 0)) {
    int src = rest ? (__ffs(rest) - 1) : (int)g.lane;
    double pnd = g.shfl(nd, src);
    uint32_t pil = g.shfl(c.il, src), pol = g.shfl(c.ol, src);
    uint32_t pl = g.shfl(c.lhs_arc, src), pr = g.shfl(c.rhs_arc, src);
    if (rest) {
      rest &= rest - 1;
      if (take_rule(pnd, cur_id, pil, pol, s.dist, s.prev_id, s_il, s_ol)) {
        s.dist = pnd; s.prev_id = cur_id; s.rhs_arc = pr; s.lhs_arc = pl; s_il = pil; s_ol = pol; changed = true;
      }
    }
  }
  if (leader && (changed || !found)) a.table[pos] = s;
  // queue action: new tuple, or distance strictly lowered (see file header)
  bool need = leader && (!found || s.dist < old_dist);
  unsigned long long k = (unsigned long long)__double_as_longlong(s.dist);
  bool to_ready = need && k == q.last;
  bool to_bucket = need && k != q.last;
  uint32_t tid = s.id_flags;
  ready_insert(g, a, q, to_ready, tid);
  if (g.any(to_bucket)) bucket_push(g, a, q, to_bucket, tid, to_bucket ? bucket_of(k, q.last) : 1u, st.overflow);
  g.sync();
}

// Advance to the next distance level: redistribute the lowest non-empty bucket.
// Returns false when no valid entry remains anywhere (search finished).
template <int G>
__device__ inline bool advance_level(const Group<G>& g, const SearchParams& p, const FastArena& a, FastQueue& q, FastState& st,
                                     bool have_best, double best_total, bool& stop_early) {
  stop_early = false;
  while (q.occupied) {
    uint32_t b0 = __ffsll((long long)q.occupied);   // bucket number 1..64
    uint2 hb = q.bucket[b0 - 1];
    // pass 1: smallest valid key in the bucket
    unsigned long long m = ~0ull;
    {
      uint32_t c = hb.x, cnt = hb.y;
      while (c != kNoChunk) {
        const uint32_t* ch = a.chunks + (uint64_t)c * 32;
        uint32_t next = ch[0];
        for (uint32_t e = g.lane; e < cnt; e += G) {
          uint32_t id = ch[1 + e];
          uint32_t pos; fast_probe(a, a.key_of[id], pos);
          unsigned long long k = (unsigned long long)__double_as_longlong(a.table[pos].dist);
          if (k > q.last && bucket_of(k, q.last) == b0 && k < m) m = k;
        }
        c = next; cnt = kChunkIds;
      }
      for (int o = G / 2; o > 0; o >>= 1) { unsigned long long t = g.shfl(m, (int)(g.lane ^ o)); if (t < m) m = t; }
    }
    // detach the bucket
    q.occupied &= ~(1ull << (b0 - 1));
    g.sync();
    if (m == ~0ull) {
      // only stale entries: recycle the chunks
      uint32_t c = hb.x;
      while (c != kNoChunk) { uint32_t next = a.chunks[(uint64_t)c * 32]; g.sync(); if (g.lane == 0) a.chunks[(uint64_t)c * 32] = q.free_head; q.free_head = c; g.sync(); c = next; }
      continue;
    }
    double md = __longlong_as_double((long long)m);
    if (!p.exhaustive && have_best && md > best_total) { stop_early = true; return false; }
    // pass 2: redistribute relative to the new level key m
    const unsigned long long old_last = q.last;
    q.last = m;
    uint32_t c = hb.x, cnt = hb.y;
    while (c != kNoChunk && !st.overflow) {
      const uint32_t* ch = a.chunks + (uint64_t)c * 32;
      uint32_t next = ch[0];
      for (uint32_t eb = 0; eb < cnt; eb += G) {
        uint32_t e = eb + g.lane;
        bool valid = false; uint32_t id = 0; unsigned long long k = 0;
        if (e < cnt) {
          id = ch[1 + e];
          uint32_t pos; fast_probe(a, a.key_of[id], pos);
          k = (unsigned long long)__double_as_longlong(a.table[pos].dist);
          valid = k > old_last && bucket_of(k, old_last) == b0;
        }
        ready_insert(g, a, q, valid && k == m, id);
        bool tb = valid && k != m;
        if (g.any(tb)) bucket_push(g, a, q, tb, id, tb ? bucket_of(k, m) : 1u, st.overflow);
        if (st.overflow) break;
      }
      g.sync();
      if (g.lane == 0) a.chunks[(uint64_t)c * 32] = q.free_head;   // recycle
      q.free_head = c;
      g.sync();
      c = next; cnt = kChunkIds;
    }
    return true;
  }
  return false;
}

template <int G, class Lhs>
__device__ inline int32_t search_one_fast(const Group<G>& g, const SearchParams& p, const Lhs& lhs, const FastArena& a, uint2* sm_bucket,
                                          uint32_t* out_path_len, uint64_t* out_pool_off, double* out_final_w,
                                          uint32_t* out_n_tuples, unsigned long long* out_relax) {
  FastState st; st.n_tuples = 0; st.overflow = false; st.relax_calls = 0;
  FastQueue q; q.bucket = sm_bucket; q.occupied = 0; q.top = 0; q.chunk_next = 0; q.free_head = kNoChunk; q.last = 0;
  *out_path_len = 0; *out_pool_off = 0; *out_final_w = d_inf(); *out_n_tuples = 0; *out_relax = 0;
  const DevFstView& F = p.fst;
  if (lhs.start() == kNone || F.start == kNone) return kStNoPath;

  // initial tuple: id 0, dist One, ready at level 0
  if (g.lane == 0) {
    unsigned long long k0 = pack_key(lhs.start(), F.start, 0);
    uint32_t pos; fast_probe(a, k0, pos);
    TupleSlot s; s.key = k0; s.dist = 0.0; s.id_flags = 0; s.prev_id = kNone; s.rhs_arc = kNone; s.lhs_arc = kNone;
    a.table[pos] = s;
    a.key_of[0] = k0;
  }
  g.sync();
  st.n_tuples = 1;
  ready_insert(g, a, q, g.lane == 0, 0u);

  bool have_best = false; uint32_t best_id = 0; double best_fw = d_inf(), best_total = d_inf();
  double cur_dist = 0.0;

  for (;;) {
    if (st.overflow) break;
    if (q.top == 0) {
      bool stop_early;
      if (!advance_level(g, p, a, q, st, have_best, best_total, stop_early)) break;
      cur_dist = __longlong_as_double((long long)q.last);
      continue;
    }
    const uint32_t cur_id = ready_pop(g, a, q);
    const unsigned long long ckey = a.key_of[cur_id];
    const uint32_t s1 = (uint32_t)(ckey >> 34), s2 = (uint32_t)(ckey >> 2), filt = (uint32_t)(ckey & 3u);

    double fw1 = lhs.final_w(s1);
    if (!d_isinf(fw1)) {
      double fw2 = F.final_w[s2];
      if (!d_isinf(fw2)) {
        double final_w = d_times(fw1, fw2);
        double total = d_times(cur_dist, final_w);
        if (!have_best || total < best_total || (total == best_total && cur_id < best_id)) {
          have_best = true; best_id = cur_id; best_fw = final_w; best_total = total;
        }
      }
    }

    const uint4 rec = __ldg(&F.state_rec[s2]);
    const uint32_t n1 = lhs.n_arcs(s1);
    const uint32_t a1base = lhs.arc_base(s1);

    for (uint32_t k = 0; k < n1 && !st.overflow; k++) {
      uint32_t il1, ol1, nx1; double w1;
      lhs.arc(a1base + k, il1, ol1, w1, nx1);
      if (ol1 == 0) continue;
      uint32_t lo, hi;
      equal_range(g, F.ilabel, rec.x, rec.z, ol1, lo, hi);
      for (uint32_t cb = lo; cb < hi && !st.overflow; cb += G) {
        bool active = cb + g.lane < hi;
        Cand c; c.key = 0; c.ew = 0; c.il = il1; c.ol = 0; c.lhs_arc = a1base + k; c.rhs_arc = cb + g.lane;
        if (active) {
          uint4 pl = __ldg(&F.payload[c.rhs_arc]);
          c.ol = pl.x;
          c.ew = d_times(w1, __hiloint2double((int)pl.w, (int)pl.z));
          c.key = pack_key(nx1, pl.y, 0);
        }
        relax_chunk_fast<G>(g, p, lhs, a, q, st, cur_id, cur_dist, active, c);
      }
    }
    if (filt != 1) {
      for (uint32_t kb = 0; kb < n1 && !st.overflow; kb += G) {
        uint32_t k = kb + g.lane;
        bool active = false;
        Cand c; c.key = 0; c.ew = 0; c.il = 0; c.ol = 0; c.lhs_arc = a1base + k; c.rhs_arc = kNone;
        if (k < n1) {
          uint32_t il1, ol1, nx1; double w1;
          lhs.arc(a1base + k, il1, ol1, w1, nx1);
          if (ol1 == 0) { active = true; c.il = il1; c.ew = w1; c.key = pack_key(nx1, s2, filt == 0 ? 2u : filt); }
        }
        if (g.any(active)) relax_chunk_fast<G>(g, p, lhs, a, q, st, cur_id, cur_dist, active, c);
      }
    }
    if (filt != 2) {
      for (uint32_t cb = rec.x; cb < rec.y && !st.overflow; cb += G) {
        bool active = cb + g.lane < rec.y;
        Cand c; c.key = 0; c.ew = 0; c.il = 0; c.ol = 0; c.lhs_arc = kNone; c.rhs_arc = cb + g.lane;
        if (active) {
          uint4 pl = __ldg(&F.payload[c.rhs_arc]);
          c.ol = pl.x;
          c.ew = __hiloint2double((int)pl.w, (int)pl.z);
          c.key = pack_key(s1, pl.y, filt == 0 ? 1u : filt);
        }
        relax_chunk_fast<G>(g, p, lhs, a, q, st, cur_id, cur_dist, active, c);
      }
    }
    if (filt == 0 && rec.y > rec.x) {
      for (uint32_t k = 0; k < n1 && !st.overflow; k++) {
        uint32_t il1, ol1, nx1; double w1;
        lhs.arc(a1base + k, il1, ol1, w1, nx1);
        if (ol1 != 0) continue;
        for (uint32_t cb = rec.x; cb < rec.y && !st.overflow; cb += G) {
          bool active = cb + g.lane < rec.y;
          Cand c; c.key = 0; c.ew = 0; c.il = il1; c.ol = 0; c.lhs_arc = a1base + k; c.rhs_arc = cb + g.lane;
          if (active) {
            uint4 pl = __ldg(&F.payload[c.rhs_arc]);
            c.ol = pl.x;
            c.ew = d_times(w1, __hiloint2double((int)pl.w, (int)pl.z));
            c.key = pack_key(nx1, pl.y, 0);
          }
          relax_chunk_fast<G>(g, p, lhs, a, q, st, cur_id, cur_dist, active, c);
        }
      }
    }
  }

  int32_t status = kStPath;
  uint32_t plen = 0;
  unsigned long long poff = 0;
  uint32_t* scratch = a.chunks;   // queue storage is dead now (>= 4 bytes per tuple by construction)
  if (st.overflow) {
    status = kStRetry;
  } else if (!have_best) {
    status = kStNoPath;
  } else {
    if (g.lane == 0) {
      uint32_t cur = best_id;
      while (cur != 0) {
        uint32_t sl; fast_probe(a, a.key_of[cur], sl);
        uint32_t prev = a.table[sl].prev_id;
        if (prev == kNone) { status = kStNoPath; break; }
        if (plen >= st.n_tuples) { status = kStCycle; break; }
        scratch[plen++] = sl;
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
      for (uint32_t i = g.lane; i < plen; i += G) {
        TupleSlot s = a.table[scratch[i]];
        uint32_t il = 0, ol = 0; double w1 = 0.0, w2 = 0.0, w;
        if (s.lhs_arc != kNone) { uint32_t o, n; lhs.arc(s.lhs_arc, il, o, w1, n); }
        if (s.rhs_arc != kNone) { uint4 pl = __ldg(&F.payload[s.rhs_arc]); ol = pl.x; w2 = __hiloint2double((int)pl.w, (int)pl.z); }
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
  // restore the arena invariants for the next problem: table empty, bitmap zero.
  // Two phases: resolve every tuple's slot first (probing needs intact chains), then clear.
  g.sync();
  uint32_t* slot_tmp = reinterpret_cast<uint32_t*>(a.key_of);
  for (uint32_t base = 0; base < st.n_tuples; base += G) {
    uint32_t i = base + g.lane;
    uint32_t sl = 0;
    if (i < st.n_tuples) fast_probe(a, a.key_of[i], sl);
    g.sync();                                   // all keys of this stripe read before any is overwritten
    if (i < st.n_tuples) slot_tmp[i] = sl;      // slot_tmp[i] aliases key_of[i/2]: stripes before `base` only
    g.sync();
  }
  for (uint32_t i = g.lane; i < st.n_tuples; i += G) a.table[slot_tmp[i]].key = kEmptyKey;
  {
    uint32_t w0 = (st.n_tuples + 63) / 64, w1 = (w0 + 63) / 64, w2 = (w1 + 63) / 64;
    for (uint32_t i = g.lane; i < w0; i += G) a.l0[i] = 0;
    for (uint32_t i = g.lane; i < w1; i += G) a.l1[i] = 0;
    for (uint32_t i = g.lane; i < w2; i += G) a.l2[i] = 0;
  }
  g.sync();
  *out_path_len = plen; *out_pool_off = poff; *out_final_w = (status == kStPath) ? best_fw : d_inf();
  *out_n_tuples = st.n_tuples; *out_relax = st.relax_calls;
  return status;
}

template <int G>
__global__ void __launch_bounds__(128) csp_batch_fast_kernel(SearchParams p) {
  extern __shared__ uint2 sm_buckets[];   // [groups_per_block][64]
  Group<G> g;
  const uint32_t groups_per_block = blockDim.x / G;
  const uint32_t gib = threadIdx.x / G;
  const uint32_t gslot = blockIdx.x * groups_per_block + gib;
  FastArena a = fast_arena_at(p, gslot);
  uint2* my_buckets = sm_buckets + gib * 64;
  unsigned long long relax_total = 0, tuple_total = 0;
  for (;;) {
    uint32_t item = 0;
    if (g.lane == 0) item = atomicAdd(p.queue_head, 1u);
    item = g.shfl(item, 0);
    if (item >= p.n_items) break;
    uint32_t idx = p.order ? p.order[item] : item;
    LhsBytes lhs; lhs.s = p.bytes + p.offsets[idx]; lhs.len = (uint32_t)(p.offsets[idx + 1] - p.offsets[idx]);
    uint32_t plen; uint64_t poff; double fw; uint32_t nt; unsigned long long nr;
    int32_t status = search_one_fast<G>(g, p, lhs, a, my_buckets, &plen, &poff, &fw, &nt, &nr);
    if (g.lane == 0) {
      p.status[idx] = status; p.path_len[idx] = plen; p.pool_off[idx] = poff; p.final_w[idx] = fw; p.n_tuples[idx] = nt;
    }
    relax_total += nr; tuple_total += nt;
  }
  if (g.lane == 0) {
    if (relax_total) atomicAdd(p.relax_counter, relax_total);
    if (tuple_total) atomicAdd(p.tuple_counter, tuple_total);
  }
}

__global__ void __launch_bounds__(32) csp_general_fast_kernel(SearchParams p) {
  __shared__ uint2 sm_b[64];
  Group<32> g;
  FastArena a = fast_arena_at(p, 0);
  LhsCsr lhs; lhs.v = p.lhs;
  uint32_t plen; uint64_t poff; double fw; uint32_t nt; unsigned long long nr;
  int32_t status = search_one_fast<32>(g, p, lhs, a, sm_b, &plen, &poff, &fw, &nt, &nr);
  if (g.lane == 0) {
    p.status[0] = status; p.path_len[0] = plen; p.pool_off[0] = poff; p.final_w[0] = fw; p.n_tuples[0] = nt;
    atomicAdd(p.relax_counter, nr); atomicAdd(p.tuple_counter, (unsigned long long)nt);
  }
}

}  // namespace fstb200
