// Search kernels: lazy compose of a left operand with the frozen transducer and
// n=1 shortest path, one G-lane group per problem, reproducing the reference's
// discovery-order tie-breaking exactly.
//
// What is reproduced (reference src/ops/compose-shortest-path.zig):
//   ids in first-touch order (:70-89), pops in (dist, id) order with lazy
//   deletion (:55-61, :159-163), final pick by (total, id) (:165-179), the four
//   expansion groups in fixed order (:182-365), the relax/tie rule (:91-144),
//   back-track (:372-380) and the result chain (:382-400).
//
// How it is parallelised without changing any observable result:
//   * lanes of a group take the arcs of ONE popped tuple (the reference's inner
//     loops); candidates that hit the same target tuple are folded in lane order
//     in registers, so the memory-visible outcome equals the sequential one;
//   * ids of tuples discovered by one expansion are a ballot prefix in lane (==
//     arc) order;
//   * several pushes of one target inside one expansion collapse into one push of
//     the last taken distance (the earlier entries could only ever be popped as
//     stale, :162).
// Inputs with negative weights make the reference's result depend on stale
// distances (dist[cur] can change during cur's own expansion); those run in
// SERIAL mode: one candidate at a time on lane 0, literally.
#pragma once
#include "device_types.cuh"

namespace fstb200 {

__device__ __forceinline__ bool d_isinf(double x) { return isinf(x); }
// src/weight.zig:19-23
__device__ __forceinline__ double d_times(double a, double b) {
  return (d_isinf(a) || d_isinf(b)) ? __longlong_as_double(0x7FF0000000000000LL) : a + b;
}
__device__ __forceinline__ double d_inf() { return __longlong_as_double(0x7FF0000000000000LL); }

__device__ __forceinline__ unsigned long long pack_key(uint32_t s1, uint32_t s2, uint32_t f) {
  return ((unsigned long long)s1 << 34) | ((unsigned long long)s2 << 2) | f;
}
__device__ __forceinline__ uint32_t hash_key(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return (uint32_t)x;
}

// ── sub-warp group helpers ──
template <int G>
struct Group {
  static_assert(G == 32 || G == 16 || G == 8 || G == 4 || G == 2 || G == 1, "group size");
  unsigned lane;   // lane within the group
  unsigned base;   // first warp lane of the group
  unsigned mask;   // warp-level member mask
  static constexpr unsigned kBits = (G == 32) ? 0xFFFFFFFFu : ((1u << (G & 31)) - 1u);
  __device__ Group() {
    unsigned wl = threadIdx.x & 31u;
    lane = wl % G;
    base = wl - lane;
    mask = (G == 32) ? 0xFFFFFFFFu : (kBits << base);
  }
  __device__ __forceinline__ unsigned ballot(bool p) const { return (__ballot_sync(mask, p) >> base) & kBits; }
  __device__ __forceinline__ bool any(bool p) const { return ballot(p) != 0; }
  template <class T> __device__ __forceinline__ T shfl(T v, int src) const { return __shfl_sync(mask, v, src, G); }
  __device__ __forceinline__ void sync() const { __syncwarp(mask); }
  __device__ __forceinline__ unsigned match_any(unsigned long long v) const {
    return (__match_any_sync(mask, v) >> base) & kBits;
  }
  __device__ __forceinline__ unsigned match_any32(unsigned v) const { return (__match_any_sync(mask, v) >> base) & kBits; }
  __device__ __forceinline__ unsigned lt_mask() const { return (1u << lane) - 1u; }
};

// ── left operand policies ──
// Linear acceptor compiled from bytes: reference src/string.zig:24-50 with
// input == output: state p has one arc (b+1 : b+1 / One -> p+1); only state len is
// final (One).  The empty string is a single final start state (:30-36).
struct LhsBytes {
  const uint8_t* s;
  uint32_t len;
  __device__ uint32_t start() const { return 0; }
  __device__ uint32_t n_arcs(uint32_t s1) const { return s1 < len ? 1u : 0u; }
  __device__ uint32_t arc_base(uint32_t s1) const { return s1; }
  __device__ void arc(uint32_t aid, uint32_t& il, uint32_t& ol, double& w, uint32_t& next) const {
    il = ol = (uint32_t)__ldg(s + aid) + 1u; w = 0.0; next = aid + 1u;
  }
  __device__ double final_w(uint32_t s1) const { return s1 == len ? 0.0 : d_inf(); }
};
// Arbitrary mutable left operand uploaded as CSR in stored arc order.
struct LhsCsr {
  DevLhsCsr v;
  __device__ uint32_t start() const { return v.start; }
  __device__ uint32_t n_arcs(uint32_t s1) const { return v.arc_off[s1 + 1] - v.arc_off[s1]; }
  __device__ uint32_t arc_base(uint32_t s1) const { return v.arc_off[s1]; }
  __device__ void arc(uint32_t aid, uint32_t& il, uint32_t& ol, double& w, uint32_t& next) const {
    il = v.ilabel[aid]; ol = v.olabel[aid]; w = v.weight[aid]; next = v.next[aid];
  }
  __device__ double final_w(uint32_t s1) const { return v.final_w[s1]; }
};

// ── per-group arena view ──
struct Arena {
  TupleSlot* table;     // [hash_cap]
  uint32_t* slot_of;    // [tuple_cap]  id -> slot
  double* heap_dist;    // [heap_cap]
  uint32_t* heap_id;    // [heap_cap]
  uint32_t hash_mask, tuple_cap, heap_cap;
};
__host__ __device__ inline uint64_t arena_bytes(uint32_t hash_cap, uint32_t tuple_cap, uint32_t heap_cap) {
  uint64_t b = (uint64_t)hash_cap * sizeof(TupleSlot);
  b += ((uint64_t)tuple_cap * 4 + 31) & ~31ull;
  b += ((uint64_t)heap_cap * 8 + 31) & ~31ull;
  b += ((uint64_t)heap_cap * 4 + 31) & ~31ull;
  return (b + 255) & ~255ull;
}
__device__ inline Arena arena_at(const SearchParams& p, uint32_t slot_idx) {
  Arena a;
  uint8_t* base = p.arena + (uint64_t)slot_idx * p.arena_stride;
  a.table = reinterpret_cast<TupleSlot*>(base);
  base += (uint64_t)p.hash_cap * sizeof(TupleSlot);
  a.slot_of = reinterpret_cast<uint32_t*>(base);
  base += ((uint64_t)p.tuple_cap * 4 + 31) & ~31ull;
  a.heap_dist = reinterpret_cast<double*>(base);
  base += ((uint64_t)p.heap_cap * 8 + 31) & ~31ull;
  a.heap_id = reinterpret_cast<uint32_t*>(base);
  a.hash_mask = p.hash_cap - 1; a.tuple_cap = p.tuple_cap; a.heap_cap = p.heap_cap;
  return a;
}

// ── binary min-heap on (dist, id), operated by one lane (reference :55-61) ──
__device__ __forceinline__ bool key_less(double ad, uint32_t ai, double bd, uint32_t bi) {
  return ad < bd || (ad == bd && ai < bi);
}
__device__ inline void heap_push(const Arena& a, uint32_t& size, double d, uint32_t id) {
  uint32_t i = size++;
  while (i > 0) {
    uint32_t par = (i - 1) >> 1;
    double pd = a.heap_dist[par]; uint32_t pi = a.heap_id[par];
    if (!key_less(d, id, pd, pi)) break;
    a.heap_dist[i] = pd; a.heap_id[i] = pi;
    i = par;
  }
  a.heap_dist[i] = d; a.heap_id[i] = id;
}
__device__ inline void heap_pop(const Arena& a, uint32_t& size, double& d, uint32_t& id) {
  d = a.heap_dist[0]; id = a.heap_id[0];
  uint32_t n = --size;
  if (n == 0) return;
  double ld = a.heap_dist[n]; uint32_t li = a.heap_id[n];
  uint32_t i = 0;
  for (;;) {
    uint32_t c = 2 * i + 1;
    if (c >= n) break;
    double cd = a.heap_dist[c]; uint32_t ci = a.heap_id[c];
    if (c + 1 < n) {
      double rd = a.heap_dist[c + 1]; uint32_t ri = a.heap_id[c + 1];
      if (key_less(rd, ri, cd, ci)) { c++; cd = rd; ci = ri; }
    }
    if (!key_less(cd, ci, ld, li)) break;
    a.heap_dist[i] = cd; a.heap_id[i] = ci;
    i = c;
  }
  a.heap_dist[i] = ld; a.heap_id[i] = li;
}

// Equal range of `x` in ilabel[begin, end) (reference src/fst.zig:112-136),
// group-cooperative: one ballot when the state has <= G arcs, G-ary narrowing else.
template <int G>
__device__ inline void equal_range(const Group<G>& g, const uint32_t* __restrict__ il, uint32_t begin, uint32_t end,
                                   uint32_t x, uint32_t& lo_out, uint32_t& hi_out) {
  // lower bound
  uint32_t lo = begin, hi = end;
  while (hi - lo > (uint32_t)G) {
    uint32_t len = hi - lo;
    uint32_t idx = lo + (uint32_t)(((unsigned long long)(g.lane + 1) * len) / (G + 1));
    uint32_t cnt = __popc(g.ballot(__ldg(il + idx) < x));
    uint32_t nlo = cnt == 0 ? lo : lo + (uint32_t)(((unsigned long long)cnt * len) / (G + 1)) + 1;
    uint32_t nhi = cnt == (uint32_t)G ? hi : lo + (uint32_t)(((unsigned long long)(cnt + 1) * len) / (G + 1));
    lo = nlo; hi = nhi;
  }
  {
    bool valid = lo + g.lane < hi;
    uint32_t v = valid ? __ldg(il + lo + g.lane) : 0u;
    lo += __popc(g.ballot(valid && v < x));
  }
  lo_out = lo;
  // upper bound, starting from lo like the reference
  hi = end;
  uint32_t l2 = lo;
  while (hi - l2 > (uint32_t)G) {
    uint32_t len = hi - l2;
    uint32_t idx = l2 + (uint32_t)(((unsigned long long)(g.lane + 1) * len) / (G + 1));
    uint32_t cnt = __popc(g.ballot(__ldg(il + idx) <= x));
    uint32_t nlo = cnt == 0 ? l2 : l2 + (uint32_t)(((unsigned long long)cnt * len) / (G + 1)) + 1;
    uint32_t nhi = cnt == (uint32_t)G ? hi : l2 + (uint32_t)(((unsigned long long)(cnt + 1) * len) / (G + 1));
    l2 = nlo; hi = nhi;
  }
  {
    bool valid = l2 + g.lane < hi;
    uint32_t v = valid ? __ldg(il + l2 + g.lane) : 0u;
    l2 += __popc(g.ballot(valid && v <= x));
  }
  hi_out = l2;
}

// One relaxation candidate held by a lane.
struct Cand {
  unsigned long long key;
  double ew;         // edge weight (reference: the value passed to relax, :105)
  uint32_t il, ol;   // labels of the composed arc
  uint32_t lhs_arc, rhs_arc;
};

// Search state of one group (uniform across the lanes of the group).
struct SearchState {
  uint32_t n_tuples;
  uint32_t heap_size;
  bool overflow;
  unsigned long long relax_calls;
};

// Labels of the back-pointer stored in a slot (only needed on the rare tie where
// the stored predecessor is the tuple being expanded).
template <class Lhs>
__device__ inline void backptr_labels(const SearchParams& p, const Lhs& lhs, const TupleSlot& s, uint32_t& il, uint32_t& ol) {
  il = 0; ol = 0;
  if (s.lhs_arc != kNone) { uint32_t a, b, n; double w; lhs.arc(s.lhs_arc, a, b, w, n); il = a; }
  if (s.rhs_arc != kNone) ol = __ldg(&p.fst.payload[s.rhs_arc]).x;
}

// reference :112-126 — decide whether candidate (nd, cur, il, ol) replaces state S.
__device__ __forceinline__ bool take_rule(double nd, uint32_t cur, uint32_t il, uint32_t ol, double s_dist,
                                          uint32_t s_prev, uint32_t s_il, uint32_t s_ol) {
  if (d_isinf(s_dist) || nd < s_dist) return true;
  if (nd == s_dist) {
    if (s_prev == kNone) return true;
    return cur < s_prev || (cur == s_prev && (il < s_il || (il == s_il && ol < s_ol)));
  }
  return false;
}

// Find `key` or the empty slot where it would go (read-only).
__device__ __forceinline__ bool probe(const Arena& a, unsigned long long key, uint32_t& pos) {
  uint32_t i = hash_key(key) & a.hash_mask;
  for (;;) {
    unsigned long long k = a.table[i].key;
    if (k == key) { pos = i; return true; }
    if (k == kEmptyKey) { pos = i; return false; }
    i = (i + 1) & a.hash_mask;
  }
}
// Claim a slot for `key` starting at `pos` (other lanes of the group may be
// inserting different keys concurrently).
__device__ __forceinline__ uint32_t claim(const Arena& a, unsigned long long key, uint32_t pos) {
  for (;;) {
    unsigned long long old = atomicCAS(&a.table[pos].key, kEmptyKey, key);
    if (old == kEmptyKey) return pos;
    pos = (pos + 1) & a.hash_mask;
  }
}

// Sequential relax of one candidate by a single lane (SERIAL mode and the
// initial tuple).  Literal restatement of reference :91-144.
template <class Lhs>
__device__ inline void relax_one(const SearchParams& p, const Lhs& lhs, const Arena& a, SearchState& st, uint32_t cur_id,
                                 uint32_t cur_slot, const Cand& c) {
  st.relax_calls++;
  uint32_t pos;
  bool found = probe(a, c.key, pos);
  if (!found) {
    if (st.n_tuples >= a.tuple_cap) { st.overflow = true; return; }
    a.table[pos].key = c.key;
    TupleSlot ns; ns.key = c.key; ns.dist = d_inf(); ns.id_flags = st.n_tuples; ns.prev_id = kNone; ns.rhs_arc = kNone; ns.lhs_arc = kNone;
    a.table[pos] = ns;
    a.slot_of[st.n_tuples] = pos;
    st.n_tuples++;
  }
  TupleSlot s = a.table[pos];
  double nd = d_times(a.table[cur_slot].dist, c.ew);   // dist[cur] re-read every time (:108)
  uint32_t s_il = 0, s_ol = 0;
  if (s.prev_id == cur_id) backptr_labels(p, lhs, s, s_il, s_ol);
  if (!take_rule(nd, cur_id, c.il, c.ol, s.dist, s.prev_id, s_il, s_ol)) return;
  s.dist = nd; s.prev_id = cur_id; s.rhs_arc = c.rhs_arc; s.lhs_arc = c.lhs_arc;
  a.table[pos] = s;
  if (!(s.id_flags & kSettledBit)) {
    if (st.heap_size >= a.heap_cap) { st.overflow = true; return; }
    heap_push(a, st.heap_size, nd, s.id_flags);
  }
}

// Relax up to G candidates of ONE expansion in parallel; equivalent to relaxing
// them sequentially in lane order (see file header).
template <int G, bool SERIAL, class Lhs>
__device__ inline void relax_chunk(const Group<G>& g, const SearchParams& p, const Lhs& lhs, const Arena& a, SearchState& st,
                                   uint32_t cur_id, uint32_t cur_slot, double cur_dist, bool active, const Cand& c) {
  unsigned act = g.ballot(active);
  if (act == 0) return;
  if (SERIAL) {
    // literal: one candidate at a time on lane 0
    unsigned m = act;
    while (m) {
      int src = __ffs(m) - 1; m &= m - 1;
      Cand b;
      b.key = g.shfl(c.key, src); b.ew = g.shfl(c.ew, src); b.il = g.shfl(c.il, src); b.ol = g.shfl(c.ol, src);
      b.lhs_arc = g.shfl(c.lhs_arc, src); b.rhs_arc = g.shfl(c.rhs_arc, src);
      if (g.lane == 0 && !st.overflow) relax_one(p, lhs, a, st, cur_id, cur_slot, b);
      g.sync();
      st.n_tuples = g.shfl(st.n_tuples, 0);
      st.heap_size = g.shfl(st.heap_size, 0);
      st.overflow = g.shfl((int)st.overflow, 0) != 0;
      st.relax_calls = g.shfl(st.relax_calls, 0);
    }
    return;
  }
  st.relax_calls += __popc(act);
  // group candidates by target tuple; the first lane of each group leads
  unsigned long long mkey = active ? c.key : (0xFFFFFFFFFFFFFF00ull | g.lane);
  unsigned peers = g.match_any(mkey);
  bool leader = active && ((unsigned)(__ffs(peers) - 1) == g.lane);
  uint32_t pos = 0; bool found = false;
  if (leader) found = probe(a, c.key, pos);
  unsigned newmask = g.ballot(leader && !found);
  uint32_t n_new = __popc(newmask);
  if (st.n_tuples + n_new > a.tuple_cap) { st.overflow = true; return; }
  TupleSlot s;
  if (leader) {
    if (!found) {
      pos = claim(a, c.key, pos);
      uint32_t my_id = st.n_tuples + __popc(newmask & g.lt_mask());   // discovery order == lane (arc) order
      a.slot_of[my_id] = pos;
      s.key = c.key; s.dist = d_inf(); s.id_flags = my_id; s.prev_id = kNone; s.rhs_arc = kNone; s.lhs_arc = kNone;
    } else {
      s = a.table[pos];
    }
  }
  st.n_tuples += n_new;
  // fold the candidates of each target in lane order (registers only)
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
  // one push per changed, unsettled target (heap is operated by lane 0)
  unsigned pushm = g.ballot(leader && changed && !(s.id_flags & kSettledBit));
  if (st.heap_size + __popc(pushm) > a.heap_cap) { st.overflow = true; return; }
  while (pushm) {
    int src = __ffs(pushm) - 1; pushm &= pushm - 1;
    double pd = g.shfl(s.dist, src);
    uint32_t pid = g.shfl(s.id_flags, src) & ~kSettledBit;
    if (g.lane == 0) heap_push(a, st.heap_size, pd, pid);
  }
  g.sync();
  st.heap_size = g.shfl(st.heap_size, 0);
}

// Search one problem.  Returns the status; on kStPath the reversed path is in the
// pool at [*pool_off, *pool_off + *path_len).
template <int G, bool SERIAL, class Lhs>
__device__ inline int32_t search_one(const Group<G>& g, const SearchParams& p, const Lhs& lhs, const Arena& a,
                                     uint32_t* out_path_len, uint64_t* out_pool_off, double* out_final_w,
                                     uint32_t* out_n_tuples, unsigned long long* out_relax) {
  SearchState st; st.n_tuples = 0; st.heap_size = 0; st.overflow = false; st.relax_calls = 0;
  *out_path_len = 0; *out_pool_off = 0; *out_final_w = d_inf(); *out_n_tuples = 0; *out_relax = 0;
  const DevFstView& F = p.fst;
  // reference :30-32 (n == 1 always here; n == 0 / n > 1 are handled on the host)
  if (lhs.start() == kNone || F.start == kNone) return kStNoPath;

  // :146-153 initial tuple, id 0, dist One
  if (g.lane == 0) {
    unsigned long long k0 = pack_key(lhs.start(), F.start, 0);
    uint32_t pos; probe(a, k0, pos);
    TupleSlot s; s.key = k0; s.dist = 0.0; s.id_flags = 0; s.prev_id = kNone; s.rhs_arc = kNone; s.lhs_arc = kNone;
    a.table[pos] = s;
    a.slot_of[0] = pos;
    heap_push(a, st.heap_size, 0.0, 0);
  }
  g.sync();
  st.n_tuples = 1; st.heap_size = 1;

  bool have_best = false; uint32_t best_id = 0; double best_fw = d_inf(), best_total = d_inf();

  while (st.heap_size > 0 && !st.overflow) {   // :159
    double pd = 0; uint32_t cur_id = 0;
    if (g.lane == 0) heap_pop(a, st.heap_size, pd, cur_id);
    g.sync();
    st.heap_size = g.shfl(st.heap_size, 0);
    pd = g.shfl(pd, 0); cur_id = g.shfl(cur_id, 0);
    uint32_t cur_slot = a.slot_of[cur_id];
    TupleSlot cs = a.table[cur_slot];
    if (cs.id_flags & kSettledBit) continue;   // :161
    if (pd != cs.dist) continue;                // :162 stale
    if (g.lane == 0) a.table[cur_slot].id_flags = cs.id_flags | kSettledBit;   // :163
    g.sync();
    const uint32_t s1 = (uint32_t)(cs.key >> 34), s2 = (uint32_t)(cs.key >> 2), filt = (uint32_t)(cs.key & 3u);
    const double cur_dist = cs.dist;

    // early exit that cannot change the result: every
    // remaining pop has dist >= cur_dist; if even a zero final weight cannot reach
    // best_total, neither the final pick nor any back-pointer on the best path can
    // change.  Disabled in exhaustive mode and in SERIAL (negative-weight) mode.
    if (!SERIAL && !p.exhaustive && have_best && cur_dist > best_total) break;

    // :165-179 final check
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

    const uint4 rec = __ldg(&F.state_rec[s2]);   // {arc_begin, eps_end, arc_end}
    const uint32_t n1 = lhs.n_arcs(s1);
    const uint32_t a1base = lhs.arc_base(s1);

    // :182-202 non-epsilon matches
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
        relax_chunk<G, SERIAL>(g, p, lhs, a, st, cur_id, cur_slot, cur_dist, active, c);
      }
    }
    // :227-252 left operand consumes an output-epsilon arc
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
        relax_chunk<G, SERIAL>(g, p, lhs, a, st, cur_id, cur_slot, cur_dist, active, c);
      }
    }
    // :254-278 transducer consumes an input-epsilon arc
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
        relax_chunk<G, SERIAL>(g, p, lhs, a, st, cur_id, cur_slot, cur_dist, active, c);
      }
    }
    // :307-336 both consume epsilon
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
          relax_chunk<G, SERIAL>(g, p, lhs, a, st, cur_id, cur_slot, cur_dist, active, c);
        }
      }
    }
  }

  int32_t status = kStPath;
  uint32_t plen = 0;
  unsigned long long poff = 0;
  uint32_t* scratch = reinterpret_cast<uint32_t*>(a.heap_dist);   // heap is dead now; >= 2*heap_cap u32
  if (st.overflow) {
    status = kStRetry;
  } else if (!have_best) {
    status = kStNoPath;                                            // :368-370
  } else {
    // :372-380 back-track on lane 0 (pointer chase), recording slots
    if (g.lane == 0) {
      uint32_t cur = best_id;
      while (cur != 0) {
        uint32_t sl = a.slot_of[cur];
        uint32_t prev = a.table[sl].prev_id;
        if (prev == kNone) { status = kStNoPath; break; }          // :375-377
        if (plen >= st.n_tuples) { status = kStCycle; break; }     // hazard H1 (reference: OOM)
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
      // emit the reversed chain: arc = (ilabel of lhs arc | eps, olabel of rhs arc | eps, edge weight)
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
  // leave the table empty for the next problem of this group (cost ~ N, not table size)
  g.sync();
  for (uint32_t i = g.lane; i < st.n_tuples; i += G) a.table[a.slot_of[i]].key = kEmptyKey;
  g.sync();
  *out_path_len = plen; *out_pool_off = poff; *out_final_w = (status == kStPath) ? best_fw : d_inf();
  *out_n_tuples = st.n_tuples; *out_relax = st.relax_calls;
  return status;
}

// Persistent batch kernel: each G-lane group pulls strings from a global queue.
template <int G, bool SERIAL>
__global__ void __launch_bounds__(128) csp_batch_kernel(SearchParams p) {
  Group<G> g;
  const uint32_t groups_per_block = blockDim.x / G;
  const uint32_t gslot = blockIdx.x * groups_per_block + threadIdx.x / G;
  Arena a = arena_at(p, gslot);
  unsigned long long relax_total = 0, tuple_total = 0;
  for (;;) {
    uint32_t item = 0;
    if (g.lane == 0) item = atomicAdd(p.queue_head, 1u);
    item = g.shfl(item, 0);
    if (item >= p.n_items) break;
    uint32_t idx = p.order ? p.order[item] : item;
    LhsBytes lhs; lhs.s = p.bytes + p.offsets[idx]; lhs.len = (uint32_t)(p.offsets[idx + 1] - p.offsets[idx]);
    uint32_t plen = 0; uint64_t poff = 0; double fw = d_inf(); uint32_t nt = 0; unsigned long long nr = 0;
    const int32_t pre = p.skip ? p.skip[idx] : kStPath;   // an earlier pipeline stage failed: pass its status through
    int32_t status = pre != kStPath ? pre : search_one<G, SERIAL>(g, p, lhs, a, &plen, &poff, &fw, &nt, &nr);
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

// One general left operand (the fst_compose_frozen_shortest_path drop-in).
template <bool SERIAL>
__global__ void __launch_bounds__(32) csp_general_kernel(SearchParams p) {
  Group<32> g;
  Arena a = arena_at(p, 0);
  LhsCsr lhs; lhs.v = p.lhs;
  uint32_t plen; uint64_t poff; double fw; uint32_t nt; unsigned long long nr;
  int32_t status = search_one<32, SERIAL>(g, p, lhs, a, &plen, &poff, &fw, &nt, &nr);
  if (g.lane == 0) {
    p.status[0] = status; p.path_len[0] = plen; p.pool_off[0] = poff; p.final_w[0] = fw; p.n_tuples[0] = nt;
    atomicAdd(p.relax_counter, nr); atomicAdd(p.tuple_counter, (unsigned long long)nt);
  }
}

// Emit kernel: un-reverse each string's pool segment into the ordered flat output
// and (optionally) write the output-tape bytes (reference src/string.zig:64-97).
__global__ void csp_emit_kernel(EmitParams e) {
  const uint32_t warps_per_block = blockDim.x >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t i = blockIdx.x * warps_per_block + (threadIdx.x >> 5); i < e.n_strings; i += gridDim.x * warps_per_block) {
    if (e.status[i] != kStPath && e.status[i] != kStNotBytes) continue;
    const bool bytes_ok = e.status[i] == kStPath;
    uint32_t n = e.path_len[i];
    uint64_t src = e.pool_off[i], dst = e.path_offsets[i];
    if (dst + n > e.path_capacity) continue;
    uint64_t ob = e.out_bytes ? e.out_offsets[i] : 0;
    for (uint32_t base = 0; base < n; base += 32) {
      uint32_t k = base + lane;
      bool valid = k < n;
      PoolArc a; a.ilabel = 0; a.olabel = 0; a.weight = 0.0;
      if (valid) {
        a = e.pool[src + (n - 1 - k)];
        e.ilabels[dst + k] = a.ilabel; e.olabels[dst + k] = a.olabel; e.weights[dst + k] = a.weight;
      }
      if (e.out_bytes && bytes_ok) {
        unsigned m = __ballot_sync(0xFFFFFFFFu, valid && a.olabel != 0);
        if (valid && a.olabel != 0) e.out_bytes[ob + __popc(m & ((1u << lane) - 1u))] = (uint8_t)(a.olabel - 1u);
        ob += __popc(m);
      }
    }
  }
}

// Output-tape byte count per string (for the exclusive scan that places out_bytes).
// A path with an output label above 256 has no byte form: its status becomes kStNotBytes, its output string is empty.
__global__ void csp_count_out_kernel(int32_t* status, const uint32_t* path_len, const uint64_t* pool_off,
                                     const PoolArc* pool, uint32_t n_strings, uint32_t* out_len) {
  const uint32_t warps_per_block = blockDim.x >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t i = blockIdx.x * warps_per_block + (threadIdx.x >> 5); i < n_strings; i += gridDim.x * warps_per_block) {
    uint32_t cnt = 0;
    bool wide = false;
    if (status[i] == kStPath) {
      uint32_t n = path_len[i]; uint64_t src = pool_off[i];
      for (uint32_t k = lane; k < n; k += 32) { const uint32_t ol = pool[src + k].olabel; cnt += ol != 0; wide |= ol > 256u; }
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    wide = __any_sync(0xFFFFFFFFu, wide);
    if (lane == 0) { out_len[i] = wide ? 0u : cnt; if (wide) status[i] = kStNotBytes; }
  }
}

}  // namespace fstb200
