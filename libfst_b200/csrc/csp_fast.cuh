// FAST lean kernel: the hot configuration of the batched path with everything that is not needed per pop moved
// out of the pop loop.  Same search, same arena, same rare paths as csp_lean.cuh (the window switch, the level
// advance, the back-track and the arena clean-up are the functions of that file); what is new is the STEP:
//
//   configuration   8 lanes per string reading the INTEGER leader slab (`islab`), dense table of compact 8-byte
//                   records (every weight of the transducer is an integer in 0..4095); lazy semantics
//                   (compose-shortest-path.zig:26-401) and, as a second instantiation, the eager pair
//                   (compose.zig:29-198 + shortest-path.zig:18-139).  Everything else stays with csp_batch_lean_kernel.
//   integer relax   the compact record is  dist:20 | prev:22 | id:22  (most significant first).  With
//                   cand = (new dist, popped id) in the same position, the reference's rule (:109-126)
//                       take  <=>  new < old  ||  (new == old && (no back-pointer || popped id < prev))
//                   is ONE unsigned compare  cand < record >> 22  (an untouched record is all ones: always
//                   taken; "no back-pointer" is the largest prev), "strictly lowered" (:109-114, :137-142) is
//                   (cand | low 12 bits) < high word, and the new record is cand with the tuple's id below it.
//                   The slab carries the weight pre-shifted to the distance field: no f64 in the loop.
//   service split   the POP LOOP (fast_pop_loop) is an out-of-line leaf function of straight-line code that every
//                   lane of the warp executes: warp-wide votes, no calls, its own register allocation.  It
//                   returns when a group needs SERVICE (next window line, next level, evicting the window below
//                   a new smaller id, the end of the string, the next string: fast_service) or when a step has a
//                   RARE TAIL (final check at the end of the string, an id below the window, a push into the
//                   radix heap once it exists, a state wider than the slab: fast_rare); the kernel runs that for
//                   the groups that asked and re-enters the loop.  The hot values travel through a struct in local
//                   memory (FastCold), the launch parameters through __constant__ memory (a leaf reads
//                   constant-bank operands for free).
//   idle groups     a group without a string reads the dummy state row S of the slab (no record matches), so
//                   the hot path needs no "is my group alive" branches.
//
// The reference's observable behaviour is untouched: ids in first-touch order (:70-89) = ballot prefix in lane
// order (lane order of the leader slab = expansion order, match arcs :182-202 before input-epsilon arcs :254-278),
// pops in (dist, id) order (:55-61, :159-163) = lowest set bit of the ready set at the current level.
#pragma once
#include "csp_lean.cuh"

namespace fstb200 {

#ifndef FAST_POP_FN
#define FAST_POP_FN __noinline__
#endif
#ifndef FAST_COLD_FN
#define FAST_COLD_FN __forceinline__
#endif
constexpr uint32_t kFastNoLabel = 0xFFFFFFFDu;
enum : uint32_t { kFastFetch = 0, kFastRun = 1, kFastFinish = 2, kFastDone = 3, kFastBfs = 4 };

// Everything the pop loop does not touch, in LOCAL memory: the out-of-line service code works on it, the pop loop
// keeps its handful of hot values in registers and syncs them only around a service call.
struct FastCold {
  LeanState st;
  LhsBytes lhs;
  uint32_t phase, idx, steps, lev12, fmin12, rc;
  unsigned long long relax_total, tuple_total;
  // event of the step that left the pop loop (fast_pop_loop -> kernel)
  uint32_t ev_rare, ev_flags, ev_id, ev_nd, ev_bigno, ev_cur, ev_P, ev_s2, ev_x, ev_own;
};

__device__ __forceinline__ LeanCtx fast_ctx(const SearchParams& p) {
  extern __shared__ __align__(16) uint32_t smem_all[];
  const uint32_t gib = threadIdx.x / 8u;
  LeanCtx c;
  c.base = p.arena + (uint64_t)(blockIdx.x * (blockDim.x / 8u) + gib) * p.arena_stride;
  c.sm = smem_all + (size_t)gib * p.smem_words;
  return c;
}

// Final check of a popped tuple at the end of the string; group-local.  Search phase (compose-shortest-path.zig
// :165-179): best = smallest (total, id).  BFS phase of the eager semantics (shortest-path.zig:88-104): the pops come in
// increasing state number, so only a strictly smaller total replaces the best; `dist` is the state's own distance.
__device__ FAST_COLD_FN void fast_final_check(const SearchParams& p, const FastCold& f, uint32_t s2, uint32_t cur_id, bool bfs, double dist) {
  constexpr int G = 8;
  const double fw2 = p.fst.final_w[s2];
  if (d_isinf(fw2) || (bfs && d_isinf(dist))) return;
  const Group<G> g;
  const LeanCtx c = fast_ctx(p);
  uint32_t* cold = LEAN_COLD(p, c, G);
  const double final_w = 0.0 + fw2;
  const double total = (bfs ? dist : __longlong_as_double((long long)f.st.last)) + final_w;
  const bool have_best = cold[kcHaveBest] != 0;
  const double best_total = __hiloint2double((int)cold[kcBestTotHi], (int)cold[kcBestTotLo]);
  const uint32_t best_id = cold[kcBestId];
  g.sync();
  if (!have_best || total < best_total || (!bfs && total == best_total && cur_id < best_id)) {
    if (g.lane == 0) {
      cold[kcHaveBest] = 1; cold[kcBestId] = cur_id;
      cold[kcBestFwLo] = (uint32_t)__double2loint(final_w); cold[kcBestFwHi] = (uint32_t)__double2hiint(final_w);
      cold[kcBestTotLo] = (uint32_t)__double2loint(total); cold[kcBestTotHi] = (uint32_t)__double2hiint(total);
    }
  }
  g.sync();
}

// Rare tail of a hot step, group-local and out of the pop loop: the final check of a popped tuple at the end of the
// string, a ready id below the window, future-set pushes once the radix heap exists (a second distance level was
// needed), and the expansion of a state wider than the leader slab (marker record; nothing was relaxed in the step):
// the arcs that can match come from the state's label index when they fit the group, else from the generic loops of
// lean_step (binary-searched match range, then the epsilon prefix, G arcs per relax step).
template <bool EAGER>
__device__ FAST_COLD_FN void fast_rare(const SearchParams& p, FastCold& f, bool fin, bool low, bool fut, uint32_t id, uint32_t nd, bool big,
                                       uint32_t bigno, uint32_t cur_id, uint32_t s1, uint32_t s2, uint32_t x, uint32_t own_d) {
  constexpr int G = 8, DENSE = EAGER ? 3 : 2;
  const bool bfs = EAGER && f.phase == kFastBfs;
  const double bfs_dist = own_d == 0xFFFFFu ? d_inf() : (double)own_d;   // BFS phase: the popped state's own distance
  const DevFstView& F = p.fst;
  const Group<G> g;
  const LeanCtx c = fast_ctx(p);
  LeanState& st = f.st;
  if (fin) fast_final_check(p, f, s2, cur_id, bfs, bfs_dist);
  if (st.n_tuples > p.tuple_cap) st.overflow = true;
  if (g.any(low)) st.low_pending = true;
  if (st.sorted && g.any(fut)) {
    const unsigned long long k = (unsigned long long)__double_as_longlong((double)nd);
    lean_bucket_push<G>(p, g, c, st, fut, id, fut ? bucket_of(k, st.last) : 1u, k);
  }
  if (!big) return;
  if (F.bigidx) {
    const uint2* bi = F.bigidx + (uint64_t)bigno * 257u;
    const uint2 e = __ldg(bi);
    const uint2 m = x <= 256u ? __ldg(bi + x) : make_uint2(0u, 0u);
    if (m.y + e.y <= (uint32_t)G) {
      uint4 r = make_uint4(0xFFFFFFFFu, 0x80000000u, 0u, 0u);
      const bool is_match = g.lane < m.y, is_eps = !is_match && g.lane < m.y + e.y;
      if (is_match) r = __ldg(&F.sarc[m.x + g.lane]);
      else if (is_eps) r = __ldg(&F.sarc[e.x + (g.lane - m.y)]);
      const bool hit = is_match || is_eps;
      lean_relax<G, DENSE, false, EAGER>(p, g, c, st, cur_id, hit ? 1u : 0u, hit && !(r.y >> 31), is_match ? s1 + 1u : s1, (r.y << 1) | (is_match ? 0u : 1u),
                                     __hiloint2double((int)r.w, (int)r.z), Group<G>::kBits, bfs, bfs_dist);
      g.sync();
      return;
    }
  }
  const uint4 rec = __ldg(&F.state_rec[s2]);
  uint32_t lo = 0, hi = 0;
  if (s1 < f.lhs.len) equal_range(g, F.ilabel, rec.x, rec.z, x, lo, hi);
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

// One service call of a group: whatever its string needs that is not a pop (see the file header).
template <bool EAGER>
__device__ FAST_COLD_FN void fast_service(const SearchParams& p, FastCold& f) {
  constexpr int G = 8, DENSE = EAGER ? 3 : 2;
  const DevFstView& F = p.fst;
  const Group<G> g;
  const LeanCtx c = fast_ctx(p);
  LeanState& st = f.st;
  uint32_t* const win = LEAN_WIN(c);
  if (f.phase == kFastRun) {
    if (f.fmin12 != ~0u) {
      const unsigned long long k = (unsigned long long)__double_as_longlong((double)(f.fmin12 >> 12));
      if (k < st.future_min) st.future_min = k;
      f.fmin12 = ~0u;
    }
    if (st.n_tuples > p.tuple_cap) st.overflow = true;
    if (++f.steps > 8u * p.tuple_cap + 4096u) st.stuck = true;   // safety valve: a service call retires a window line, a level or a string
    if (st.overflow || st.heap_overflow || st.stuck) {
      f.phase = kFastFinish;
    } else {
      if (st.low_pending) { lean_window_evict<G>(p, g, c, st); st.low_pending = false; }
      if (!g.any(win[g.lane] != 0)) {
        if (!lean_window_next<G, DENSE>(p, g, c, st)) {
          if (!lean_advance_level<G, DENSE>(p, g, c, st)) {
            if (EAGER && !st.heap_overflow) {
              // eager semantics: the distances are known; number the lattice in FIFO order (compose.zig:29-198) and give
              // every state its first tight relaxer as back-pointer (shortest-path.zig:75-78) — the BFS steps of the pop loop
              lean_bfs_begin<G, DENSE>(p, g, c, st);
              f.rc = 0;
              f.phase = kFastBfs;
            } else {
              f.phase = kFastFinish;
            }
          } else {
            const double lv = __longlong_as_double((long long)st.last);
            // the compact record holds distances below 2^20 - 1: a step adds at most 4095
            if (lv > kCrecMaxDist - 4096.0) { st.wide = true; f.phase = kFastFinish; }
            else f.lev12 = __double2uint_rn(lv) << 12;
          }
        }
      }
    }
  }
  else if (EAGER && f.phase == kFastBfs) {
    // the FIFO is drained (the lattice is complete) or the id -> key array is full
    if (st.n_tuples > p.tuple_cap) st.overflow = true;
    f.phase = kFastFinish;
  }
  if (f.phase == kFastFinish) {
    uint32_t plen; uint64_t poff; double fw;
    st.relax_calls += f.rc;
    if (EAGER && p.lat_state_base != nullptr) lean_emit_lattice<G, DENSE>(p, g, c, st, f.lhs, f.idx);
    const int32_t status = lean_finish<G, DENSE>(p, g, c, st, f.lhs, &plen, &poff, &fw);
    if (g.lane == 0) {
      p.status[f.idx] = status; p.path_len[f.idx] = plen; p.pool_off[f.idx] = poff; p.final_w[f.idx] = fw; p.n_tuples[f.idx] = st.n_tuples;
    }
    f.relax_total += st.relax_calls; f.tuple_total += st.n_tuples;
    f.lhs.len = 0;
    f.phase = kFastFetch;
  } else if (f.phase == kFastFetch) {
    uint32_t item = 0;
    if (g.lane == 0) item = atomicAdd(p.queue_head, 1u);
    item = g.shfl(item, 0);
    if (item >= p.n_items) {
      f.phase = kFastDone;
      // an idle group keeps stepping with its warp: it pops slot 0, whose key names the slab's all-idle row S
      if (g.lane == 0) reinterpret_cast<uint32_t*>(LEAN_KEYOF(p, c))[0] = F.num_states << 1;
      g.sync();
    } else {
      const uint32_t idx = p.order ? p.order[item] : item;
      f.idx = idx;
      const int32_t pre = p.skip ? p.skip[idx] : kStPath;
      if (F.start == kNone || pre != kStPath) {
        if (g.lane == 0) { p.status[idx] = pre != kStPath ? pre : kStNoPath; p.path_len[idx] = 0; p.pool_off[idx] = 0; p.final_w[idx] = d_inf(); p.n_tuples[idx] = 0; }
      } else {
        f.lhs.s = p.bytes + p.offsets[idx]; f.lhs.len = (uint32_t)(p.offsets[idx + 1] - p.offsets[idx]);
        lean_begin<G, DENSE>(p, g, c, st);
        f.lev12 = 0; f.fmin12 = ~0u; f.rc = 0; f.steps = 0;
        f.phase = kFastRun;
      }
    }
  }
}

#ifndef FSTB_FAST_MINBLOCKS
#define FSTB_FAST_MINBLOCKS 8
#endif

// Launch parameters of the fast kernel in CONSTANT memory (set with cudaMemcpyToSymbolAsync on the launch stream):
// the pop loop is an out-of-line leaf function, and a leaf reads constant-bank operands for free where a reference to
// the kernel's own parameter block would cost a register or a load per use.
__constant__ SearchParams c_fp;

__device__ __forceinline__ uint32_t fast_lds(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void fast_red_and(uint32_t a, uint32_t v) { asm volatile("red.shared.and.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void fast_red_or(uint32_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// The pop loop: runs steps until a group of the warp needs a service iteration or the rare tail of a step.  Nothing
// in here calls or branches to cold code; the hot values travel through `f`.  Everything that only service code
// changes (window line, string, level, the service flags) is loop-invariant.
// EAGER instantiation: a group is either in the search phase (steps as in the lazy kernel) or in the BFS phase of the
// eager semantics (csp_lean.cuh file header): pops are the lattice states in FIFO order (a cursor), every target
// without a BFS number gets the next one, and a target takes the popped state as back-pointer if it has none yet and
// the arc is tight (shortest-path.zig:75-78: the first tight relaxer is the smallest-numbered one).  Both kinds of
// step share the loads, the votes and the id -> key store, so groups of one warp can be in different phases.
template <bool EAGER>
__device__ FAST_POP_FN void fast_pop_loop(FastCold& f) {
  constexpr unsigned FULL = 0xFFFFFFFFu;
  const SearchParams& p = c_fp;
  const DevFstView& F = p.fst;
  const LeanCtx c = fast_ctx(p);
  const bool running = f.phase == kFastRun;
  const bool bfsm = EAGER && f.phase == kFastBfs;
  f.ev_rare = 0;
  // service wanted before any step: a smaller id below the window, a full heap pool / arena, a drained FIFO, a group
  // without a string
  f.ev_flags = (running ? (f.st.low_pending || f.st.heap_overflow || f.st.overflow)
                        : (bfsm ? (f.st.overflow || f.st.wline >= f.st.n_tuples) : f.phase != kFastDone)) ? 1u : 0u;
  if (__any_sync(FULL, f.ev_flags != 0)) return;
  uint32_t n_tuples = f.st.n_tuples, fmin12 = f.fmin12, rc = f.rc;
  uint32_t cursor = f.st.wline;                       // BFS phase: the FIFO cursor
  const uint32_t wline = f.st.wline, lev12 = f.lev12;
  // an idle group pops its slot 0 forever: fast_service left the idle key there (slab row S: no record ever hits)
  const uint32_t len = (running || bfsm) ? f.lhs.len : 0xFFFFFFFFu;
  const uint8_t* const str = (running || bfsm) ? f.lhs.s : p.bytes;
  const bool sorted = f.st.sorted;
  uint32_t* const key_of = reinterpret_cast<uint32_t*>(LEAN_KEYOF(p, c));
  uint32_t* const l0 = LEAN_L0(p, c);
  uint2* const tab = reinterpret_cast<uint2*>(c.base);
  // lane constants and shared-memory addresses pinned in registers (the compiler otherwise re-derives them from
  // SR_TID / SR_CgaCtaId in every step)
  unsigned lane = threadIdx.x & 7u, gbase = threadIdx.x & 24u;
  unsigned gmask = 0xFFu << gbase, ltmw = ((1u << lane) - 1u) << gbase;
  uint32_t win_s = (uint32_t)__cvta_generic_to_shared(LEAN_WIN(c));
  asm volatile("" : "+r"(lane), "+r"(gbase), "+r"(gmask), "+r"(ltmw), "+r"(win_s));
  const uint4* const slab_lane = F.islab + lane;
  const uint32_t smask = (1u << p.key_sbits) - 1u;
  for (;;) {
    const uint32_t w = fast_lds(win_s + lane * 4u);
    const unsigned gb = (__ballot_sync(FULL, w != 0) >> gbase) & 0xFFu;
    bool svc = running && gb == 0u;                                   // an empty window
    if (EAGER) svc = svc || (bfsm && cursor >= n_tuples);             // the FIFO is drained: the lattice is complete
    if (__any_sync(FULL, svc)) { f.ev_flags = svc ? 1u : 0u; break; }
    // pop the smallest ready id (search) / the next state of the FIFO (BFS) of every group and relax its leader records
    const int src = __ffs(gb) - 1;
    const uint32_t ww = __shfl_sync(FULL, w, (int)gbase + src);
    const uint32_t bit = __ffs(ww) - 1;
    uint32_t cur_id = running ? (wline << 8) + ((uint32_t)src << 5) + bit : 0u;
    if (EAGER && bfsm) cur_id = cursor++;
    if ((int)lane == src) fast_red_and(win_s + (uint32_t)src * 4u, ~(1u << bit));
    const uint32_t key = key_of[cur_id];
    const uint32_t P = key >> p.key_sbits, s2 = (key & smask) >> 1;
    const uint4 sa = __ldg(slab_lane + (size_t)s2 * kWaveSlots);    // {ilabel, next << 1 | eps, weight << 12, arcs folded}
    // dense index (SearchParams::pos_h): sa.y = next << 1 | epsilon; the per-pop parts are uniform over the group
    const uint32_t pb = p.pos_c + P * p.pos_m2;
    uint32_t own_d = 0u;                                              // BFS: the popped state's own distance (20 bits, all ones = none)
    if (EAGER && bfsm) own_d = tab[(key & smask) * p.pos_h + ((key & 1u) ? p.pos_k : 0u) + pb].y >> 12;
    uint32_t x = kFastNoLabel;
    if (P < len) x = (uint32_t)__ldg(str + P) + 1u;
    const bool is_match = sa.x == x, hit = is_match || sa.x == 0u;
    if (hit) rc += sa.w;
    const uint32_t pos = sa.y * p.pos_h + (is_match ? pb + p.pos_m2 : pb + p.pos_k);
    uint2* const recp = tab + pos;
    uint2 rec = make_uint2(~0u, ~0u);
    if (hit) rec = *recp;
    const uint32_t cand_hi = (lev12 | (cur_id >> 10)) + sa.z;        // new dist:20 | popped id bits 21..10
    bool is_new = hit && rec.y == 0xFFFFFFFFu;
    if (EAGER && bfsm) is_new = hit && ((rec.x & rec.y) == 0xFFFFFFFFu || !(rec.x & kCrecBfsBit));   // no BFS number yet (compose.zig:77-91)
    const unsigned nv = __ballot_sync(FULL, is_new);
    const uint32_t my_id = is_new ? n_tuples + __popc(nv & ltmw) : (rec.x & kCrecNone);
    bool lowered = hit && (cand_hi | 0xFFFu) < rec.y;                // :109-114, :137-142
    if (EAGER && bfsm) {
      lowered = false;
      // tight: both distances known and own + w == target's (a target the search never reached is never tight)
      const bool tight = own_d != 0xFFFFFu && (rec.y >> 12) != 0xFFFFFu && (own_d << 12) + sa.z == (rec.y & 0xFFFFF000u);
      const bool no_prev = (rec.y & 0xFFFu) == 0xFFFu && (rec.x >> 22) == 0x3FFu;
      const uint32_t pv = tight ? cur_id : kCrecNone;
      if (is_new) *recp = make_uint2((pv << 22) | my_id | kCrecBfsBit, (rec.y & 0xFFFFF000u) | (pv >> 10));
      else if (hit && tight && no_prev) *recp = make_uint2((cur_id << 22) | (rec.x & kCrecNone), (rec.y & 0xFFFFF000u) | (cur_id >> 10));
    } else {
      const unsigned long long cand = ((unsigned long long)cand_hi << 32) | (cur_id & 0x3FFu);
      const unsigned long long old = ((unsigned long long)rec.y << 32) | (rec.x >> 22);
      if (hit && cand < old) *recp = make_uint2((cur_id << 22) | my_id, cand_hi);   // :109-126
    }
    const uint32_t tkey = ((is_match ? P + 1u : P) << p.key_sbits) | sa.y;
    if (is_new) key_of[my_id] = tkey;
    n_tuples += __popc(nv & gmask);
    // queue: ready set at the current level, else the future set
    const bool fut = lowered && sa.z != 0u;
    bool low = false;
    if (lowered && sa.z == 0u) {
      const uint32_t line = my_id >> 8, bm = 1u << (my_id & 31u);
      if (line == wline) {
        fast_red_or(win_s + ((my_id >> 3) & 28u), bm);
      } else {
        atomicOr(&l0[my_id >> 5], bm);
        fast_red_or(win_s + 32u + ((line >> 5) << 2), 1u << (line & 31u));
        low = line < wline;
      }
#ifndef FSTB_FAST_NO_CHILDPF
      // a tuple with an old id that this step made ready is usually popped next: its slab row towards L1
      if (!is_new) asm volatile("prefetch.global.L1 [%0];" ::"l"(F.islab + (size_t)(sa.y >> 1) * kWaveSlots));
#endif
    }
    if (fut && cand_hi < fmin12) fmin12 = cand_hi;
    // the rare tail: final check at the end of the string, an id below the window, a future push into the radix
    // heap, a state wider than the slab, a full arena
    const bool fin = P == len && !(EAGER && bfsm && own_d == 0xFFFFFu), big = sa.x == kWaveBig;
    const unsigned rare = __ballot_sync(FULL, fin || low || (fut && sorted) || big || n_tuples > p.tuple_cap);
    if (rare) {
      f.ev_rare = rare; f.ev_flags = (fin ? 2u : 0u) | (low ? 4u : 0u) | (fut ? 8u : 0u) | (big ? 16u : 0u);
      f.ev_id = my_id; f.ev_nd = cand_hi >> 12; f.ev_bigno = sa.z; f.ev_cur = cur_id; f.ev_P = P; f.ev_s2 = s2; f.ev_x = x; f.ev_own = own_d;
      break;
    }
    __syncwarp();
  }
  f.st.n_tuples = n_tuples; f.fmin12 = fmin12; f.rc = rc;
  if (EAGER && bfsm) { f.st.wline = cursor; if (f.st.occ < n_tuples) f.st.occ = n_tuples; }
}

template <bool EAGER>
__global__ void __launch_bounds__(128, FSTB_FAST_MINBLOCKS) csp_batch_fast_kernel() {
  constexpr int G = 8;
  constexpr unsigned FULL = 0xFFFFFFFFu;
  const SearchParams& p = c_fp;
  const Group<G> g;
  const LeanCtx c = fast_ctx(p);
  for (uint32_t i = 128 + g.lane; i < p.smem_words; i += G) c.sm[i] = 0;   // window, summary and cold state start empty
  g.sync();
  FastCold f;
  f.st.n_tuples = 0; f.st.wline = kNone; f.st.relax_calls = 0; f.st.last = 0; f.st.future_min = ~0ull;
  f.st.low_pending = false; f.st.overflow = false; f.st.heap_overflow = false; f.st.sorted = false; f.st.bfs_started = false; f.st.stuck = false;
  f.st.occ = 0; f.st.wide = false;
  f.lhs.s = nullptr; f.lhs.len = 0;
  f.phase = kFastFetch; f.idx = 0; f.steps = 0; f.lev12 = 0; f.fmin12 = ~0u; f.rc = 0; f.relax_total = 0; f.tuple_total = 0;
  for (;;) {
    fast_pop_loop<EAGER>(f);
    if (f.ev_rare) {
      if (f.ev_rare & g.mask)   // group-local
        fast_rare<EAGER>(p, f, (f.ev_flags & 2u) != 0, (f.ev_flags & 4u) != 0, (f.ev_flags & 8u) != 0, f.ev_id, f.ev_nd, (f.ev_flags & 16u) != 0,
                         f.ev_bigno, f.ev_cur, f.ev_P, f.ev_s2, f.ev_x, f.ev_own);
      __syncwarp();
    } else {
      if (f.ev_flags & 1u) fast_service<EAGER>(p, f);
      if (__all_sync(FULL, f.phase == kFastDone)) break;
    }
  }
  unsigned long long relax_total = f.relax_total;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) relax_total += __shfl_xor_sync(FULL, relax_total, o);
  if ((threadIdx.x & 31u) == 0 && relax_total) atomicAdd(p.relax_counter, relax_total);
  if (g.lane == 0 && f.tuple_total) atomicAdd(p.tuple_counter, f.tuple_total);
}

}  // namespace fstb200
