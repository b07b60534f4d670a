// Wavefront exact search: ONE WARP PER STRING, up to 32 pops per step.
//
// STATUS: exact (parity-tested like every other engine) but NOT used by default — measured 4x slower than the lean
// kernel on the bench transducers (profiles/README.md): the reference's pop order keeps jumping back to tuples the
// current expansion has just lowered, so only 1.6-2 consecutive pops are order-independent on average and most
// chunks are cut to a single tuple.  Kept selectable (`engine = 4..6`) for transducers whose levels are wide
// breadth-first sweeps (all-zero weights without re-lowering), where whole ready words do pop together.
//
// The lean kernel (csp_lean.cuh) emulates the reference's pop sequence one tuple at a time and takes all of
// its parallelism from the batch; the per-string latency (one dependent chain window -> key -> arcs -> table
// per pop) times the number of strings that fit HBM bounds its throughput.  This kernel keeps the same data
// structures (table, id -> key array, ready bitmap + window, radix heap, back-track; all reused from
// csp_lean.cuh with G = 32) but pops a whole READY WORD at once: the up to 32 unsettled tuples of the current
// distance level whose ids share one 32-id word of the ready bitmap, lane b <-> id base + b.  Their
// expansions run in parallel, one lane per popped tuple looping over the (at most kWaveSlots) folded search
// records of its transducer state, and the relaxations of the whole chunk are merged so that the table, the
// ids and the queues end up exactly as if the tuples had been popped one after the other in id order
// (compose-shortest-path.zig:159-366):
//
//   * ORDER.  A candidate relaxation is (lane r, slot j); the reference's order is r-major (pops in id order),
//     then the expansion order of one pop (match arcs :182-202, then input-epsilon arcs :254-278) — the wave
//     slab stores a state's records in exactly that order, so `order = r * kWaveSlots + j`.
//   * ARBITRATION (shared memory, per warp).  Candidates are grouped by target tuple in a small open-addressing
//     table keyed by the compact tuple key; per target it collects  first = min order  (the first toucher:
//     creates the tuple if it is new, so new ids are numbered in `first` order = prefix sums over lanes, :70-89),
//     ndmin = min new distance, and  best = min order among the candidates reaching ndmin  (all pops of a chunk
//     have the level distance D, so the sequential relax rule :91-144 ends with dist = min(old, ndmin) and, where
//     the distance is lowered or tied, prev = the smallest popped id among the candidates reaching it).
//     The first toucher alone then reads the table record, applies the rule once and writes it back: one global
//     load + one store per DISTINCT target of the chunk and no global atomics.
//   * VALIDITY.  Popping u_0 < u_1 < ... together is the reference's sequence iff no expansion of the chunk
//     inserts an id below a later member into the ready set.  New tuples get ids above every existing id; the
//     only other insertions are EXISTING tuples lowered to the level distance.  If one of those has an id below
//     the chunk's largest member, nothing has been written yet: the chunk is abandoned and ONE tuple is popped
//     by the lean step (always exact).  States wider than the slab and sparse ready words go the same way.
//   * A stale future-set push of the reference (target created above the level and lowered to it later in the
//     same chunk) is simply not made; stale entries are skipped by the reference (:161-162) and by the radix heap.
#pragma once
#include "csp_lean.cuh"

namespace fstb200 {

constexpr uint32_t kWaveArb = 512;                   // arbitration slots per warp (>= 2 x 32 x kWaveSlots candidates)
constexpr uint32_t kWaveArbWords = kWaveArb * 5;     // ndmin u64 + key u32 + first u32 + best u32
constexpr uint32_t kWaveMinPops = 2;                 // ready words with fewer tuples take the single-pop path

__host__ __device__ inline uint32_t wave_lean_words(uint32_t n1) { return (128u + 32u + n1 + kLeanColdWords + 3u) & ~3u; }

struct WaveArb {
  unsigned long long* nd;   // [kWaveArb] smallest new distance (bit pattern), ~0 = none
  uint32_t* key;            // [kWaveArb] compact tuple key, kNone = empty
  uint32_t* first;          // [kWaveArb] smallest candidate order
  uint32_t* best;           // [kWaveArb] smallest candidate order among those reaching nd
};

__device__ __forceinline__ uint32_t wave_arb_insert(const WaveArb& a, uint32_t key) {
  uint32_t h = (key * 0x9E3779B1u) >> 23;   // 9 bits
  for (;;) {
    const uint32_t old = atomicCAS(&a.key[h], kNone, key);
    if (old == kNone || old == key) return h;
    h = (h + 1u) & (kWaveArb - 1u);
  }
}
__device__ __forceinline__ void wave_arb_reset(const WaveArb& a, uint32_t h) {
  a.key[h] = kNone; a.first[h] = kNone; a.best[h] = kNone; a.nd[h] = ~0ull;
}

// One chunk: pop every set bit of window word `src` (`ww`).  True = done (or the search state says stop);
// false = nothing was changed, take the single-pop path for this step.
template <bool DENSE>
__device__ __forceinline__ bool wave_chunk(const SearchParams& p, const Group<32>& g, const LeanCtx& c, LeanState& st, const LhsBytes& lhs,
                                           const WaveArb& arb, uint32_t src, uint32_t ww, unsigned long long& rc_lane) {
  const DevFstView& F = p.fst;
  const uint32_t lane = g.lane;
  bool popping = (ww >> lane) & 1u;
  const uint32_t cur_id = ((st.wline * 32u + src) << 5) + lane;
  uint32_t s1 = 0, sf = 0;
  if (popping) lean_keyof_load<DENSE>(p, c, cur_id, s1, sf);
  const uint32_t s2 = sf >> 1;
  // the state's search records: matching labels in frozen order, then the input-epsilon records
  uint4 sa[kWaveSlots];
#pragma unroll
  for (uint32_t j = 0; j < kWaveSlots; j++) sa[j] = make_uint4(kNone, 0u, 0u, 0u);
  if (popping) {
    const uint4* rp = F.wslab + (uint64_t)s2 * kWaveSlots;
#pragma unroll
    for (uint32_t j = 0; j < kWaveSlots; j++) sa[j] = __ldg(rp + j);
  }
  const unsigned bigm = g.ballot(popping && sa[0].x == kWaveBig);
  if (bigm) {
    const uint32_t fb = __ffs(bigm) - 1;
    ww &= (1u << fb) - 1u;
    if (ww == 0) return false;          // the first tuple of the word is wide: single-pop path
    popping = popping && lane < fb;     // pop the tuples before it
  }
  const uint32_t x = (popping && s1 < lhs.len) ? (uint32_t)__ldg(lhs.s + s1) + 1u : 0xFFFDu;
  const unsigned long long Db = st.last;
  const double D = __longlong_as_double((long long)Db);
  const uint32_t sbits = p.key_sbits;

  // ── phase A: candidates enter the arbitration table ──
  uint32_t candbits = 0;
  uint32_t ck[kWaveSlots], slot[kWaveSlots];
  unsigned long long ndb[kWaveSlots];
#pragma unroll
  for (uint32_t j = 0; j < kWaveSlots; j++) {
    const uint32_t lab = sa[j].x & 0xFFFFu;
    const bool ism = lab == x, ise = lab == 0u;
    ck[j] = 0; slot[j] = 0; ndb[j] = 0;
    if (popping && (ism || ise)) {
      rc_lane += sa[j].x >> 16;                                          // relax calls of the reference: every folded arc
      ck[j] = ((s1 + (ism ? 1u : 0u)) << sbits) | (sa[j].y << 1) | (ise ? 1u : 0u);
      ndb[j] = (unsigned long long)__double_as_longlong(D + __hiloint2double((int)sa[j].w, (int)sa[j].z));
      const uint32_t h = wave_arb_insert(arb, ck[j]);
      slot[j] = h;
      atomicMin(&arb.first[h], lane * kWaveSlots + j);
      atomicMin(&arb.nd[h], ndb[j]);
      candbits |= 1u << j;
    }
  }
  __syncwarp();
  // ── phase B: first candidate reaching the smallest distance ──
#pragma unroll
  for (uint32_t j = 0; j < kWaveSlots; j++) {
    if (((candbits >> j) & 1u) && arb.nd[slot[j]] == ndb[j]) atomicMin(&arb.best[slot[j]], lane * kWaveSlots + j);
  }
  __syncwarp();
  // ── phase C: the first toucher of every target reads its record ──
  const uint32_t maxpop = ((st.wline * 32u + src) << 5) + (31u - __clz(ww));
  uint32_t repbits = 0, newbits = 0;
  bool viol = false;
  uint32_t pos[kWaveSlots], oid[kWaveSlots], oprev[kWaveSlots];
  double odist[kWaveSlots];
#pragma unroll
  for (uint32_t j = 0; j < kWaveSlots; j++) {
    pos[j] = 0; oid[j] = kNone; oprev[j] = kNone; odist[j] = d_inf();
    if (((candbits >> j) & 1u) && arb.first[slot[j]] == lane * kWaveSlots + j) {
      repbits |= 1u << j;
      lean_lookup<DENSE>(p, c, ck[j] >> sbits, ck[j] & ((1u << sbits) - 1u), pos[j], odist[j], oid[j], oprev[j]);
      if (oid[j] == kNone) {
        newbits |= 1u << j;
      } else {
        // an existing tuple lowered to the level distance joins the ready set: it must pop after the whole chunk
        const unsigned long long ndm = arb.nd[slot[j]];
        if (ndm == Db && (unsigned long long)__double_as_longlong(odist[j]) > Db && oid[j] < maxpop) viol = true;
      }
    }
  }
  // new ids: prefix sums in candidate order (:80-87)
  const uint32_t nnew = __popc(newbits);
  uint32_t pre = nnew;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, pre, o); if ((int)lane >= o) pre += t; }
  const uint32_t total_new = __shfl_sync(0xFFFFFFFFu, pre, 31);
  pre -= nnew;
  const bool too_many = st.n_tuples + total_new > p.tuple_cap;
  if (g.any(viol) || too_many) {
    // nothing was written to the search state: forget the chunk
#pragma unroll
    for (uint32_t j = 0; j < kWaveSlots; j++) if ((candbits >> j) & 1u) wave_arb_reset(arb, slot[j]);
    __syncwarp();
    if (too_many) { st.overflow = true; return true; }
    return false;
  }
  // ── phase D: commit ──
#pragma unroll
  for (uint32_t j = 0; j < kWaveSlots; j++) {
    const bool rep = (repbits >> j) & 1u;
    const uint32_t b = rep ? arb.best[slot[j]] : 0u;
    const uint32_t ub = __shfl_sync(0xFFFFFFFFu, cur_id, (int)(b / kWaveSlots));
    bool to_future = false; uint32_t my_id = 0; unsigned long long ndm = 0;
    if (rep) {
      ndm = arb.nd[slot[j]];
      wave_arb_reset(arb, slot[j]);
      const double nd = __longlong_as_double((long long)ndm);
      const uint32_t P = ck[j] >> sbits, SF = ck[j] & ((1u << sbits) - 1u);
      bool lowered;
      if ((newbits >> j) & 1u) {
        my_id = st.n_tuples + pre + __popc(newbits & ((1u << j) - 1u));
        lean_keyof_store<DENSE>(p, c, my_id, P, SF);
        const uint32_t ps = lean_claim<DENSE>(p, c, P, SF, pos[j]);
        lean_store<DENSE>(c, ps, P, SF, nd, my_id, ub);
        lowered = true;
      } else {
        my_id = oid[j];
        lowered = nd < odist[j];                                                                  // :109-114, :137-142
        const bool take = lowered || (nd == odist[j] && (oprev[j] == kNone || ub < oprev[j]));    // :115-126
        if (take) lean_store<DENSE>(c, pos[j], P, SF, nd, my_id, ub);
      }
      if (lowered) {
        if (ndm == Db) {   // ready at this level; never below the window (see VALIDITY)
          const uint32_t line = my_id >> 10;
          const uint32_t bit = 1u << (my_id & 31u);
          if (line == st.wline) {
            atomicOr(&LEAN_WIN(c)[(my_id >> 5) & 31u], bit);
          } else {
            atomicOr(&LEAN_L0(p, c)[my_id >> 5], bit);
            atomicOr(&LEAN_L1(c, 32)[line >> 5], 1u << (line & 31u));
          }
        } else {
          to_future = true;
          if (ndm < st.future_min) st.future_min = ndm;
        }
      }
    }
    if (st.sorted && !st.heap_overflow) {
      if (g.any(to_future)) lean_bucket_push<32>(p, g, c, st, to_future, my_id, to_future ? bucket_of(ndm, st.last) : 1u, ndm);
    }
  }
  st.n_tuples += total_new;
  // ── final check of the popped tuples (:165-179): smallest (total, id) ──
  {
    bool fin = false; double final_w = 0.0, total = 0.0;
    if (popping && s1 == lhs.len) {
      const double fw2 = F.final_w[s2];
      if (!d_isinf(fw2)) { fin = true; final_w = 0.0 + fw2; total = D + final_w; }
    }
    if (g.any(fin)) {
      unsigned long long tk = fin ? (unsigned long long)__double_as_longlong(total) : ~0ull;
      uint32_t tid = fin ? cur_id : kNone;
      unsigned long long fk = (unsigned long long)__double_as_longlong(final_w);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long tk2 = __shfl_xor_sync(0xFFFFFFFFu, tk, o);
        const uint32_t tid2 = __shfl_xor_sync(0xFFFFFFFFu, tid, o);
        const unsigned long long fk2 = __shfl_xor_sync(0xFFFFFFFFu, fk, o);
        if (tk2 < tk || (tk2 == tk && tid2 < tid)) { tk = tk2; tid = tid2; fk = fk2; }
      }
      uint32_t* cold = LEAN_COLD(p, c, 32);
      const bool have_best = cold[kcHaveBest] != 0;
      const double best_total = __hiloint2double((int)cold[kcBestTotHi], (int)cold[kcBestTotLo]);
      const uint32_t best_id = cold[kcBestId];
      const double tot = __longlong_as_double((long long)tk);
      __syncwarp();
      if (!have_best || tot < best_total || (tot == best_total && tid < best_id)) {
        if (lane == 0) {
          cold[kcHaveBest] = 1; cold[kcBestId] = tid;
          cold[kcBestFwLo] = (uint32_t)fk; cold[kcBestFwHi] = (uint32_t)(fk >> 32);
          cold[kcBestTotLo] = (uint32_t)tk; cold[kcBestTotHi] = (uint32_t)(tk >> 32);
        }
      }
    }
  }
  // the popped tuples are settled: clear their bits (commutes with the inserts above: other bits)
  if (lane == 0) atomicAnd(&LEAN_WIN(c)[src], ~ww);
  __syncwarp();
  return true;
}

// One step of a string: a chunk if the lowest ready word allows it, else the lean single-pop step (which also
// switches the window line and advances the distance level).
template <bool DENSE>
__device__ __forceinline__ bool wave_step(const SearchParams& p, const Group<32>& g, const LeanCtx& c, LeanState& st, const LhsBytes& lhs,
                                          const WaveArb& arb, unsigned long long& rc_lane, uint32_t* stats) {
  bool done = false;
  if (!(st.overflow || st.heap_overflow || st.low_pending)) {
    const uint32_t w = LEAN_WIN(c)[g.lane];
    const unsigned bal = g.ballot(w != 0);
    if (bal) {
      const uint32_t src = __ffs(bal) - 1;
      const uint32_t ww = g.shfl(w, (int)src);
      if (__popc(ww) >= (int)kWaveMinPops) {
        done = wave_chunk<DENSE>(p, g, c, st, lhs, arb, src, ww, rc_lane);
        if (done) { stats[0]++; stats[1] += __popc(ww); } else stats[3]++;
      }
    }
  }
  if (done) return true;
  stats[2]++;
  return lean_step<32, DENSE, false, false>(p, g, c, st, lhs, 1u);
}

// Persistent batch kernel: one warp per string.
template <bool DENSE>
__global__ void __launch_bounds__(128, 4) csp_batch_wave_kernel(const __grid_constant__ SearchParams p) {
  extern __shared__ __align__(16) uint32_t smem_all[];
  const Group<32> g;
  const uint32_t wib = threadIdx.x / 32;
  const uint32_t gslot = blockIdx.x * (blockDim.x / 32) + wib;
  LeanCtx c;
  c.base = p.arena + (uint64_t)gslot * p.arena_stride;
  c.sm = smem_all + (size_t)wib * p.smem_words;
  const uint32_t lw = wave_lean_words(p.n1);
  for (uint32_t i = 128 + g.lane; i < lw; i += 32) c.sm[i] = 0;                        // window, summary, cold state
  for (uint32_t i = lw + g.lane; i < lw + kWaveArbWords; i += 32) c.sm[i] = kNone;     // arbitration table: all empty
  WaveArb arb;
  arb.nd = reinterpret_cast<unsigned long long*>(c.sm + lw);
  arb.key = c.sm + lw + 2 * kWaveArb;
  arb.first = arb.key + kWaveArb;
  arb.best = arb.first + kWaveArb;
  __syncwarp();
  LeanState st;
  st.n_tuples = 0; st.wline = kNone; st.relax_calls = 0; st.last = 0; st.future_min = ~0ull;
  st.low_pending = false; st.overflow = false; st.heap_overflow = false; st.sorted = false; st.bfs_started = false; st.stuck = false; st.occ = 0;
  unsigned long long relax_total = 0, tuple_total = 0;
  uint32_t stats[4] = {0, 0, 0, 0};   // chunk steps, tuples popped by chunks, single-pop steps, abandoned chunks
  for (;;) {
    uint32_t item = 0;
    if (g.lane == 0) item = atomicAdd(p.queue_head, 1u);
    item = g.shfl(item, 0);
    if (item >= p.n_items) break;
    const uint32_t idx = p.order ? p.order[item] : item;
    LhsBytes lhs; lhs.s = p.bytes + p.offsets[idx]; lhs.len = (uint32_t)(p.offsets[idx + 1] - p.offsets[idx]);
    const int32_t pre = p.skip ? p.skip[idx] : kStPath;
    if (p.fst.start == kNone || pre != kStPath) {
      if (g.lane == 0) { p.status[idx] = pre != kStPath ? pre : kStNoPath; p.path_len[idx] = 0; p.pool_off[idx] = 0; p.final_w[idx] = d_inf(); p.n_tuples[idx] = 0; }
      continue;
    }
    lean_begin<32, DENSE>(p, g, c, st);
    unsigned long long rc_lane = 0;
    uint32_t steps = 0;
    while (wave_step<DENSE>(p, g, c, st, lhs, arb, rc_lane, stats)) {
      if (++steps > 8u * p.tuple_cap + 4096u) { st.stuck = true; break; }   // safety valve, see csp_lean.cuh
    }
    uint32_t plen; uint64_t poff; double fw;
    const int32_t status = lean_finish<32, DENSE>(p, g, c, st, lhs, &plen, &poff, &fw);
    if (g.lane == 0) {
      p.status[idx] = status; p.path_len[idx] = plen; p.pool_off[idx] = poff; p.final_w[idx] = fw; p.n_tuples[idx] = st.n_tuples;
    }
    relax_total += st.relax_calls + rc_lane; tuple_total += st.n_tuples;   // relax counts are per-lane partial sums
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) relax_total += __shfl_xor_sync(0xFFFFFFFFu, relax_total, o);
  if (g.lane == 0) {
    if (relax_total) atomicAdd(p.relax_counter, relax_total);
    if (tuple_total) atomicAdd(p.tuple_counter, tuple_total);
    if (p.wave_stats) { for (int k = 0; k < 4; k++) atomicAdd(p.wave_stats + k, (unsigned long long)stats[k]); }
  }
}

}  // namespace fstb200
