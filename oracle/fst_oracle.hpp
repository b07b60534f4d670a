// CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// A C++17 restatement of the reference (ontypehq/libfst, Zig) algorithms on the
// compose+shortest-path hot path.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may build, link or call this.
// The product library (libfst_b200.so) never includes or links anything here.
//
// Every function cites the reference file:line it follows (paths are relative
// to the reference tree).  The Zig reference cannot be compiled in this image
// (no zig toolchain), so this restatement is pinned by:
//   * the reference's own unit-test vectors for this path
//     (src/ops/compose-shortest-path.zig:424-471, src/ops/shortest-path.zig:143-216,
//      src/ops/compose.zig:311-354, src/ops/rewrite.zig:244-498, src/fst.zig:295-492),
//   * the survey's independent Python restatement signatures.
// Tie-break behaviour is NOT exercised by any reference test ("parity unpinned by
// tests; pinned by source reading") — see DESIGN.md.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <queue>
#include <string>
#include <vector>

namespace orc {

using Label = uint32_t;    // src/arc.zig:4
using StateId = uint32_t;  // src/arc.zig:7
constexpr Label kEpsilon = 0;                  // src/arc.zig:10
constexpr StateId kNoState = 0xFFFFFFFFu;      // src/arc.zig:13
constexpr double kInf = std::numeric_limits<double>::infinity();

// ── Tropical semiring: src/weight.zig:5-37 ──
inline bool w_is_zero(double a) { return std::isinf(a); }  // weight.zig:30-32 (isInf: ±inf)
inline double w_times(double a, double b) {                 // weight.zig:19-23
  if (w_is_zero(a) || w_is_zero(b)) return kInf;
  return a + b;
}
// weight.zig:34-37 — math.order on the raw f64 (NaN is rejected before we get here).
inline int w_compare(double a, double b) { return a < b ? -1 : (a > b ? 1 : 0); }

struct Arc {  // src/arc.zig:17-22
  Label ilabel;
  Label olabel;
  double weight;
  StateId nextstate;
};

// src/arc.zig:46-54 — total order used by freeze.
inline bool compare_by_ilabel(const Arc& a, const Arc& b) {
  if (a.ilabel != b.ilabel) return a.ilabel < b.ilabel;
  if (a.olabel != b.olabel) return a.olabel < b.olabel;
  int c = w_compare(a.weight, b.weight);
  if (c < 0) return true;
  if (c > 0) return false;
  return a.nextstate < b.nextstate;
}

// ── MutableFst: src/mutable-fst.zig:45-218 (only what the path uses) ──
struct MutableFst {
  std::vector<double> finals;
  std::vector<std::vector<Arc>> arcs_;
  StateId start_ = kNoState;

  StateId add_state() {  // mutable-fst.zig:96-101
    finals.push_back(kInf);
    arcs_.emplace_back();
    return (StateId)(finals.size() - 1);
  }
  void add_states(size_t n) { for (size_t i = 0; i < n; i++) add_state(); }  // :103-109
  void set_start(StateId s) { start_ = s; }                                   // :111-116
  void set_final(StateId s, double w) { finals[s] = w; }                      // :118-122
  void add_arc(StateId src, const Arc& a) { arcs_[src].push_back(a); }        // :124-127
  void sort_all_arcs() {  // mutable-fst.zig:148-153 (std.mem.sort is stable)
    for (auto& v : arcs_) std::stable_sort(v.begin(), v.end(), compare_by_ilabel);
  }
  StateId start() const { return start_; }                                    // :157
  double final_weight(StateId s) const { return finals[s]; }                  // :161
  size_t num_states() const { return finals.size(); }                         // :169
  const std::vector<Arc>& arcs(StateId s) const { return arcs_[s]; }          // :183
  size_t total_arcs() const { size_t t = 0; for (auto& v : arcs_) t += v.size(); return t; }
};

// ── Frozen layout: src/fst.zig:16-40 ──
struct StateEntry { uint32_t arc_offset; uint32_t num_arcs; double final_weight; };
struct PackedArc { uint32_t ilabel; uint32_t olabel; double weight; uint32_t nextstate; uint32_t _pad; };
struct Header {
  uint32_t magic; uint16_t version; uint8_t weight_type; uint8_t flags;
  uint32_t num_states; uint32_t num_arcs; uint32_t start_state; uint32_t _padding;
};
static_assert(sizeof(StateEntry) == 16 && sizeof(PackedArc) == 24 && sizeof(Header) == 24, "layout");
constexpr uint32_t kMagic = 0x46535421;  // fst.zig:12
constexpr uint16_t kVersion = 1;         // fst.zig:13

struct ArcSpan {
  const PackedArc* p; size_t n;
  const PackedArc* begin() const { return p; }
  const PackedArc* end() const { return p + n; }
  size_t size() const { return n; }
};

struct Fst {
  std::vector<uint8_t> bytes;  // Header | StateEntry[] | PackedArc[]

  const Header& header() const { return *reinterpret_cast<const Header*>(bytes.data()); }
  const StateEntry* state_table() const {
    return reinterpret_cast<const StateEntry*>(bytes.data() + sizeof(Header));
  }
  const PackedArc* arc_table() const {
    return reinterpret_cast<const PackedArc*>(bytes.data() + sizeof(Header) +
                                              (size_t)header().num_states * sizeof(StateEntry));
  }
  StateId start() const { return header().start_state; }                       // fst.zig:84-86
  uint32_t num_states() const { return header().num_states; }                  // :88-90
  double final_weight(StateId s) const { return state_table()[s].final_weight; }  // :96-98
  ArcSpan arcs(StateId s) const {                                              // :104-107
    const StateEntry& e = state_table()[s];
    return {arc_table() + e.arc_offset, e.num_arcs};
  }
  // fst.zig:112-136 — lower bound then upper bound (starting from lo).
  ArcSpan arcs_by_ilabel(StateId s, Label ilabel) const {
    ArcSpan sa = arcs(s);
    size_t lo = 0, hi = sa.n;
    while (lo < hi) {
      size_t mid = lo + (hi - lo) / 2;
      if (sa.p[mid].ilabel < ilabel) lo = mid + 1; else hi = mid;
    }
    size_t start_idx = lo;
    hi = sa.n;
    while (lo < hi) {
      size_t mid = lo + (hi - lo) / 2;
      if (sa.p[mid].ilabel <= ilabel) lo = mid + 1; else hi = mid;
    }
    return {sa.p + start_idx, lo - start_idx};
  }

  // fst.zig:160-224 — freeze (sorts the mutable's arcs in place, like the reference).
  static Fst from_mutable(MutableFst& m) {
    m.sort_all_arcs();
    uint32_t ns = (uint32_t)m.num_states();
    uint32_t total = (uint32_t)m.total_arcs();
    Fst f;
    f.bytes.assign(sizeof(Header) + (size_t)ns * sizeof(StateEntry) + (size_t)total * sizeof(PackedArc), 0);
    Header* h = reinterpret_cast<Header*>(f.bytes.data());
    h->magic = kMagic; h->version = kVersion; h->weight_type = 0; h->flags = 0;
    h->num_states = ns; h->num_arcs = total; h->start_state = m.start(); h->_padding = 0;
    StateEntry* st = reinterpret_cast<StateEntry*>(f.bytes.data() + sizeof(Header));
    uint32_t off = 0;
    for (uint32_t i = 0; i < ns; i++) {
      st[i].arc_offset = off; st[i].num_arcs = (uint32_t)m.arcs(i).size(); st[i].final_weight = m.final_weight(i);
      off += st[i].num_arcs;
    }
    PackedArc* ap = reinterpret_cast<PackedArc*>(f.bytes.data() + sizeof(Header) + (size_t)ns * sizeof(StateEntry));
    size_t ai = 0;
    for (uint32_t i = 0; i < ns; i++)
      for (const Arc& a : m.arcs(i)) { ap[ai].ilabel = a.ilabel; ap[ai].olabel = a.olabel; ap[ai].weight = a.weight; ap[ai].nextstate = a.nextstate; ap[ai]._pad = 0; ai++; }
    return f;
  }

  // fst.zig:227-273 — validation of an external byte image.  Returns false on
  // any of the reference's error.Invalid* conditions.
  static bool from_bytes(const uint8_t* data, size_t len, Fst* out) {
    if (len < sizeof(Header)) return false;
    Header h; std::memcpy(&h, data, sizeof(h));
    if (h.magic != kMagic) return false;
    if (h.version != kVersion) return false;
    if (h.weight_type != 0) return false;
    size_t expected = sizeof(Header) + (size_t)h.num_states * sizeof(StateEntry) + (size_t)h.num_arcs * sizeof(PackedArc);
    if (len != expected) return false;
    if (h.num_states > 0 && h.start_state != kNoState && h.start_state >= h.num_states) return false;
    if (h.num_states == 0 && h.start_state != kNoState) return false;
    Fst f; f.bytes.assign(data, data + len);
    const StateEntry* st = f.state_table();
    const PackedArc* all = f.arc_table();
    for (uint32_t i = 0; i < h.num_states; i++) {
      if (st[i].arc_offset > h.num_arcs) return false;
      if (st[i].num_arcs > h.num_arcs - st[i].arc_offset) return false;
      bool have = false; Label last = 0;
      for (uint32_t j = 0; j < st[i].num_arcs; j++) {
        const PackedArc& a = all[st[i].arc_offset + j];
        if (a.nextstate >= h.num_states) return false;
        if (have && a.ilabel < last) return false;
        last = a.ilabel; have = true;
      }
    }
    *out = std::move(f);
    return true;
  }
};

// ── String helpers: src/string.zig ──
// string.zig:24-50 (compileStringTransducer; compileString passes input twice).
inline MutableFst compile_string_transducer(const uint8_t* in, size_t in_len, const uint8_t* out, size_t out_len) {
  MutableFst f;
  size_t max_len = std::max(in_len, out_len);
  if (max_len == 0) { StateId s = f.add_state(); f.set_start(s); f.set_final(s, 0.0); return f; }
  f.add_states(max_len + 1);
  f.set_start(0);
  f.set_final((StateId)max_len, 0.0);
  for (size_t i = 0; i < max_len; i++) {
    Label il = i < in_len ? (Label)in[i] + 1 : kEpsilon;
    Label ol = i < out_len ? (Label)out[i] + 1 : kEpsilon;
    f.add_arc((StateId)i, Arc{il, ol, 0.0, (StateId)(i + 1)});
  }
  return f;
}
inline MutableFst compile_string(const uint8_t* in, size_t n) { return compile_string_transducer(in, n, in, n); }

// string.zig:64-97.  Returns false for "null" (not a linear chain / no start).
// Deviation (documented): the reference loops forever on a final-less cycle and
// has checked-UB for labels > 256; we return false in both cases.
inline bool print_string_from_tape(const MutableFst& f, bool output_tape, std::string* out) {
  StateId cur = f.start();
  if (cur == kNoState) return false;
  out->clear();
  size_t steps = 0;
  while (true) {
    if (!w_is_zero(f.final_weight(cur)) && f.arcs(cur).empty()) break;
    const auto& sa = f.arcs(cur);
    if (sa.size() != 1) return false;
    const Arc& a = sa[0];
    Label l = output_tape ? a.olabel : a.ilabel;
    if (l != kEpsilon) { if (l > 256) return false; out->push_back((char)(uint8_t)(l - 1)); }
    cur = a.nextstate;
    if (cur == kNoState) return false;
    if (++steps > f.num_states()) return false;
  }
  return true;
}

// ── Left-operand views (the reference duck-types fst1; we template on it) ──
struct MutableLhs {
  const MutableFst* f;
  StateId start() const { return f->start(); }
  double final_weight(StateId s) const { return f->final_weight(s); }
  const std::vector<Arc>& arcs(StateId s) const { return f->arcs(s); }
};

// ── Search statistics (not in the reference; used to size and report work) ──
struct SearchStats {
  uint64_t tuples = 0;       // N: tuples created (compose-shortest-path.zig:80-87)
  uint64_t relax_calls = 0;  // R: calls of relax (:91) == arcs of compose(a,b)
  uint64_t pushes = 0;       // queue pushes (:138, :153)
  uint64_t retakes = 0;      // equal-distance takes with an existing back-pointer (:115-122)
  uint64_t pops = 0;         // non-stale pops (:163)
};

enum class Status : int { kOk = 0, kEmpty = 1, kUnsupportedN = 2, kBacktrackCycle = 3 };

struct PathArc { Label ilabel; Label olabel; double weight; };
struct PathResult {
  Status status = Status::kEmpty;
  std::vector<PathArc> arcs;  // chain arc i: state i -> i+1
  double final_weight = kInf; // final weight on the last state
  SearchStats stats;
  // identification of the chosen final tuple (diagnostics)
  uint32_t final_s1 = kNoState, final_s2 = kNoState, final_filter = 0;
  double total() const {      // what a caller sums: arcs left to right, then final
    double t = 0.0; for (auto& a : arcs) t += a.weight; return t + final_weight;
  }
};

// Open-addressing tuple → id map.  The reference uses std.AutoHashMapUnmanaged
// (compose-shortest-path.zig:63); only exact key→value semantics are observable.
class TupleMap {
 public:
  TupleMap() { keys_.assign(1024, kEmptyKey); vals_.assign(1024, 0); mask_ = 1023; }
  static uint64_t pack(uint32_t s1, uint32_t s2, uint8_t f) {
    // s1 < 2^30 is asserted by the callers; 2 filter bits.
    return ((uint64_t)s1 << 34) | ((uint64_t)s2 << 2) | f;
  }
  bool get(uint64_t k, uint32_t* v) const {
    size_t i = hash(k) & mask_;
    while (keys_[i] != kEmptyKey) { if (keys_[i] == k) { *v = vals_[i]; return true; } i = (i + 1) & mask_; }
    return false;
  }
  void put(uint64_t k, uint32_t v) {
    if ((count_ + 1) * 10 > (mask_ + 1) * 7) grow();
    size_t i = hash(k) & mask_;
    while (keys_[i] != kEmptyKey) i = (i + 1) & mask_;
    keys_[i] = k; vals_[i] = v; count_++;
  }
 private:
  static constexpr uint64_t kEmptyKey = ~0ull;
  static uint64_t hash(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
  }
  void grow() {
    std::vector<uint64_t> ok; std::vector<uint32_t> ov; ok.swap(keys_); ov.swap(vals_);
    size_t n = (mask_ + 1) * 2; keys_.assign(n, kEmptyKey); vals_.assign(n, 0); mask_ = n - 1; count_ = 0;
    for (size_t i = 0; i < ok.size(); i++) if (ok[i] != kEmptyKey) put(ok[i], ov[i]);
  }
  std::vector<uint64_t> keys_; std::vector<uint32_t> vals_; size_t mask_ = 0, count_ = 0;
};

// ── composeShortestPath: src/ops/compose-shortest-path.zig:26-401 ──
// Frozen rhs variant (rhs_has_label_lookup == true).
template <class Lhs>
PathResult compose_shortest_path(const Lhs& fst1, const Fst& fst2, uint32_t n) {
  PathResult res;
  // :30-33
  if (fst1.start() == kNoState || fst2.start() == kNoState || n == 0) { res.status = Status::kEmpty; return res; }
  if (n != 1) { res.status = Status::kUnsupportedN; return res; }

  struct StateTuple { StateId s1, s2; uint8_t filter; };      // :39-43
  struct BackPtr { bool has; uint32_t prev_id; Label ilabel, olabel; double weight; };  // :44-49 (+optional)
  struct QueueItem { uint32_t tuple_id; double dist; };        // :50-53
  struct QueueCompare {                                         // :55-61 (min-heap on (dist, id))
    bool operator()(const QueueItem& a, const QueueItem& b) const {
      int c = w_compare(a.dist, b.dist);
      if (c != 0) return c > 0;
      return a.tuple_id > b.tuple_id;
    }
  };

  TupleMap tuple_to_id;                 // :63
  std::vector<StateTuple> tuples;       // :64
  std::vector<double> dist;             // :65
  std::vector<BackPtr> back;            // :66
  std::vector<uint8_t> settled;         // :67
  std::priority_queue<QueueItem, std::vector<QueueItem>, QueueCompare> queue;  // :68
  SearchStats& st = res.stats;

  auto get_or_create = [&](StateTuple t) -> uint32_t {  // :70-89
    uint64_t k = TupleMap::pack(t.s1, t.s2, t.filter);
    uint32_t id;
    if (tuple_to_id.get(k, &id)) return id;
    id = (uint32_t)tuples.size();
    tuples.push_back(t);
    dist.push_back(kInf);
    back.push_back(BackPtr{false, 0, 0, 0, 0.0});
    settled.push_back(0);
    tuple_to_id.put(k, id);
    st.tuples++;
    return id;
  };

  auto relax = [&](uint32_t curr_id, StateTuple next_tuple, Label ilabel, Label olabel, double edge_weight) {  // :91-144
    st.relax_calls++;
    uint32_t next_id = get_or_create(next_tuple);
    double new_dist = w_times(dist[curr_id], edge_weight);
    double old_dist = dist[next_id];
    int by_dist = w_compare(new_dist, old_dist);
    bool take = false;
    if (w_is_zero(old_dist) || by_dist < 0) {
      take = true;
    } else if (by_dist == 0) {
      const BackPtr& bp = back[next_id];
      if (bp.has) {
        if (curr_id < bp.prev_id ||
            (curr_id == bp.prev_id && (ilabel < bp.ilabel || (ilabel == bp.ilabel && olabel < bp.olabel)))) {
          take = true;
          st.retakes++;
        }
      } else {
        take = true;
      }
    }
    if (!take) return;
    dist[next_id] = new_dist;
    back[next_id] = BackPtr{true, curr_id, ilabel, olabel, edge_weight};
    if (!settled[next_id]) { queue.push(QueueItem{next_id, new_dist}); st.pushes++; }
  };

  // :146-153
  uint32_t init_id = get_or_create(StateTuple{fst1.start(), fst2.start(), 0});
  dist[init_id] = 0.0;
  queue.push(QueueItem{init_id, 0.0}); st.pushes++;

  bool have_best = false; uint32_t best_final_id = 0;  // :155-157
  double best_final_weight = kInf, best_total = kInf;

  while (!queue.empty()) {  // :159
    QueueItem item = queue.top(); queue.pop();
    uint32_t curr_id = item.tuple_id;
    if (settled[curr_id]) continue;                              // :161
    if (w_compare(item.dist, dist[curr_id]) != 0) continue;      // :162
    settled[curr_id] = 1;                                        // :163
    st.pops++;

    StateTuple t = tuples[curr_id];                              // :165
    double fw1 = fst1.final_weight(t.s1), fw2 = fst2.final_weight(t.s2);
    if (!w_is_zero(fw1) && !w_is_zero(fw2)) {                    // :168-179
      double final_w = w_times(fw1, fw2);
      double total = w_times(dist[curr_id], final_w);
      if (!have_best || w_compare(total, best_total) < 0 ||
          (w_compare(total, best_total) == 0 && curr_id < best_final_id)) {
        have_best = true; best_final_id = curr_id; best_final_weight = final_w; best_total = total;
      }
    }

    // NOTE: relax() may grow `tuples`; the lhs/rhs arc containers are stable.
    // :182-202 non-epsilon matches
    for (const Arc& a1 : fst1.arcs(t.s1)) {
      if (a1.olabel == kEpsilon) continue;
      for (const PackedArc& a2 : fst2.arcs_by_ilabel(t.s2, a1.olabel))
        relax(curr_id, StateTuple{a1.nextstate, a2.nextstate, 0}, a1.ilabel, a2.olabel, w_times(a1.weight, a2.weight));
    }
    // :227-252 lhs consumes an output epsilon
    if (t.filter != 1) {
      for (const Arc& a1 : fst1.arcs(t.s1)) {
        if (a1.olabel != kEpsilon) continue;
        uint8_t nf = t.filter == 0 ? 2 : t.filter;
        relax(curr_id, StateTuple{a1.nextstate, t.s2, nf}, a1.ilabel, kEpsilon, a1.weight);
      }
    }
    // :254-278 rhs consumes an input epsilon
    if (t.filter != 2) {
      for (const PackedArc& a2 : fst2.arcs_by_ilabel(t.s2, kEpsilon)) {
        uint8_t nf = t.filter == 0 ? 1 : t.filter;
        relax(curr_id, StateTuple{t.s1, a2.nextstate, nf}, kEpsilon, a2.olabel, a2.weight);
      }
    }
    // :307-336 both consume epsilon
    if (t.filter == 0) {
      ArcSpan rhs_eps = fst2.arcs_by_ilabel(t.s2, kEpsilon);
      if (rhs_eps.size() > 0) {
        for (const Arc& a1 : fst1.arcs(t.s1)) {
          if (a1.olabel != kEpsilon) continue;
          for (const PackedArc& a2 : rhs_eps)
            relax(curr_id, StateTuple{a1.nextstate, a2.nextstate, 0}, a1.ilabel, a2.olabel, w_times(a1.weight, a2.weight));
        }
      }
    }
  }

  if (!have_best) { res.status = Status::kEmpty; return res; }  // :368-370

  // :372-380 back-track.  Hazard H1: a cyclic back[] chain makes the reference
  // append until OOM; we bound the walk at N steps and flag it.
  std::vector<BackPtr> reverse;
  uint32_t curr = best_final_id;
  while (curr != init_id) {
    const BackPtr& bp = back[curr];
    if (!bp.has) { res.status = Status::kEmpty; return res; }   // :375-377
    reverse.push_back(bp);
    curr = bp.prev_id;
    if (reverse.size() > tuples.size()) { res.status = Status::kBacktrackCycle; return res; }
  }
  // :382-398
  res.status = Status::kOk;
  res.final_weight = best_final_weight;
  res.arcs.reserve(reverse.size());
  for (size_t i = reverse.size(); i > 0; i--) {
    const BackPtr& bp = reverse[i - 1];
    res.arcs.push_back(PathArc{bp.ilabel, bp.olabel, bp.weight});
  }
  res.final_s1 = tuples[best_final_id].s1; res.final_s2 = tuples[best_final_id].s2; res.final_filter = tuples[best_final_id].filter;
  return res;
}

// ── Eager compose: src/ops/compose.zig:29-198 (frozen rhs variant) ──
template <class Lhs>
MutableFst compose(const Lhs& fst1, const Fst& fst2) {
  MutableFst result;
  if (fst1.start() == kNoState || fst2.start() == kNoState) return result;  // :33-35
  struct StateTuple { StateId s1, s2; uint8_t filter; };
  TupleMap state_map;
  std::vector<StateTuple> queue;
  StateId init_state = result.add_state();   // :57-61
  result.set_start(init_state);
  state_map.put(TupleMap::pack(fst1.start(), fst2.start(), 0), init_state);
  queue.push_back(StateTuple{fst1.start(), fst2.start(), 0});
  auto get_or_create = [&](StateTuple t) -> StateId {  // :77-91
    uint64_t k = TupleMap::pack(t.s1, t.s2, t.filter);
    uint32_t id;
    if (state_map.get(k, &id)) return id;
    StateId ns = result.add_state();
    state_map.put(k, ns);
    queue.push_back(t);
    return ns;
  };
  for (size_t qi = 0; qi < queue.size(); qi++) {  // :64
    StateTuple t = queue[qi];
    uint32_t current = 0; state_map.get(TupleMap::pack(t.s1, t.s2, t.filter), &current);
    double fw1 = fst1.final_weight(t.s1), fw2 = fst2.final_weight(t.s2);   // :69-74
    if (!w_is_zero(fw1) && !w_is_zero(fw2)) result.set_final(current, w_times(fw1, fw2));
    for (const Arc& a1 : fst1.arcs(t.s1)) {  // :95-109
      if (a1.olabel == kEpsilon) continue;
      for (const PackedArc& a2 : fst2.arcs_by_ilabel(t.s2, a1.olabel)) {
        StateId ns = get_or_create(StateTuple{a1.nextstate, a2.nextstate, 0});
        result.add_arc(current, Arc{a1.ilabel, a2.olabel, w_times(a1.weight, a2.weight), ns});
      }
    }
    if (t.filter != 1) {  // :126-136
      for (const Arc& a1 : fst1.arcs(t.s1)) {
        if (a1.olabel != kEpsilon) continue;
        uint8_t nf = t.filter == 0 ? 2 : t.filter;
        StateId ns = get_or_create(StateTuple{a1.nextstate, t.s2, nf});
        result.add_arc(current, Arc{a1.ilabel, kEpsilon, a1.weight, ns});
      }
    }
    if (t.filter != 2) {  // :138-149
      for (const PackedArc& a2 : fst2.arcs_by_ilabel(t.s2, kEpsilon)) {
        uint8_t nf = t.filter == 0 ? 1 : t.filter;
        StateId ns = get_or_create(StateTuple{t.s1, a2.nextstate, nf});
        result.add_arc(current, Arc{kEpsilon, a2.olabel, a2.weight, ns});
      }
    }
    if (t.filter == 0) {  // :162-178
      ArcSpan rhs_eps = fst2.arcs_by_ilabel(t.s2, kEpsilon);
      if (rhs_eps.size() > 0) {
        for (const Arc& a1 : fst1.arcs(t.s1)) {
          if (a1.olabel != kEpsilon) continue;
          for (const PackedArc& a2 : rhs_eps) {
            StateId ns = get_or_create(StateTuple{a1.nextstate, a2.nextstate, 0});
            result.add_arc(current, Arc{a1.ilabel, a2.olabel, w_times(a1.weight, a2.weight), ns});
          }
        }
      }
    }
  }
  return result;
}

// ── Explicit-graph shortest path: src/ops/shortest-path.zig:18-139 ──
inline PathResult shortest_path(const MutableFst& fst, uint32_t n) {
  PathResult res;
  if (fst.start() == kNoState || n == 0) { res.status = Status::kEmpty; return res; }  // :21-23
  if (n != 1) { res.status = Status::kUnsupportedN; return res; }                      // :24
  size_t ns = fst.num_states();
  std::vector<double> dist(ns, kInf);                 // :35-37
  dist[fst.start()] = 0.0;
  struct BackPtr { bool has; StateId prev_state; uint32_t arc_idx; };   // :40-45
  std::vector<BackPtr> back(ns, BackPtr{false, 0, 0});
  std::vector<uint8_t> settled(ns, 0);
  struct QueueItem { StateId state; double dist; };
  struct QueueCompare {                               // :54-60
    bool operator()(const QueueItem& a, const QueueItem& b) const {
      int c = w_compare(a.dist, b.dist);
      if (c != 0) return c > 0;
      return a.state > b.state;
    }
  };
  std::priority_queue<QueueItem, std::vector<QueueItem>, QueueCompare> queue;
  queue.push(QueueItem{fst.start(), 0.0});
  while (!queue.empty()) {                            // :64-86
    QueueItem item = queue.top(); queue.pop();
    StateId s = item.state;
    if (settled[s]) continue;
    if (w_compare(item.dist, dist[s]) != 0) continue;
    settled[s] = 1;
    res.stats.pops++;
    const auto& sa = fst.arcs(s);
    for (size_t ai = 0; ai < sa.size(); ai++) {
      const Arc& a = sa[ai];
      res.stats.relax_calls++;
      StateId next = a.nextstate;
      double new_dist = w_times(dist[s], a.weight);
      double old_dist = dist[next];
      int by_dist = w_compare(new_dist, old_dist);
      StateId prev_state = back[next].has ? back[next].prev_state : kNoState;
      bool better_tie = by_dist == 0 && (prev_state == kNoState || s < prev_state);   // :75-78
      if (w_is_zero(old_dist) || by_dist < 0 || better_tie) {
        dist[next] = new_dist;
        back[next] = BackPtr{true, s, (uint32_t)ai};
        if (!settled[next]) queue.push(QueueItem{next, new_dist});
      }
    }
  }
  StateId best_final = kNoState; double best_total = kInf;   // :88-104
  for (size_t i = 0; i < ns; i++) {
    StateId s = (StateId)i;
    if (w_is_zero(dist[s])) continue;
    double fw = fst.final_weight(s);
    if (w_is_zero(fw)) continue;
    double total = w_times(dist[s], fw);
    if (best_final == kNoState || w_compare(total, best_total) < 0 ||
        (w_compare(total, best_total) == 0 && s < best_final)) { best_final = s; best_total = total; }
  }
  if (best_final == kNoState) { res.status = Status::kEmpty; return res; }
  std::vector<BackPtr> rev;                                  // :113-118
  StateId current = best_final;
  while (back[current].has) {
    rev.push_back(back[current]);
    current = back[current].prev_state;
    if (rev.size() > ns + 1) { res.status = Status::kBacktrackCycle; return res; }  // hazard H1 analogue
  }
  if (current != fst.start()) { res.status = Status::kEmpty; return res; }          // :121-123
  res.status = Status::kOk;
  res.final_weight = fst.final_weight(best_final);
  for (size_t i = rev.size(); i > 0; i--) {                  // :129-136
    const Arc& a = fst.arcs(rev[i - 1].prev_state)[rev[i - 1].arc_idx];
    res.arcs.push_back(PathArc{a.ilabel, a.olabel, a.weight});
  }
  res.stats.tuples = ns;
  return res;
}

// ── Bench generators: bench/optimize-bench.zig ──
inline MutableFst gen_repeated_label_acceptor(size_t len, Label label) {   // :182-196
  MutableFst f; f.add_states(len + 1); f.set_start(0); f.set_final((StateId)len, 0.0);
  for (size_t i = 0; i < len; i++) f.add_arc((StateId)i, Arc{label, label, 0.0, (StateId)(i + 1)});
  return f;
}
inline MutableFst gen_linear_acceptor_alphabet(size_t len, size_t alphabet) {   // :168-180
  MutableFst f; f.add_states(len + 1); f.set_start(0); f.set_final((StateId)len, 0.0);
  size_t alpha = std::max<size_t>(1, alphabet);
  for (size_t i = 0; i < len; i++) { Label l = (Label)((i % alpha) + 1); f.add_arc((StateId)i, Arc{l, l, 0.0, (StateId)(i + 1)}); }
  return f;
}
inline MutableFst gen_plain_transducer(size_t T, size_t B) {   // :290-305 (transducer_for_freeze)
  MutableFst f; f.add_states(T); f.set_start(0);
  for (size_t i = 0; i < T; i++) {
    f.set_final((StateId)i, 0.0);
    for (size_t b = 0; b < B; b++)
      f.add_arc((StateId)i, Arc{(Label)((b % 255) + 1), (Label)(((i + b) % 255) + 1), (double)b, (StateId)((i + b + 1) % T)});
  }
  return f;
}
inline MutableFst gen_epsilon_dense_transducer(size_t len, size_t B) {   // :219-248
  MutableFst f; f.add_states(len + 1); f.set_start(0);
  for (size_t i = 0; i <= len; i++) f.set_final((StateId)i, 0.0);
  for (size_t i = 0; i < len; i++) {
    f.add_arc((StateId)i, Arc{0, 0, 0.0, (StateId)(i + 1)});
    for (size_t b = 0; b < B; b++) {
      size_t jump = (b % 4) + 1;
      size_t next = std::min(i + jump, len);
      f.add_arc((StateId)i, Arc{1, (Label)(((i + b) % 255) + 1), (double)b, (StateId)next});
    }
  }
  return f;
}
inline MutableFst gen_ambiguous_chain_transducer(size_t len, size_t B) {   // :250-277
  MutableFst f; f.add_states(len + 1); f.set_start(0);
  for (size_t i = 0; i <= len; i++) f.set_final((StateId)i, 0.0);
  for (size_t i = 0; i <= len; i++) {
    f.add_arc((StateId)i, Arc{1, 1, 0.0, (StateId)i});
    size_t fanout = std::max<size_t>(1, std::min<size_t>(B, 4));
    for (size_t b = 0; b < fanout; b++) {
      size_t next = std::min(i + b + 1, len);
      f.add_arc((StateId)i, Arc{1, (Label)(((i + b) % 255) + 1), (double)b, (StateId)next});
    }
  }
  return f;
}

}  // namespace orc
