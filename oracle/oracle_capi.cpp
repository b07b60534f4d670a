// CPU ORACLE — TEST INFRASTRUCTURE ONLY (see fst_oracle.hpp header).
// Plain C ABI over the restatement so tests/ and bench.py can drive it via ctypes.
#include "fst_oracle.hpp"

#include <atomic>
#include <chrono>
#include <thread>

using namespace orc;

extern "C" {

struct OrcInfo {
  int32_t status;        // orc::Status
  uint32_t n_arcs;       // path length (may exceed cap: then arrays hold the first cap arcs)
  double final_weight;
  double total;          // left-to-right sum of arc weights + final
  uint64_t tuples, relax_calls, pushes, retakes, pops;
  uint32_t final_s1, final_s2, final_filter, _pad;
};

// ── mutable ──
void* orc_mutable_new() { return new MutableFst(); }
void orc_mutable_free(void* m) { delete (MutableFst*)m; }
uint32_t orc_mutable_add_state(void* m) { return ((MutableFst*)m)->add_state(); }
void orc_mutable_set_start(void* m, uint32_t s) { ((MutableFst*)m)->set_start(s); }
void orc_mutable_set_final(void* m, uint32_t s, double w) { ((MutableFst*)m)->set_final(s, w); }
void orc_mutable_add_arc(void* m, uint32_t src, uint32_t il, uint32_t ol, double w, uint32_t next) {
  ((MutableFst*)m)->add_arc(src, Arc{il, ol, w, next});
}
// Bulk build: arcs are appended to their source state in the order given.
void* orc_mutable_from_arrays(uint32_t num_states, uint32_t start, const double* finals, uint32_t num_arcs,
                              const uint32_t* src, const uint32_t* il, const uint32_t* ol, const double* w,
                              const uint32_t* next) {
  MutableFst* m = new MutableFst();
  m->add_states(num_states);
  m->set_start(start);
  for (uint32_t i = 0; i < num_states; i++) m->set_final(i, finals[i]);
  for (uint32_t i = 0; i < num_arcs; i++) m->add_arc(src[i], Arc{il[i], ol[i], w[i], next[i]});
  return m;
}
void* orc_compile_string(const uint8_t* s, uint32_t len) { return new MutableFst(compile_string(s, len)); }
void* orc_compile_string_transducer(const uint8_t* a, uint32_t alen, const uint8_t* b, uint32_t blen) {
  return new MutableFst(compile_string_transducer(a, alen, b, blen));
}
uint32_t orc_mutable_num_states(void* m) { return (uint32_t)((MutableFst*)m)->num_states(); }

// ── frozen ──
// c-api.zig:507-526: fst_freeze clones the mutable, then Fst.fromMutable sorts the clone.
void* orc_freeze(void* m) {
  MutableFst copy = *(MutableFst*)m;
  return new Fst(Fst::from_mutable(copy));
}
void* orc_fst_from_bytes(const uint8_t* data, uint64_t len) {
  Fst f;
  if (!Fst::from_bytes(data, (size_t)len, &f)) return nullptr;
  return new Fst(std::move(f));
}
uint64_t orc_fst_num_bytes(void* f) { return ((Fst*)f)->bytes.size(); }
void orc_fst_copy_bytes(void* f, uint8_t* out) { std::memcpy(out, ((Fst*)f)->bytes.data(), ((Fst*)f)->bytes.size()); }
void orc_fst_free(void* f) { delete (Fst*)f; }
uint32_t orc_fst_num_states(void* f) { return ((Fst*)f)->num_states(); }
uint32_t orc_fst_num_arcs_total(void* f) { return ((Fst*)f)->header().num_arcs; }

// kind: 0 plain (bench transducer_frozen), 1 epsilon-dense, 2 ambiguous chain
void* orc_gen_frozen(int kind, uint32_t T, uint32_t B) {
  MutableFst m = kind == 0 ? gen_plain_transducer(T, B)
                 : kind == 1 ? gen_epsilon_dense_transducer(T, B)
                             : gen_ambiguous_chain_transducer(T, B);
  return new Fst(Fst::from_mutable(m));
}

static void fill(const PathResult& r, uint32_t cap, uint32_t* il, uint32_t* ol, double* w, OrcInfo* info) {
  info->status = (int32_t)r.status;
  info->n_arcs = (uint32_t)r.arcs.size();
  info->final_weight = r.final_weight;
  info->total = r.status == Status::kOk ? r.total() : kInf;
  info->tuples = r.stats.tuples; info->relax_calls = r.stats.relax_calls; info->pushes = r.stats.pushes;
  info->retakes = r.stats.retakes; info->pops = r.stats.pops;
  info->final_s1 = r.final_s1; info->final_s2 = r.final_s2; info->final_filter = r.final_filter; info->_pad = 0;
  uint32_t n = std::min<uint32_t>(cap, info->n_arcs);
  for (uint32_t i = 0; i < n; i++) { if (il) il[i] = r.arcs[i].ilabel; if (ol) ol[i] = r.arcs[i].olabel; if (w) w[i] = r.arcs[i].weight; }
}

// Lazy path, arbitrary mutable lhs (the fst_compose_frozen_shortest_path contract).
void orc_csp_mutable(void* lhs, void* fst, uint32_t n, uint32_t cap, uint32_t* il, uint32_t* ol, double* w, OrcInfo* info) {
  MutableLhs l{(MutableFst*)lhs};
  fill(compose_shortest_path(l, *(Fst*)fst, n), cap, il, ol, w, info);
}
// Lazy path, lhs = compile_string(bytes).
void orc_csp_bytes(void* fst, const uint8_t* s, uint32_t len, uint32_t cap, uint32_t* il, uint32_t* ol, double* w, OrcInfo* info) {
  MutableFst m = compile_string(s, len);
  MutableLhs l{&m};
  fill(compose_shortest_path(l, *(Fst*)fst, 1), cap, il, ol, w, info);
}
// Eager pair: compose (compose.zig) then shortestPath (shortest-path.zig).
void orc_eager_mutable(void* lhs, void* fst, uint32_t n, uint32_t cap, uint32_t* il, uint32_t* ol, double* w, OrcInfo* info,
                       uint64_t* lattice_states, uint64_t* lattice_arcs) {
  MutableLhs l{(MutableFst*)lhs};
  MutableFst lat = compose(l, *(Fst*)fst);
  if (lattice_states) *lattice_states = lat.num_states();
  if (lattice_arcs) *lattice_arcs = lat.total_arcs();
  fill(shortest_path(lat, n), cap, il, ol, w, info);
}

// Batched lazy path over byte strings with a thread pool: the CPU baseline.
// Outputs: per-string status/len/final/total and flat path arrays with a fixed
// per-string capacity `cap` (arrays may be null to time only).
// Returns wall seconds of the search region.
static double batch_bytes_impl(void* fst, const uint8_t* bytes, const uint64_t* offsets, uint32_t n_strings,
                               uint32_t n_threads, uint32_t cap, uint32_t* il, uint32_t* ol, double* w,
                               int32_t* status, uint32_t* lens, double* finals, double* totals,
                               uint64_t* sum_tuples, uint64_t* sum_relax, bool eager) {
  std::atomic<uint32_t> next{0};
  std::atomic<uint64_t> st{0}, sr{0};
  const Fst& f = *(Fst*)fst;
  auto worker = [&]() {
    uint64_t lt = 0, lr = 0;
    for (;;) {
      uint32_t i = next.fetch_add(1);
      if (i >= n_strings) break;
      MutableFst m = compile_string(bytes + offsets[i], (size_t)(offsets[i + 1] - offsets[i]));
      MutableLhs l{&m};
      PathResult r;
      if (eager) {   // compose (compose.zig) then shortestPath (shortest-path.zig); work = lattice size
        MutableFst lat = compose(l, f);
        r = shortest_path(lat, 1);
        r.stats.tuples = lat.num_states(); r.stats.relax_calls = lat.total_arcs();
      } else {
        r = compose_shortest_path(l, f, 1);
      }
      lt += r.stats.tuples; lr += r.stats.relax_calls;
      if (status) status[i] = (int32_t)r.status;
      if (lens) lens[i] = (uint32_t)r.arcs.size();
      if (finals) finals[i] = r.final_weight;
      if (totals) totals[i] = r.status == Status::kOk ? r.total() : kInf;
      uint32_t n = std::min<uint32_t>(cap, (uint32_t)r.arcs.size());
      for (uint32_t k = 0; k < n; k++) {
        size_t o = (size_t)i * cap + k;
        if (il) il[o] = r.arcs[k].ilabel;
        if (ol) ol[o] = r.arcs[k].olabel;
        if (w) w[o] = r.arcs[k].weight;
      }
    }
    st += lt; sr += lr;
  };
  auto t0 = std::chrono::steady_clock::now();
  if (n_threads <= 1) worker();
  else {
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < n_threads; t++) th.emplace_back(worker);
    for (auto& t : th) t.join();
  }
  auto t1 = std::chrono::steady_clock::now();
  if (sum_tuples) *sum_tuples = st.load();
  if (sum_relax) *sum_relax = sr.load();
  return std::chrono::duration<double>(t1 - t0).count();
}

double orc_csp_batch_bytes(void* fst, const uint8_t* bytes, const uint64_t* offsets, uint32_t n_strings,
                           uint32_t n_threads, uint32_t cap, uint32_t* il, uint32_t* ol, double* w,
                           int32_t* status, uint32_t* lens, double* finals, double* totals,
                           uint64_t* sum_tuples, uint64_t* sum_relax) {
  return batch_bytes_impl(fst, bytes, offsets, n_strings, n_threads, cap, il, ol, w, status, lens, finals, totals, sum_tuples, sum_relax, false);
}
// Same driver for the eager pair (BASELINE config 5).
double orc_eager_batch_bytes(void* fst, const uint8_t* bytes, const uint64_t* offsets, uint32_t n_strings,
                             uint32_t n_threads, uint32_t cap, uint32_t* il, uint32_t* ol, double* w,
                             int32_t* status, uint32_t* lens, double* finals, double* totals,
                             uint64_t* sum_tuples, uint64_t* sum_relax) {
  return batch_bytes_impl(fst, bytes, offsets, n_strings, n_threads, cap, il, ol, w, status, lens, finals, totals, sum_tuples, sum_relax, true);
}

// Eager lattice of compile_string(s) o fst (compose.zig:29-198) as a new mutable; read it with the dump calls below.
void* orc_compose_bytes(void* fst, const uint8_t* s, uint32_t len) {
  MutableFst m = compile_string(s, len);
  MutableLhs l{&m};
  return new MutableFst(compose(l, *(Fst*)fst));
}
uint64_t orc_mutable_total_arcs(void* m) { return ((MutableFst*)m)->total_arcs(); }
uint32_t orc_mutable_start(void* m) { return ((MutableFst*)m)->start(); }
// CSR dump: arc_begin[num_states + 1], finals[num_states], arcs in stored order.
void orc_mutable_dump(void* mp, uint64_t* arc_begin, double* finals, uint32_t* il, uint32_t* ol, double* w, uint32_t* next) {
  const MutableFst& m = *(MutableFst*)mp;
  uint64_t k = 0;
  for (size_t s = 0; s < m.num_states(); s++) {
    arc_begin[s] = k; finals[s] = m.final_weight((StateId)s);
    for (const Arc& a : m.arcs((StateId)s)) { il[k] = a.ilabel; ol[k] = a.olabel; w[k] = a.weight; next[k] = a.nextstate; k++; }
  }
  arc_begin[m.num_states()] = k;
}

// string.zig:64-97 over a path result expressed as label arrays is trivial; the
// mutable variant is exposed for the fst_print_* parity tests.
int32_t orc_print_string(void* m, int output_tape, uint8_t* buf, uint32_t buf_len) {
  std::string s;
  if (!print_string_from_tape(*(MutableFst*)m, output_tape != 0, &s)) return -1;
  if (s.size() > buf_len) return -1;
  if (buf) std::memcpy(buf, s.data(), s.size());
  return (int32_t)s.size();
}

}  // extern "C"
