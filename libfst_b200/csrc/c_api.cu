// C ABI of libfst_b200.so (declared in include/libfst_b200.h).
//
// The drop-in subset mirrors the reference's src/c-api.zig conventions:
// generation-tagged handles in two separate tables (:109-277), one global mutex
// for table bookkeeping only (:279-282), left operand snapshotted under the lock
// (:754), right operand pinned for the duration of a search (:759, :210-248).
// The search itself runs on the GPU; there is no CPU fallback.
#include "../../include/libfst_b200.h"

#include <atomic>
#include <chrono>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "engine.cuh"

using namespace fstb200;

namespace {

// ── generation-tagged handle table (reference src/c-api.zig:109-277) ──
template <class T>
class HandleTable {
 public:
  uint64_t insert(T* ptr) {
    uint32_t idx;
    if (!free_.empty()) {
      idx = free_.back(); free_.pop_back();
      gen_[idx]++; if (gen_[idx] == 0) gen_[idx] = 1;
      slot_[idx] = ptr; pins_[idx] = 0; pending_[idx] = false;
    } else {
      idx = (uint32_t)slot_.size();
      slot_.push_back(ptr); gen_.push_back(1); pins_.push_back(0); pending_.push_back(false);
    }
    return ((uint64_t)gen_[idx] << 32) | idx;
  }
  T* get(uint64_t h) const { uint32_t i; return validate(h, &i) ? slot_[i] : nullptr; }
  T* pin(uint64_t h) {
    uint32_t i;
    if (!validate(h, &i) || !slot_[i]) return nullptr;
    pins_[i]++;
    return slot_[i];
  }
  // returns the object to destroy (caller destroys outside any per-object use) or null
  T* unpin(uint64_t h) {
    if (h == FST_INVALID_HANDLE) return nullptr;
    uint32_t i = (uint32_t)(h & 0xFFFFFFFFu);
    if (i >= slot_.size() || pins_[i] == 0) return nullptr;
    if (--pins_[i] != 0 || !pending_[i]) return nullptr;
    T* p = slot_[i]; slot_[i] = nullptr; pending_[i] = false; free_.push_back(i);
    return p;
  }
  // remove: returns the object to destroy now, or null if deferred / invalid
  T* remove(uint64_t h) {
    uint32_t i;
    if (!validate(h, &i) || !slot_[i]) return nullptr;
    if (pins_[i] > 0) { pending_[i] = true; bump(i); return nullptr; }
    T* p = slot_[i]; slot_[i] = nullptr; bump(i); free_.push_back(i);
    return p;
  }
  std::vector<T*> drain() {
    std::vector<T*> all;
    for (T* p : slot_) if (p) all.push_back(p);
    slot_.clear(); gen_.clear(); pins_.clear(); pending_.clear(); free_.clear();
    return all;
  }
 private:
  bool validate(uint64_t h, uint32_t* idx) const {
    if (h == FST_INVALID_HANDLE) return false;
    uint32_t g = (uint32_t)(h >> 32), i = (uint32_t)(h & 0xFFFFFFFFu);
    if (g == 0 || i == 0xFFFFFFFFu || i >= slot_.size()) return false;
    if (pending_[i] || gen_[i] != g) return false;
    *idx = i;
    return true;
  }
  void bump(uint32_t i) { gen_[i]++; if (gen_[i] == 0) gen_[i] = 1; }
  std::vector<T*> slot_; std::vector<uint32_t> gen_, pins_; std::vector<bool> pending_; std::vector<uint32_t> free_;
};

// A frozen transducer plus its lazily created device images (one per device).
struct FrozenEntry {
  std::unique_ptr<HostFrozen> host;
  std::mutex dev_mu;
  std::map<int, DeviceFst*> images;
  ~FrozenEntry() { for (auto& kv : images) free_device_fst(kv.second); }
  DeviceFst* image_for(int device, cudaError_t* err) {
    std::lock_guard<std::mutex> lk(dev_mu);
    auto it = images.find(device);
    if (it != images.end()) { *err = cudaSuccess; return it->second; }
    DeviceFst* d = nullptr;
    *err = upload_fst(*host, device, &d);
    if (*err != cudaSuccess) return nullptr;
    images[device] = d;
    return d;
  }
};

std::mutex g_mu;                       // reference api_mutex (src/c-api.zig:282)
HandleTable<HostMutable> g_mutables;   // src/c-api.zig:276
HandleTable<FrozenEntry> g_frozen;     // src/c-api.zig:277

thread_local BatchCounters t_last;

bool trace_enabled() {                 // src/c-api.zig:55-64
  static int v = -1;
  if (v < 0) v = std::getenv("LIBFST_TRACE_COMPOSE") != nullptr;
  return v == 1;
}
void trace_line(const char* tag, uint64_t a, uint64_t b, size_t in_states, size_t in_arcs, size_t out_states,
                size_t out_arcs, double us) {   // src/c-api.zig:74-103 (same line format)
  if (!trace_enabled()) return;
  std::fprintf(stderr,
               "[libfst] %s op=fst_compose_frozen a=%llu b=%llu in_states=%zu in_arcs=%zu out_states=%zu out_arcs=%zu elapsed_us=%lld\n",
               tag, (unsigned long long)a, (unsigned long long)b, in_states, in_arcs, out_states, out_arcs, (long long)us);
}

uint64_t new_mutable(HostMutable&& m) {
  auto* p = new (std::nothrow) HostMutable(std::move(m));
  if (!p) return FST_INVALID_HANDLE;
  std::lock_guard<std::mutex> lk(g_mu);
  return g_mutables.insert(p);
}
uint64_t new_frozen(std::unique_ptr<HostFrozen> f) {
  if (!f) return FST_INVALID_HANDLE;
  auto* e = new (std::nothrow) FrozenEntry();
  if (!e) return FST_INVALID_HANDLE;
  e->host = std::move(f);
  std::lock_guard<std::mutex> lk(g_mu);
  return g_frozen.insert(e);
}

bool device_available() {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess && n > 0;
}

struct PinGuard {   // unpin on scope exit (src/c-api.zig:777-781)
  uint64_t h;
  ~PinGuard() {
    FrozenEntry* dead;
    { std::lock_guard<std::mutex> lk(g_mu); dead = g_frozen.unpin(h); }
    delete dead;
  }
};

struct BatchResultImpl {
  FstB200BatchResult pub;
  void* pinned = nullptr;   // one pinned allocation backing all arrays
  size_t pinned_bytes = 0;
};

// Pinned result buffers are recycled: page-locking 16 MB per call is milliseconds of driver work.
static std::mutex g_pin_mu;
static std::vector<std::pair<void*, size_t>> g_pin_free;
static void* pinned_alloc(size_t bytes, size_t* got) {
  {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    int best = -1;
    for (int i = 0; i < (int)g_pin_free.size(); i++)
      if (g_pin_free[i].second >= bytes && (best < 0 || g_pin_free[i].second < g_pin_free[best].second)) best = i;
    if (best >= 0) {
      void* p = g_pin_free[best].first; *got = g_pin_free[best].second;
      g_pin_free.erase(g_pin_free.begin() + best);
      return p;
    }
  }
  void* p = nullptr;
  const size_t want = bytes + bytes / 8 + 4096;
  if (cudaMallocHost(&p, want) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  *got = want;
  return p;
}
static void pinned_release(void* p, size_t bytes) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    size_t held = 0;
    for (auto& e : g_pin_free) held += e.second;
    // one buffer per chunk of a multi-GPU call is recycled; the pool is bounded in entries and in bytes
    if (g_pin_free.size() < 64 && held + bytes <= (48ull << 30)) { g_pin_free.emplace_back(p, bytes); return; }
  }
  cudaFreeHost(p);
}
static void pinned_drain() {
  std::lock_guard<std::mutex> lk(g_pin_mu);
  for (auto& e : g_pin_free) cudaFreeHost(e.first);
  g_pin_free.clear();
}

}  // namespace

extern "C" {

// ── mutable lifecycle (src/c-api.zig:436-505) ──
FstMutableHandle fst_mutable_new(void) { return new_mutable(HostMutable()); }

FstMutableHandle fst_mutable_clone(FstMutableHandle handle) {
  HostMutable copy;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    HostMutable* m = g_mutables.get(handle);
    if (!m) return FST_INVALID_HANDLE;
    copy = *m;
  }
  return new_mutable(std::move(copy));
}

void fst_mutable_free(FstMutableHandle handle) {
  HostMutable* dead;
  { std::lock_guard<std::mutex> lk(g_mu); dead = g_mutables.remove(handle); }
  delete dead;
}

uint32_t fst_mutable_add_state(FstMutableHandle handle) {
  std::lock_guard<std::mutex> lk(g_mu);
  HostMutable* m = g_mutables.get(handle);
  if (!m) return FST_NO_STATE;
  return m->add_state();
}

FstError fst_mutable_set_start(FstMutableHandle handle, uint32_t state) {
  std::lock_guard<std::mutex> lk(g_mu);
  HostMutable* m = g_mutables.get(handle);
  if (!m) return FST_INVALID_ARG;
  if (state >= m->num_states()) return FST_INVALID_STATE;
  m->start = state;
  return FST_OK;
}

FstError fst_mutable_set_final(FstMutableHandle handle, uint32_t state, double weight) {
  std::lock_guard<std::mutex> lk(g_mu);
  HostMutable* m = g_mutables.get(handle);
  if (!m) return FST_INVALID_ARG;
  if (state >= m->num_states()) return FST_INVALID_STATE;
  m->finals[state] = weight;
  return FST_OK;
}

FstError fst_mutable_add_arc(FstMutableHandle handle, uint32_t src, uint32_t ilabel, uint32_t olabel, double weight,
                             uint32_t nextstate) {
  std::lock_guard<std::mutex> lk(g_mu);
  HostMutable* m = g_mutables.get(handle);
  if (!m) return FST_INVALID_ARG;
  if (src >= m->num_states() || nextstate >= m->num_states()) return FST_INVALID_STATE;
  m->arcs[src].push_back(HostArc{ilabel, olabel, weight, nextstate});
  return FST_OK;
}

// Bulk builders (new; same effect as the per-call builders above, one lock).
FstError fst_b200_mutable_add_states(FstMutableHandle handle, uint32_t n) {
  std::lock_guard<std::mutex> lk(g_mu);
  HostMutable* m = g_mutables.get(handle);
  if (!m) return FST_INVALID_ARG;
  m->add_states(n);
  return FST_OK;
}
FstError fst_b200_mutable_set_finals(FstMutableHandle handle, uint32_t n, const uint32_t* states, const double* weights) {
  std::lock_guard<std::mutex> lk(g_mu);
  HostMutable* m = g_mutables.get(handle);
  if (!m || (n && (!states || !weights))) return FST_INVALID_ARG;
  for (uint32_t i = 0; i < n; i++) if (states[i] >= m->num_states()) return FST_INVALID_STATE;
  for (uint32_t i = 0; i < n; i++) m->finals[states[i]] = weights[i];
  return FST_OK;
}
FstError fst_b200_mutable_add_arcs(FstMutableHandle handle, uint64_t n, const uint32_t* src, const uint32_t* ilabel,
                                   const uint32_t* olabel, const double* weight, const uint32_t* nextstate) {
  std::lock_guard<std::mutex> lk(g_mu);
  HostMutable* m = g_mutables.get(handle);
  if (!m || (n && (!src || !ilabel || !olabel || !weight || !nextstate))) return FST_INVALID_ARG;
  for (uint64_t i = 0; i < n; i++) if (src[i] >= m->num_states() || nextstate[i] >= m->num_states()) return FST_INVALID_STATE;
  for (uint64_t i = 0; i < n; i++) m->arcs[src[i]].push_back(HostArc{ilabel[i], olabel[i], weight[i], nextstate[i]});
  return FST_OK;
}

// ── mutable queries (src/c-api.zig:1376-1424) ──
uint32_t fst_mutable_start(FstMutableHandle handle) {
  std::lock_guard<std::mutex> lk(g_mu);
  HostMutable* m = g_mutables.get(handle);
  return m ? m->start : FST_NO_STATE;
}
uint32_t fst_mutable_num_states(FstMutableHandle handle) {
  std::lock_guard<std::mutex> lk(g_mu);
  HostMutable* m = g_mutables.get(handle);
  return m ? m->num_states() : 0;
}
uint32_t fst_mutable_num_arcs(FstMutableHandle handle, uint32_t state) {
  std::lock_guard<std::mutex> lk(g_mu);
  HostMutable* m = g_mutables.get(handle);
  if (!m || state >= m->num_states()) return 0;
  return (uint32_t)m->arcs[state].size();
}
double fst_mutable_final_weight(FstMutableHandle handle, uint32_t state) {
  std::lock_guard<std::mutex> lk(g_mu);
  HostMutable* m = g_mutables.get(handle);
  if (!m || state >= m->num_states()) return std::numeric_limits<double>::infinity();
  return m->finals[state];
}
uint32_t fst_mutable_get_arcs(FstMutableHandle handle, uint32_t state, FstArc* buf, uint32_t buf_len) {
  std::lock_guard<std::mutex> lk(g_mu);
  HostMutable* m = g_mutables.get(handle);
  if (!m || state >= m->num_states()) return 0;
  const auto& v = m->arcs[state];
  uint32_t cnt = std::min<uint32_t>((uint32_t)v.size(), buf_len);
  if (buf) for (uint32_t i = 0; i < cnt; i++) { buf[i].ilabel = v[i].ilabel; buf[i].olabel = v[i].olabel; buf[i].weight = v[i].weight; buf[i].nextstate = v[i].nextstate; }
  return cnt;
}

// ── freeze (src/c-api.zig:507-526) ──
FstHandle fst_freeze(FstMutableHandle mutable_handle) {
  HostMutable snap;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    HostMutable* m = g_mutables.get(mutable_handle);
    if (!m) return FST_INVALID_HANDLE;
    snap = *m;
  }
  return new_frozen(HostFrozen::from_mutable(snap));
}

// ── frozen lifecycle + queries (src/c-api.zig:530-583) ──
void fst_free(FstHandle handle) {
  FrozenEntry* dead;
  { std::lock_guard<std::mutex> lk(g_mu); dead = g_frozen.remove(handle); }
  delete dead;
}
uint32_t fst_start(FstHandle handle) {
  std::lock_guard<std::mutex> lk(g_mu);
  FrozenEntry* f = g_frozen.get(handle);
  return f ? f->host->start() : FST_NO_STATE;
}
uint32_t fst_num_states(FstHandle handle) {
  std::lock_guard<std::mutex> lk(g_mu);
  FrozenEntry* f = g_frozen.get(handle);
  return f ? f->host->num_states() : 0;
}
uint32_t fst_num_arcs(FstHandle handle, uint32_t state) {
  std::lock_guard<std::mutex> lk(g_mu);
  FrozenEntry* f = g_frozen.get(handle);
  if (!f || state >= f->host->num_states()) return 0;
  return f->host->states()[state].num_arcs;
}
double fst_final_weight(FstHandle handle, uint32_t state) {
  std::lock_guard<std::mutex> lk(g_mu);
  FrozenEntry* f = g_frozen.get(handle);
  if (!f || state >= f->host->num_states()) return std::numeric_limits<double>::infinity();
  return f->host->states()[state].final_weight;
}
uint32_t fst_get_arcs(FstHandle handle, uint32_t state, FstArc* buf, uint32_t buf_len) {
  std::lock_guard<std::mutex> lk(g_mu);
  FrozenEntry* f = g_frozen.get(handle);
  if (!f || state >= f->host->num_states()) return 0;
  const ImgState& st = f->host->states()[state];
  const ImgArc* a = f->host->all_arcs() + st.arc_offset;
  uint32_t cnt = std::min<uint32_t>(st.num_arcs, buf_len);
  if (buf) for (uint32_t i = 0; i < cnt; i++) { buf[i].ilabel = a[i].ilabel; buf[i].olabel = a[i].olabel; buf[i].weight = a[i].weight; buf[i].nextstate = a[i].nextstate; }
  return cnt;
}

// ── native binary I/O (src/c-api.zig:601-639) ──
FstHandle fst_load(const char* path) {
  if (!path) return FST_INVALID_HANDLE;
  return new_frozen(HostFrozen::load_file(path));
}
FstError fst_save(FstHandle handle, const char* path) {
  if (!path) return FST_INVALID_ARG;
  FrozenEntry* f;
  { std::lock_guard<std::mutex> lk(g_mu); f = g_frozen.pin(handle); }
  if (!f) return FST_INVALID_ARG;
  PinGuard pg{handle};
  return f->host->save_file(path) ? FST_OK : FST_IO_ERROR;
}

// ── string helpers (src/c-api.zig:1334-1372) ──
FstMutableHandle fst_compile_string(const uint8_t* input, uint32_t len) {
  if (!input) return FST_INVALID_HANDLE;
  return new_mutable(HostMutable::from_bytes_string(input, len));
}
static int32_t print_tape(FstMutableHandle handle, bool output, uint8_t* buf, uint32_t buf_len) {
  std::string s;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    HostMutable* m = g_mutables.get(handle);
    if (!m) return -1;
    if (!m->print_tape(output, &s)) return -1;
  }
  if (s.size() > buf_len) return -1;
  if (buf && !s.empty()) std::memcpy(buf, s.data(), s.size());
  return (int32_t)s.size();
}
int32_t fst_print_string(FstMutableHandle handle, uint8_t* buf, uint32_t buf_len) { return print_tape(handle, false, buf, buf_len); }
int32_t fst_print_output_string(FstMutableHandle handle, uint8_t* buf, uint32_t buf_len) { return print_tape(handle, true, buf, buf_len); }

// Is `a` exactly what fst_compile_string builds (src/string.zig:24-50 with input == output: states 0..n, state i has
// the one arc (b+1 : b+1 / One -> i+1), only state n is final with weight One)?  Then the single call is the batched
// search of one byte string (the lean / fast kernels instead of the general-left-operand warp kernel: same search,
// ~4x lower latency).
static bool as_byte_string(const HostMutable& a, std::vector<uint8_t>* bytes) {
  const uint32_t ns = a.num_states();
  if (ns == 0 || a.start != 0) return false;
  const uint32_t n = ns - 1;
  auto bits = [](double x) { unsigned long long b; std::memcpy(&b, &x, 8); return b; };
  const unsigned long long inf_bits = bits(std::numeric_limits<double>::infinity());
  bytes->resize(n);
  for (uint32_t i = 0; i < n; i++) {
    if (a.arcs[i].size() != 1 || bits(a.finals[i]) != inf_bits) return false;
    const HostArc& x = a.arcs[i][0];
    if (x.ilabel != x.olabel || x.ilabel == 0 || x.ilabel > 256 || x.nextstate != i + 1 || bits(x.weight) != 0ull) return false;
    (*bytes)[i] = (uint8_t)(x.ilabel - 1);
  }
  return a.arcs[n].empty() && bits(a.finals[n]) == 0ull;
}

static FstError host_batch(FstHandle b, FstHandle b2, const uint8_t* bytes, const uint64_t* offsets,
                           uint32_t n_strings, FstB200BatchResult** out, int semantics = -1, uint32_t flags = 0);
void fst_b200_batch_free(FstB200BatchResult* r);

// ── THE HOT PATH, single problem (src/c-api.zig:744-811) ──
FstMutableHandle fst_compose_frozen_shortest_path(FstMutableHandle a_handle, FstHandle b_handle, uint32_t n) {
  const bool trace = trace_enabled();
  auto t0 = std::chrono::steady_clock::now();
  auto us = [&]() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(); };
  HostMutable a;
  FrozenEntry* fb;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    HostMutable* pa = g_mutables.get(a_handle);
    if (!pa) { trace_line("sp_invalid_a", a_handle, b_handle, 0, 0, 0, 0, us()); return FST_INVALID_HANDLE; }
    a = *pa;   // snapshot under the lock (:754)
    fb = g_frozen.pin(b_handle);
    if (!fb) { trace_line("sp_invalid_b", a_handle, b_handle, a.num_states(), a.total_arcs(), 0, 0, us()); return FST_INVALID_HANDLE; }
  }
  PinGuard pg{b_handle};
  const size_t in_states = a.num_states(), in_arcs = a.total_arcs();
  auto fail = [&]() { trace_line("sp_compose_error", a_handle, b_handle, in_states, in_arcs, 0, 0, us()); return FST_INVALID_HANDLE; };

  HostMutable result;
  // compose-shortest-path.zig:30-33
  if (a.start == kNoState || fb->host->start() == kNoState || n == 0) {
    // empty FST
  } else if (n != 1) {
    return fail();
  } else {
    // documented deviations: NaN weights are rejected (reference: UB in math.order);
    // left-operand state ids must fit 30 bits (tuple key packing).
    bool lhs_neg = false, lhs_nan = false;
    for (uint32_t s = 0; s < a.num_states(); s++) {
      if (std::isnan(a.finals[s])) lhs_nan = true;
      if (a.finals[s] < 0) lhs_neg = true;
      for (const HostArc& x : a.arcs[s]) { if (std::isnan(x.weight)) lhs_nan = true; if (x.weight < 0) lhs_neg = true; }
    }
    if (lhs_nan || fb->host->has_nan || a.num_states() >= (1u << 30)) return fail();
    if (!device_available()) {
      std::fprintf(stderr, "[libfst_b200] no CUDA device: fst_compose_frozen_shortest_path has no CPU fallback\n");
      return fail();
    }
    cudaError_t err;
    Engine* en = Engine::for_current_device(&err);
    if (!en) return fail();
    DeviceFst* img = fb->image_for(en->device, &err);
    if (!img) return fail();
    int32_t status = kStNoPath; double fw = 0;
    std::vector<uint32_t> il, ol; std::vector<double> w;
    std::vector<uint8_t> str;
    if (as_byte_string(a, &str)) {
      // a compiled string: one-string batch (always the lazy semantics of this entry point)
      const uint64_t off[2] = {0, str.size()};
      const uint8_t dummy = 0;
      FstB200BatchResult* res = nullptr;
      if (host_batch(b_handle, FST_INVALID_HANDLE, str.empty() ? &dummy : str.data(), off, 1, &res, 0, 0) != FST_OK) return fail();
      const int32_t s1 = res->status[0];
      status = (s1 == FST_B200_PATH || s1 == FST_B200_NOT_BYTES) ? kStPath : (s1 == FST_B200_NO_PATH ? kStNoPath : kStInternal);
      if (status == kStPath) {
        const uint64_t lo = res->path_offsets[0], hi = res->path_offsets[1];
        il.assign(res->ilabels + lo, res->ilabels + hi); ol.assign(res->olabels + lo, res->olabels + hi); w.assign(res->weights + lo, res->weights + hi);
        fw = res->final_weights[0];
      }
      fst_b200_batch_free(res);
    } else {
      std::lock_guard<std::mutex> lk(en->mu);
      BatchCounters bc;
      err = en->run_general(img, a, lhs_neg, &status, &il, &ol, &w, &fw, &bc);
      t_last = bc;
    }
    if (err != cudaSuccess) return fail();
    if (status == kStPath) result = HostMutable::chain(il.data(), ol.data(), w.data(), il.size(), fw);
    else if (status == kStNoPath) { /* empty */ }
    else return fail();   // cycle hazard (reference: OOM) or too large
  }
  trace_line("sp_ok", a_handle, b_handle, in_states, in_arcs, result.num_states(), result.total_arcs(), us());
  uint64_t h = new_mutable(std::move(result));
  if (h == FST_INVALID_HANDLE) trace_line("sp_new_handle_oom", a_handle, b_handle, in_states, in_arcs, 0, 0, us());
  return h;
}

// ── teardown (src/c-api.zig:295-329) ──
void fst_teardown(void) {
  std::vector<HostMutable*> ms; std::vector<FrozenEntry*> fs;
  { std::lock_guard<std::mutex> lk(g_mu); ms = g_mutables.drain(); fs = g_frozen.drain(); }
  for (auto* p : ms) delete p;
  for (auto* p : fs) delete p;
  if (device_available()) {
    cudaError_t err;
    Engine* en = Engine::for_current_device(&err);
    if (en) { std::lock_guard<std::mutex> lk(en->mu); en->release_all(); }
    pinned_drain();
  }
}

// ── batched entry points ──
void fst_b200_batch_free(FstB200BatchResult* r);
static int32_t map_status(int32_t s) {
  switch (s) { case kStPath: return FST_B200_PATH; case kStNoPath: return FST_B200_NO_PATH; case kStCycle: return FST_B200_CYCLE;
               case kStInternal: return FST_B200_INTERNAL; case kStNotBytes: return FST_B200_NOT_BYTES; default: return FST_B200_TOO_LARGE; }
}

// Host-buffer batch against one transducer (b2 == FST_INVALID_HANDLE) or the two-stage pipeline b then b2.
static FstError host_batch(FstHandle b, FstHandle b2, const uint8_t* bytes, const uint64_t* offsets,
                           uint32_t n_strings, FstB200BatchResult** out, int semantics, uint32_t flags) {
  if (!out) return FST_INVALID_ARG;
  *out = nullptr;
  if (!offsets || (!bytes && n_strings && offsets[n_strings] > 0)) return FST_INVALID_ARG;
  for (uint32_t i = 0; i < n_strings; i++) if (offsets[i + 1] < offsets[i]) return FST_INVALID_ARG;
  const bool two = b2 != FST_INVALID_HANDLE;
  FrozenEntry *fb, *fb2 = nullptr;
  { std::lock_guard<std::mutex> lk(g_mu); fb = g_frozen.pin(b); }
  if (!fb) return FST_INVALID_ARG;
  PinGuard pg{b};
  if (two) { std::lock_guard<std::mutex> lk(g_mu); fb2 = g_frozen.pin(b2); }
  if (two && !fb2) return FST_INVALID_ARG;
  PinGuard pg2{two ? b2 : FST_INVALID_HANDLE};
  if (fb->host->has_nan || (two && fb2->host->has_nan)) return FST_INVALID_ARG;
  if (!device_available()) {
    std::fprintf(stderr, "[libfst_b200] no CUDA device: the batched search has no CPU fallback\n");
    return FST_INVALID_STATE;
  }
  cudaError_t err;
  Engine* en = Engine::for_current_device(&err);
  if (!en) return FST_INVALID_STATE;
  DeviceFst* img = fb->image_for(en->device, &err);
  if (!img) return err == cudaErrorMemoryAllocation ? FST_OOM : FST_INVALID_STATE;
  DeviceFst* img2 = two ? fb2->image_for(en->device, &err) : nullptr;
  if (two && !img2) return err == cudaErrorMemoryAllocation ? FST_OOM : FST_INVALID_STATE;

  const uint64_t nbytes = n_strings ? offsets[n_strings] - offsets[0] : 0;
  uint64_t max_len = 0;
  for (uint32_t i = 0; i < n_strings; i++) max_len = std::max<uint64_t>(max_len, offsets[i + 1] - offsets[i]);
  if (max_len >= (1u << 30)) return FST_INVALID_ARG;

  std::lock_guard<std::mutex> lk(en->mu);
  cudaStream_t stream = 0;
  const uint32_t n = n_strings;
  // device buffers of the call: engine-owned, grow-only (cudaMalloc/cudaFree per call cost ~0.3 s per step when
  // the search workspace holds most of HBM)
  uint8_t* d_bytes = nullptr; uint64_t* d_offsets = nullptr; int32_t* d_status = nullptr; uint64_t* d_poff = nullptr;
  uint32_t *d_il = nullptr, *d_ol = nullptr, *d_nt = nullptr; double *d_w = nullptr, *d_fin = nullptr;
  uint64_t* d_ooff = nullptr; uint8_t* d_obytes = nullptr;
  std::vector<uint64_t> rel(n + 1);
  for (uint32_t i = 0; i <= n; i++) rel[i] = offsets[i] - offsets[0];
  uint64_t path_cap = 2 * nbytes + 16ull * n + 1024;
  auto free_dev = [&]() {};
  BatchCounters bc;
  for (int attempt = 0;; attempt++) {
    if (en->ensure_io(n, nbytes, path_cap) != cudaSuccess) { cudaGetLastError(); return FST_OOM; }
    const Engine::IoBuffers& io = en->io();
    // the flat path arrays are grow-only: use all of them (a transducer that inserts output — a tagger — needs several
    // arcs per input byte; sizing every call from the input alone would run its first attempt in vain each time)
    path_cap = std::max<uint64_t>(path_cap, io.cap_path);
    d_bytes = io.bytes; d_offsets = io.offsets; d_status = io.status; d_poff = io.path_offsets; d_il = io.il; d_ol = io.ol; d_w = io.w;
    d_fin = io.final_w; d_nt = io.n_tuples; d_ooff = io.out_offsets; d_obytes = io.out_bytes;
    {
      NvtxRange nvtx_h2d("fstb200 input H2D");
      if (nbytes) cudaMemcpyAsync(d_bytes, bytes + offsets[0], nbytes, cudaMemcpyHostToDevice, stream);
      cudaMemcpyAsync(d_offsets, rel.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, stream);
    }
    err = en->run_batch(img, d_bytes, d_offsets, n, (uint32_t)max_len, d_status, d_poff, d_il, d_ol, d_w, d_fin, d_nt, path_cap, d_ooff, d_obytes, path_cap, stream, &bc,
                        nullptr, nullptr, semantics);
    if (err != cudaSuccess) { free_dev(); return err == cudaErrorMemoryAllocation ? FST_OOM : FST_INVALID_STATE; }
    // the flat path arrays were too small (the pool itself may have been large enough from an earlier call)
    const uint64_t need = std::max<uint64_t>(bc.path_required, bc.path_total);
    if (need > path_cap && attempt < 4) { path_cap = need + need / 4 + 1024; free_dev(); continue; }
    break;
  }
  if (two) {
    // stage 2: the output-tape strings of stage 1 (still in HBM) are the inputs; strings stage 1 could not
    // transduce keep stage 1's status.  Per string this is compile_string -> compose_frozen_shortest_path(b) ->
    // print_output_string -> compile_string -> compose_frozen_shortest_path(b2) of the reference's ITN flow
    // (README.md:177-189) without leaving the device.
    const int32_t* d_status1 = d_status;
    const uint8_t* d_in2 = d_obytes; const uint64_t* d_off2 = d_ooff;
    uint64_t nbytes2 = 0; uint32_t max_len2 = 0;
    if (n) cudaMemcpy(&nbytes2, d_ooff + n, 8, cudaMemcpyDeviceToHost);
    if (en->last_max_out_len(n, stream, &max_len2) != cudaSuccess) return FST_INVALID_STATE;
    const BatchCounters bc1 = bc;
    path_cap = 2 * nbytes2 + 16ull * n + 1024;
    for (int attempt = 0;; attempt++) {
      if (en->ensure_io(n, 0, path_cap, 1) != cudaSuccess) { cudaGetLastError(); return FST_OOM; }
      const Engine::IoBuffers& io = en->io(1);
      path_cap = std::max<uint64_t>(path_cap, io.cap_path);
      d_status = io.status; d_poff = io.path_offsets; d_il = io.il; d_ol = io.ol; d_w = io.w;
      d_fin = io.final_w; d_nt = io.n_tuples; d_ooff = io.out_offsets; d_obytes = io.out_bytes;
      err = en->run_batch(img2, d_in2, d_off2, n, max_len2, d_status, d_poff, d_il, d_ol, d_w, d_fin, d_nt, path_cap, d_ooff, d_obytes, path_cap, stream, &bc, d_status1,
                          nullptr, semantics);
      if (err != cudaSuccess) return err == cudaErrorMemoryAllocation ? FST_OOM : FST_INVALID_STATE;
      const uint64_t need = std::max<uint64_t>(bc.path_required, bc.path_total);
      if (need > path_cap && attempt < 4) { path_cap = need + need / 4 + 1024; continue; }
      break;
    }
    bc.launches += bc1.launches; bc.passes += bc1.passes; bc.relax += bc1.relax; bc.tuples += bc1.tuples; bc.device_ms += bc1.device_ms;
  }
  t_last = bc;
  // assemble the pinned host result
  NvtxRange nvtx_d2h("fstb200 result D2H");
  // FST_B200_RESULT_NO_PATHS: the caller only wants the output strings — the per-arc arrays stay on the device
  const bool want_paths = !(flags & FST_B200_RESULT_NO_PATHS);
  const uint64_t total = want_paths ? bc.path_total : 0;
  uint64_t out_total = 0;
  if (n) cudaMemcpy(&out_total, d_ooff + n, 8, cudaMemcpyDeviceToHost);
  auto al = [](size_t x) { return (x + 63) & ~(size_t)63; };
  size_t o_status = 0, o_poff = o_status + al((size_t)n * 4), o_il = o_poff + al((size_t)(n + 1) * 8), o_ol = o_il + al(total * 4),
         o_w = o_ol + al(total * 4), o_fin = o_w + al(total * 8), o_nt = o_fin + al((size_t)n * 8), o_ooff = o_nt + al((size_t)n * 4),
         o_ob = o_ooff + al((size_t)(n + 1) * 8), bytes_total = o_ob + al(out_total) + 64;
  auto* r = new (std::nothrow) BatchResultImpl();
  if (r) r->pinned = pinned_alloc(bytes_total, &r->pinned_bytes);
  if (!r || !r->pinned) { delete r; free_dev(); return FST_OOM; }
  uint8_t* hb = static_cast<uint8_t*>(r->pinned);
  if (n) {
    cudaMemcpyAsync(hb + o_status, d_status, (size_t)n * 4, cudaMemcpyDeviceToHost, stream);
    cudaMemcpyAsync(hb + o_fin, d_fin, (size_t)n * 8, cudaMemcpyDeviceToHost, stream);
    cudaMemcpyAsync(hb + o_nt, d_nt, (size_t)n * 4, cudaMemcpyDeviceToHost, stream);
  }
  cudaMemcpyAsync(hb + o_poff, d_poff, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, stream);
  cudaMemcpyAsync(hb + o_ooff, d_ooff, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, stream);
  if (total) {
    cudaMemcpyAsync(hb + o_il, d_il, total * 4, cudaMemcpyDeviceToHost, stream);
    cudaMemcpyAsync(hb + o_ol, d_ol, total * 4, cudaMemcpyDeviceToHost, stream);
    cudaMemcpyAsync(hb + o_w, d_w, total * 8, cudaMemcpyDeviceToHost, stream);
  }
  if (out_total) cudaMemcpyAsync(hb + o_ob, d_obytes, out_total, cudaMemcpyDeviceToHost, stream);
  err = cudaStreamSynchronize(stream);
  free_dev();
  if (err != cudaSuccess) { pinned_release(r->pinned, r->pinned_bytes); delete r; return FST_INVALID_STATE; }
  int32_t* hs = reinterpret_cast<int32_t*>(hb + o_status);
  for (uint32_t i = 0; i < n; i++) hs[i] = map_status(hs[i]);
  r->pub.n_strings = n;
  r->pub.status = hs;
  r->pub.path_offsets = reinterpret_cast<uint64_t*>(hb + o_poff);
  r->pub.ilabels = want_paths ? reinterpret_cast<uint32_t*>(hb + o_il) : nullptr;
  r->pub.olabels = want_paths ? reinterpret_cast<uint32_t*>(hb + o_ol) : nullptr;
  r->pub.weights = want_paths ? reinterpret_cast<double*>(hb + o_w) : nullptr;
  r->pub.final_weights = reinterpret_cast<double*>(hb + o_fin);
  r->pub.n_tuples = reinterpret_cast<uint32_t*>(hb + o_nt);
  r->pub.out_offsets = reinterpret_cast<uint64_t*>(hb + o_ooff);
  r->pub.out_bytes = hb + o_ob;
  r->pub.device_ms = bc.device_ms;
  r->pub.total_tuples = bc.tuples;
  r->pub.total_relax = bc.relax;
  r->pub.launches = bc.launches;
  r->pub.passes = bc.passes;
  *out = &r->pub;
  return FST_OK;
}

FstError fst_compose_frozen_shortest_path_batch(FstHandle b, const uint8_t* bytes, const uint64_t* offsets,
                                                uint32_t n_strings, FstB200BatchResult** out) {
  return host_batch(b, FST_INVALID_HANDLE, bytes, offsets, n_strings, out);
}
FstError fst_compose_frozen_shortest_path_batch_ex(FstHandle b, const uint8_t* bytes, const uint64_t* offsets,
                                                   uint32_t n_strings, uint32_t flags, FstB200BatchResult** out) {
  if (flags & ~(uint32_t)FST_B200_RESULT_NO_PATHS) { if (out) *out = nullptr; return FST_INVALID_ARG; }
  return host_batch(b, FST_INVALID_HANDLE, bytes, offsets, n_strings, out, -1, flags);
}

// Eager pair per call (BASELINE config 5): fst_compose_frozen then fst_shortest_path(., 1) of every string, as its own
// entry point — the semantics are an argument of the call, not process state.
FstError fst_b200_compose_frozen_then_shortest_path_batch(FstHandle b, const uint8_t* bytes, const uint64_t* offsets,
                                                          uint32_t n_strings, FstB200BatchResult** out) {
  return host_batch(b, FST_INVALID_HANDLE, bytes, offsets, n_strings, out, 1);
}

FstError fst_compose_frozen_shortest_path_pipeline(FstHandle first, FstHandle second, const uint8_t* bytes, const uint64_t* offsets,
                                                   uint32_t n_strings, FstB200BatchResult** out) {
  if (second == FST_INVALID_HANDLE) { if (out) *out = nullptr; return FST_INVALID_ARG; }
  return host_batch(first, second, bytes, offsets, n_strings, out);
}

// ── one batch over several GPUs (north_star (3)) ──
// The transducer is replicated (one device image per GPU, uploaded on first use), the batch is cut into contiguous
// chunks of roughly equal estimated cost, and one host thread per GPU pulls chunks from an atomic queue; every chunk
// is an ordinary host-buffer batch on that thread's device (own stream, own engine, async D2H into the chunk's pinned
// result).  No collective: strings are independent.  Output order = input order (chunks are contiguous and listed
// in order; inside a chunk the batch entry keeps input order).
struct MultiResultImpl {
  FstB200MultiResult pub;
  std::vector<uint64_t> first;
  std::vector<FstB200BatchResult*> chunks;
  std::vector<int32_t> device;
};

FstError fst_compose_frozen_shortest_path_batch_multi(FstHandle b, const uint8_t* bytes, const uint64_t* offsets, uint32_t n_strings,
                                                      const int32_t* devices, uint32_t n_devices, uint32_t chunks_per_device,
                                                      uint32_t flags, FstB200MultiResult** out) {
  if (!out) return FST_INVALID_ARG;
  *out = nullptr;
  if (flags & ~(uint32_t)FST_B200_RESULT_NO_PATHS) return FST_INVALID_ARG;
  if (!offsets || (!bytes && n_strings && offsets[n_strings] > 0)) return FST_INVALID_ARG;
  for (uint32_t i = 0; i < n_strings; i++) if (offsets[i + 1] < offsets[i]) return FST_INVALID_ARG;
  int visible = 0;
  if (cudaGetDeviceCount(&visible) != cudaSuccess || visible <= 0) {
    cudaGetLastError();
    std::fprintf(stderr, "[libfst_b200] no CUDA device: the batched search has no CPU fallback\n");
    return FST_INVALID_STATE;
  }
  std::vector<int> devs;
  if (n_devices == 0 || !devices) { for (int d = 0; d < visible; d++) devs.push_back(d); }
  else {
    for (uint32_t k = 0; k < n_devices; k++) {
      if (devices[k] < 0 || devices[k] >= visible) return FST_INVALID_ARG;
      for (int d : devs) if (d == devices[k]) return FST_INVALID_ARG;   // a device is listed once
      devs.push_back(devices[k]);
    }
  }
  // hold the handle for the whole call: a concurrent fst_free is deferred until the last chunk is done
  FrozenEntry* fb;
  { std::lock_guard<std::mutex> lk(g_mu); fb = g_frozen.pin(b); }
  if (!fb) return FST_INVALID_ARG;
  PinGuard pg{b};
  if (fb->host->has_nan) return FST_INVALID_ARG;

  auto* r = new (std::nothrow) MultiResultImpl();
  if (!r) return FST_OOM;
  // chunks of equal estimated cost (linear in the length: the search state of a string grows with its length)
  if (chunks_per_device == 0) chunks_per_device = 2;
  const uint64_t want_chunks = std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)devs.size() * chunks_per_device, n_strings));
  r->first.push_back(0);
  if (n_strings) {
    const long double total_cost = (long double)(offsets[n_strings] - offsets[0]) + 16.0L * n_strings;
    uint64_t next_cut = 1;
    for (uint32_t i = 0; i < n_strings && next_cut < want_chunks; i++) {
      const long double done = (long double)(offsets[i + 1] - offsets[0]) + 16.0L * (i + 1);
      if (done * want_chunks >= total_cost * next_cut) {
        if (i + 1 < n_strings && (uint64_t)(i + 1) > r->first.back()) r->first.push_back(i + 1);
        next_cut++;
      }
    }
    r->first.push_back(n_strings);
  }
  const uint32_t n_chunks = (uint32_t)r->first.size() - 1;
  r->chunks.assign(n_chunks, nullptr);
  r->device.assign(n_chunks, -1);
  std::atomic<uint32_t> next{0};
  std::atomic<int> first_err{FST_OK};
  std::vector<double> dev_ms(devs.size(), 0.0);
  std::vector<BatchCounters> dev_cnt(devs.size());
  const auto t0 = std::chrono::steady_clock::now();
  auto worker = [&](size_t w) {
    if (cudaSetDevice(devs[w]) != cudaSuccess) { cudaGetLastError(); first_err = FST_INVALID_STATE; return; }
    for (;;) {
      const uint32_t k = next.fetch_add(1);
      if (k >= n_chunks || first_err.load() != FST_OK) break;
      const uint32_t lo = (uint32_t)r->first[k], hi = (uint32_t)r->first[k + 1];
      FstB200BatchResult* res = nullptr;
      const FstError e = host_batch(b, FST_INVALID_HANDLE, bytes, offsets + lo, hi - lo, &res, -1, flags);
      if (e != FST_OK) { int expect = FST_OK; first_err.compare_exchange_strong(expect, (int)e); break; }
      r->chunks[k] = res; r->device[k] = devs[w];
      dev_ms[w] += res->device_ms;
      dev_cnt[w].launches += res->launches; dev_cnt[w].passes += res->passes; dev_cnt[w].relax += res->total_relax; dev_cnt[w].tuples += res->total_tuples;
    }
  };
  if (devs.size() == 1) {
    int cur = 0; cudaGetDevice(&cur);
    worker(0);
    cudaSetDevice(cur);
  } else {
    std::vector<std::thread> th;
    for (size_t w = 0; w < devs.size(); w++) th.emplace_back(worker, w);
    for (auto& t : th) t.join();
  }
  if (first_err.load() != FST_OK) {
    for (auto* c : r->chunks) fst_b200_batch_free(c);
    delete r;
    return (FstError)first_err.load();
  }
  r->pub.n_strings = n_strings; r->pub.n_chunks = n_chunks;
  r->pub.chunk_first = r->first.data(); r->pub.chunks = r->chunks.data(); r->pub.chunk_device = r->device.data();
  r->pub.n_devices = (uint32_t)devs.size();
  r->pub.wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  r->pub.device_ms = 0; r->pub.total_relax = 0; r->pub.total_tuples = 0; r->pub.launches = 0;
  for (size_t w = 0; w < devs.size(); w++) {
    r->pub.device_ms = std::max(r->pub.device_ms, dev_ms[w]);
    r->pub.total_relax += dev_cnt[w].relax; r->pub.total_tuples += dev_cnt[w].tuples; r->pub.launches += dev_cnt[w].launches;
  }
  *out = &r->pub;
  return FST_OK;
}

void fst_b200_multi_free(FstB200MultiResult* r) {
  if (!r) return;
  auto* impl = reinterpret_cast<MultiResultImpl*>(r);   // pub is the first member
  for (auto* c : impl->chunks) fst_b200_batch_free(c);
  delete impl;
}

// ── eager lattices ──
struct LatticeResultImpl {
  FstB200LatticeResult pub;
  std::vector<int32_t> status; std::vector<uint64_t> state_off, arc_off; std::vector<uint32_t> arc_begin, il, ol, next;
  std::vector<double> fin, w;
};

FstError fst_b200_compose_frozen_lattice_batch(FstHandle b, const uint8_t* bytes, const uint64_t* offsets,
                                               uint32_t n_strings, FstB200LatticeResult** out) {
  if (!out) return FST_INVALID_ARG;
  *out = nullptr;
  if (!offsets || (!bytes && n_strings && offsets[n_strings] > 0)) return FST_INVALID_ARG;
  for (uint32_t i = 0; i < n_strings; i++) if (offsets[i + 1] < offsets[i]) return FST_INVALID_ARG;
  FrozenEntry* fb;
  { std::lock_guard<std::mutex> lk(g_mu); fb = g_frozen.pin(b); }
  if (!fb) return FST_INVALID_ARG;
  PinGuard pg{b};
  if (fb->host->has_nan) return FST_INVALID_ARG;
  if (!device_available()) {
    std::fprintf(stderr, "[libfst_b200] no CUDA device: the lattice builder has no CPU fallback\n");
    return FST_INVALID_STATE;
  }
  cudaError_t err;
  Engine* en = Engine::for_current_device(&err);
  if (!en) return FST_INVALID_STATE;
  DeviceFst* img = fb->image_for(en->device, &err);
  if (!img) return err == cudaErrorMemoryAllocation ? FST_OOM : FST_INVALID_STATE;
  if (!img->lean_ok) return FST_INVALID_ARG;   // the eager kernels need finite non-negative weights
  const uint32_t n = n_strings;
  const uint64_t nbytes = n ? offsets[n] - offsets[0] : 0;
  uint64_t max_len = 0;
  for (uint32_t i = 0; i < n; i++) max_len = std::max<uint64_t>(max_len, offsets[i + 1] - offsets[i]);
  if (max_len >= (1u << 30)) return FST_INVALID_ARG;
  std::vector<uint64_t> rel(n + 1);
  for (uint32_t i = 0; i <= n; i++) rel[i] = offsets[i] - offsets[0];

  std::lock_guard<std::mutex> lk(en->mu);
  cudaStream_t stream = 0;
  struct DevBuf { void* p = nullptr; ~DevBuf() { cudaFree(p); } };
  DevBuf b_sb, b_ab, b_ns, b_na, b_begin, b_fin, b_il, b_ol, b_nx, b_w;
  auto alloc = [](DevBuf& d, size_t bytes) { cudaFree(d.p); d.p = nullptr; return cudaMalloc(&d.p, std::max<size_t>(bytes, 16)) == cudaSuccess; };
  if (!alloc(b_sb, (size_t)n * 8) || !alloc(b_ab, (size_t)n * 8) || !alloc(b_ns, (size_t)n * 4) || !alloc(b_na, (size_t)n * 8)) return FST_OOM;
  LatticeOut lat;
  lat.state_cap = (nbytes + n) * 8 + 1024; lat.arc_cap = lat.state_cap * 4;
  BatchCounters bc;
  uint64_t path_cap = 2 * nbytes + 16ull * n + 1024;
  for (int attempt = 0;; attempt++) {
    if (!alloc(b_begin, lat.state_cap * 4) || !alloc(b_fin, lat.state_cap * 8) || !alloc(b_il, lat.arc_cap * 4) || !alloc(b_ol, lat.arc_cap * 4) ||
        !alloc(b_nx, lat.arc_cap * 4) || !alloc(b_w, lat.arc_cap * 8)) { cudaGetLastError(); return FST_OOM; }
    lat.d_state_base = static_cast<uint64_t*>(b_sb.p); lat.d_arc_base = static_cast<uint64_t*>(b_ab.p);
    lat.d_n_states = static_cast<uint32_t*>(b_ns.p); lat.d_n_arcs = static_cast<uint64_t*>(b_na.p);
    lat.d_arc_begin = static_cast<uint32_t*>(b_begin.p); lat.d_final = static_cast<double*>(b_fin.p);
    lat.d_il = static_cast<uint32_t*>(b_il.p); lat.d_ol = static_cast<uint32_t*>(b_ol.p); lat.d_next = static_cast<uint32_t*>(b_nx.p);
    lat.d_w = static_cast<double*>(b_w.p);
    if (en->ensure_io(n, nbytes, path_cap) != cudaSuccess) { cudaGetLastError(); return FST_OOM; }
    const Engine::IoBuffers& io = en->io();
    if (nbytes) cudaMemcpyAsync(io.bytes, bytes + offsets[0], nbytes, cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(io.offsets, rel.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, stream);
    err = en->run_batch(img, io.bytes, io.offsets, n, (uint32_t)max_len, io.status, io.path_offsets, io.il, io.ol, io.w, io.final_w, io.n_tuples,
                        path_cap, nullptr, nullptr, 0, stream, &bc, nullptr, &lat);
    if (err != cudaSuccess) return err == cudaErrorMemoryAllocation ? FST_OOM : FST_INVALID_STATE;
    const uint64_t need = std::max<uint64_t>(bc.path_required, bc.path_total);
    bool again = false;
    if (need > path_cap) { path_cap = need + need / 4 + 1024; again = true; }
    if (lat.states_required > lat.state_cap) { lat.state_cap = lat.states_required + lat.states_required / 8 + 1024; again = true; }
    if (lat.arcs_required > lat.arc_cap) { lat.arc_cap = lat.arcs_required + lat.arcs_required / 8 + 1024; again = true; }
    if (!again) break;
    if (attempt >= 3) return FST_OOM;
  }
  t_last = bc;
  // bring everything back and put the strings in input order (the device places them in completion order)
  auto* r = new (std::nothrow) LatticeResultImpl();
  if (!r) return FST_OOM;
  std::vector<uint64_t> sb(n), ab(n), na(n); std::vector<uint32_t> ns(n);
  r->status.resize(n);
  const Engine::IoBuffers& io = en->io();
  if (n) {
    cudaMemcpy(sb.data(), lat.d_state_base, (size_t)n * 8, cudaMemcpyDeviceToHost); cudaMemcpy(ab.data(), lat.d_arc_base, (size_t)n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(ns.data(), lat.d_n_states, (size_t)n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(na.data(), lat.d_n_arcs, (size_t)n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(r->status.data(), io.status, (size_t)n * 4, cudaMemcpyDeviceToHost);
  }
  const uint64_t S = lat.states_required, A = lat.arcs_required;
  std::vector<uint32_t> h_begin(S), h_il(A), h_ol(A), h_nx(A); std::vector<double> h_fin(S), h_w(A);
  if (S) { cudaMemcpy(h_begin.data(), lat.d_arc_begin, S * 4, cudaMemcpyDeviceToHost); cudaMemcpy(h_fin.data(), lat.d_final, S * 8, cudaMemcpyDeviceToHost); }
  if (A) { cudaMemcpy(h_il.data(), lat.d_il, A * 4, cudaMemcpyDeviceToHost); cudaMemcpy(h_ol.data(), lat.d_ol, A * 4, cudaMemcpyDeviceToHost);
           cudaMemcpy(h_nx.data(), lat.d_next, A * 4, cudaMemcpyDeviceToHost); cudaMemcpy(h_w.data(), lat.d_w, A * 8, cudaMemcpyDeviceToHost); }
  if (cudaDeviceSynchronize() != cudaSuccess) { delete r; return FST_INVALID_STATE; }
  r->state_off.assign(n + 1, 0); r->arc_off.assign(n + 1, 0);
  for (uint32_t i = 0; i < n; i++) { r->state_off[i + 1] = r->state_off[i] + ns[i]; r->arc_off[i + 1] = r->arc_off[i] + na[i]; }
  r->arc_begin.resize(S); r->fin.resize(S); r->il.resize(A); r->ol.resize(A); r->next.resize(A); r->w.resize(A);
  for (uint32_t i = 0; i < n; i++) {
    // the search status of the eager pair says whether a best path exists; the lattice itself exists whenever the
    // string was processed (status PATH / NO_PATH / CYCLE): report PATH for a built lattice
    const int32_t s = r->status[i];
    const bool built = (s == kStPath || s == kStNoPath || s == kStCycle) && (ns[i] > 0);
    r->status[i] = built ? FST_B200_PATH : (ns[i] == 0 && (s == kStPath || s == kStNoPath) ? FST_B200_NO_PATH : map_status(s));
    if (!built) { r->state_off[i + 1] = r->state_off[i]; r->arc_off[i + 1] = r->arc_off[i]; continue; }
    std::copy(h_begin.begin() + sb[i], h_begin.begin() + sb[i] + ns[i], r->arc_begin.begin() + r->state_off[i]);
    std::copy(h_fin.begin() + sb[i], h_fin.begin() + sb[i] + ns[i], r->fin.begin() + r->state_off[i]);
    std::copy(h_il.begin() + ab[i], h_il.begin() + ab[i] + na[i], r->il.begin() + r->arc_off[i]);
    std::copy(h_ol.begin() + ab[i], h_ol.begin() + ab[i] + na[i], r->ol.begin() + r->arc_off[i]);
    std::copy(h_nx.begin() + ab[i], h_nx.begin() + ab[i] + na[i], r->next.begin() + r->arc_off[i]);
    std::copy(h_w.begin() + ab[i], h_w.begin() + ab[i] + na[i], r->w.begin() + r->arc_off[i]);
  }
  r->pub.n_strings = n; r->pub.status = r->status.data(); r->pub.state_offsets = r->state_off.data(); r->pub.arc_offsets = r->arc_off.data();
  r->pub.arc_begin = r->arc_begin.data(); r->pub.final_weights = r->fin.data(); r->pub.ilabels = r->il.data(); r->pub.olabels = r->ol.data();
  r->pub.weights = r->w.data(); r->pub.nextstates = r->next.data(); r->pub.device_ms = bc.device_ms; r->pub.launches = bc.launches;
  *out = &r->pub;
  return FST_OK;
}

void fst_b200_lattice_free(FstB200LatticeResult* r) {
  if (r) delete reinterpret_cast<LatticeResultImpl*>(r);   // pub is the first member
}

void fst_b200_batch_free(FstB200BatchResult* r) {
  if (!r) return;
  auto* impl = reinterpret_cast<BatchResultImpl*>(r);   // pub is the first member
  pinned_release(impl->pinned, impl->pinned_bytes);
  delete impl;
}

FstError fst_b200_batch_device(FstHandle b, const uint8_t* d_bytes, const uint64_t* d_offsets, uint32_t n_strings,
                               uint32_t max_len, const FstB200DeviceOut* o, void* stream) {
  if (!o || !d_offsets || !o->d_status || !o->d_path_offsets || !o->d_final_weights || !o->d_n_tuples) return FST_INVALID_ARG;
  if (max_len >= (1u << 30)) return FST_INVALID_ARG;
  FrozenEntry* fb;
  { std::lock_guard<std::mutex> lk(g_mu); fb = g_frozen.pin(b); }
  if (!fb) return FST_INVALID_ARG;
  PinGuard pg{b};
  if (fb->host->has_nan) return FST_INVALID_ARG;
  if (!device_available()) return FST_INVALID_STATE;
  cudaError_t err;
  Engine* en = Engine::for_current_device(&err);
  if (!en) return FST_INVALID_STATE;
  DeviceFst* img = fb->image_for(en->device, &err);
  if (!img) return err == cudaErrorMemoryAllocation ? FST_OOM : FST_INVALID_STATE;
  std::lock_guard<std::mutex> lk(en->mu);
  BatchCounters bc;
  // the dense table is sized by the longest string: measure it (the caller's max_len is a hint; a longer string would
  // write past its arena)
  uint32_t measured = 0;
  if (en->measure_max_len(d_offsets, n_strings, static_cast<cudaStream_t>(stream), &measured) != cudaSuccess) return FST_INVALID_STATE;
  if (measured >= (1u << 30)) return FST_INVALID_ARG;
  max_len = std::max(max_len, measured);
  err = en->run_batch(img, d_bytes, d_offsets, n_strings, max_len, o->d_status, o->d_path_offsets, o->d_ilabels, o->d_olabels, o->d_weights,
                      o->d_final_weights, o->d_n_tuples, o->path_capacity, nullptr, nullptr, 0, static_cast<cudaStream_t>(stream), &bc);
  t_last = bc;
  if (err != cudaSuccess) return err == cudaErrorMemoryAllocation ? FST_OOM : FST_INVALID_STATE;
  if (bc.path_required > o->path_capacity || bc.path_total > o->path_capacity) return FST_OOM;
  return FST_OK;
}

FstError fst_b200_configure(const FstB200Config* cfg) {
  if (!cfg) return FST_INVALID_ARG;
  uint32_t g = cfg->lanes_per_string;
  if (!(g == 0 || g == 4 || g == 8 || g == 16 || g == 32)) return FST_INVALID_ARG;
  Config& c = global_config();
  if (cfg->engine > 7 || cfg->semantics > 1) return FST_INVALID_ARG;
  c.workspace_bytes = cfg->workspace_bytes; c.lanes_per_string = g; c.tuples_hint = cfg->tuples_hint; c.exhaustive = cfg->exhaustive; c.engine = cfg->engine; c.semantics = cfg->semantics;
  return FST_OK;
}

uint64_t fst_b200_last_path_required(void) { return std::max<uint64_t>(t_last.path_required, t_last.path_total); }

void fst_b200_last_occupancy(uint32_t* resident, uint32_t* capacity) {
  if (resident) *resident = t_last.resident;
  if (capacity) *capacity = t_last.capacity;
}

void fst_b200_last_counters(uint32_t* launches, uint64_t* relaxations, double* device_ms) {
  if (launches) *launches = t_last.launches;
  if (relaxations) *relaxations = t_last.relax;
  if (device_ms) *device_ms = t_last.device_ms;
}

int32_t fst_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char* fst_b200_version(void) { return "libfst_b200 0.1 (sm_100a)"; }

}  // extern "C"
