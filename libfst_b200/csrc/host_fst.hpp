// Host-side FST containers for the B200 engine.
//
// These mirror the two-phase model of the reference (build-time mutable graph,
// frozen contiguous image) only as far as the hot path needs them:
//   * HostMutable  ~ MutableFst(W)   reference src/mutable-fst.zig:45-218
//   * HostFrozen   ~ Fst(W) bytes    reference src/fst.zig:16-40, :160-273
// The frozen image is byte-compatible with the reference's native binary format
// (Header 24 B | StateEntry 16 B[] | PackedArc 24 B[]) so fst_load/fst_save
// interoperate with files written by the Zig library.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <vector>

namespace fstb200 {

constexpr uint32_t kNoState = 0xFFFFFFFFu;   // src/arc.zig:13
constexpr uint32_t kImageMagic = 0x46535421; // src/fst.zig:12 ("FST!")
constexpr uint16_t kImageVersion = 1;        // src/fst.zig:13

struct HostArc {
  uint32_t ilabel, olabel;
  double weight;
  uint32_t nextstate;
};

// Byte image records (src/fst.zig:16-40).
struct ImgHeader {
  uint32_t magic; uint16_t version; uint8_t weight_type; uint8_t flags;
  uint32_t num_states; uint32_t num_arcs; uint32_t start_state; uint32_t pad;
};
struct ImgState { uint32_t arc_offset; uint32_t num_arcs; double final_weight; };
struct ImgArc { uint32_t ilabel; uint32_t olabel; double weight; uint32_t nextstate; uint32_t pad; };
static_assert(sizeof(ImgHeader) == 24 && sizeof(ImgState) == 16 && sizeof(ImgArc) == 24, "image layout");

inline bool weight_is_zero(double w) { return std::isinf(w); }  // src/weight.zig:30-32

// Arc order of a frozen state (src/arc.zig:46-54).
inline bool arc_less(const HostArc& a, const HostArc& b) {
  if (a.ilabel != b.ilabel) return a.ilabel < b.ilabel;
  if (a.olabel != b.olabel) return a.olabel < b.olabel;
  if (a.weight < b.weight) return true;
  if (a.weight > b.weight) return false;
  return a.nextstate < b.nextstate;
}

class HostMutable {
 public:
  uint32_t start = kNoState;
  std::vector<double> finals;
  std::vector<std::vector<HostArc>> arcs;

  uint32_t num_states() const { return (uint32_t)finals.size(); }
  uint32_t add_state() { finals.push_back(std::numeric_limits<double>::infinity()); arcs.emplace_back(); return num_states() - 1; }
  void add_states(size_t n) { finals.resize(finals.size() + n, std::numeric_limits<double>::infinity()); arcs.resize(arcs.size() + n); }
  size_t total_arcs() const { size_t t = 0; for (auto& v : arcs) t += v.size(); return t; }

  // src/string.zig:24-50 with input == output (compileString).
  static HostMutable from_bytes_string(const uint8_t* s, size_t n) {
    HostMutable m;
    m.add_states(n + 1);
    m.start = 0;
    m.finals[n] = 0.0;
    for (size_t i = 0; i < n; i++) m.arcs[i].push_back(HostArc{(uint32_t)s[i] + 1, (uint32_t)s[i] + 1, 0.0, (uint32_t)(i + 1)});
    return m;
  }

  // Build the result chain of a search (src/ops/compose-shortest-path.zig:382-400).
  static HostMutable chain(const uint32_t* il, const uint32_t* ol, const double* w, size_t k, double final_w) {
    HostMutable m;
    m.add_states(k + 1);
    m.start = 0;
    m.finals[k] = final_w;
    for (size_t i = 0; i < k; i++) m.arcs[i].push_back(HostArc{il[i], ol[i], w[i], (uint32_t)(i + 1)});
    return m;
  }

  // src/string.zig:64-97.  false == the reference's `null` (-1 at the C ABI).
  // Deviations, both documented in DESIGN.md: a chain that never reaches a final
  // state (the reference spins) and labels > 256 (checked UB) return false.
  bool print_tape(bool output_tape, std::string* out) const {
    if (start == kNoState) return false;
    out->clear();
    uint32_t cur = start;
    size_t steps = 0;
    for (;;) {
      if (!weight_is_zero(finals[cur]) && arcs[cur].empty()) return true;
      if (arcs[cur].size() != 1) return false;
      const HostArc& a = arcs[cur][0];
      uint32_t l = output_tape ? a.olabel : a.ilabel;
      if (l != 0) { if (l > 256) return false; out->push_back((char)(uint8_t)(l - 1)); }
      cur = a.nextstate;
      if (cur == kNoState || cur >= num_states()) return false;
      if (++steps > finals.size()) return false;
    }
  }
};

// Frozen image + derived facts the device upload needs.
class HostFrozen {
 public:
  std::vector<uint64_t> storage;  // 8-byte aligned backing store
  size_t nbytes = 0;
  bool has_nan = false;
  bool has_negative = false;      // any arc or final weight < 0 (incl. -inf): non-monotone search
  bool fully_sorted = true;       // every state's arcs are in the full freeze order (src/arc.zig:46-54); fromBytes only
                                  // checks the ilabels (src/fst.zig:227-273), so a loaded image may not be
  uint32_t max_out_degree = 0;

  const uint8_t* bytes() const { return reinterpret_cast<const uint8_t*>(storage.data()); }
  uint8_t* bytes_mut() { return reinterpret_cast<uint8_t*>(storage.data()); }
  const ImgHeader& header() const { return *reinterpret_cast<const ImgHeader*>(bytes()); }
  const ImgState* states() const { return reinterpret_cast<const ImgState*>(bytes() + sizeof(ImgHeader)); }
  const ImgArc* all_arcs() const {
    return reinterpret_cast<const ImgArc*>(bytes() + sizeof(ImgHeader) + (size_t)header().num_states * sizeof(ImgState));
  }
  uint32_t num_states() const { return header().num_states; }
  uint32_t num_arcs() const { return header().num_arcs; }
  uint32_t start() const { return header().start_state; }

  void alloc(size_t n) { nbytes = n; storage.assign((n + 7) / 8, 0); }

  void scan_facts() {
    has_nan = has_negative = false; max_out_degree = 0;
    const ImgState* st = states(); const ImgArc* ar = all_arcs();
    for (uint32_t i = 0; i < num_states(); i++) {
      max_out_degree = std::max(max_out_degree, st[i].num_arcs);
      double f = st[i].final_weight;
      if (std::isnan(f)) has_nan = true;
      if (f < 0) has_negative = true;
    }
    for (uint32_t i = 0; i < num_arcs(); i++) {
      double w = ar[i].weight;
      if (std::isnan(w)) has_nan = true;
      if (w < 0) has_negative = true;
    }
    fully_sorted = true;
    for (uint32_t i = 0; i < num_states() && fully_sorted; i++) {
      const ImgArc* a = ar + st[i].arc_offset;
      for (uint32_t j = 1; j < st[i].num_arcs; j++) {
        const HostArc x{a[j - 1].ilabel, a[j - 1].olabel, a[j - 1].weight, a[j - 1].nextstate}, y{a[j].ilabel, a[j].olabel, a[j].weight, a[j].nextstate};
        if (arc_less(y, x)) { fully_sorted = false; break; }
      }
    }
  }

  // Freeze (src/fst.zig:160-224): stable-sort each state's arcs with arc_less
  // (std.mem.sort is stable; src/mutable-fst.zig:148-153), then pack.
  static std::unique_ptr<HostFrozen> from_mutable(const HostMutable& m) {
    auto f = std::make_unique<HostFrozen>();
    uint32_t ns = m.num_states();
    size_t total = m.total_arcs();
    if (total > 0xFFFFFFFFull) return nullptr;
    f->alloc(sizeof(ImgHeader) + (size_t)ns * sizeof(ImgState) + total * sizeof(ImgArc));
    ImgHeader* h = reinterpret_cast<ImgHeader*>(f->bytes_mut());
    *h = ImgHeader{kImageMagic, kImageVersion, 0, 0, ns, (uint32_t)total, m.start, 0};
    ImgState* st = reinterpret_cast<ImgState*>(f->bytes_mut() + sizeof(ImgHeader));
    ImgArc* ar = reinterpret_cast<ImgArc*>(f->bytes_mut() + sizeof(ImgHeader) + (size_t)ns * sizeof(ImgState));
    uint32_t off = 0;
    std::vector<HostArc> tmp;
    for (uint32_t s = 0; s < ns; s++) {
      tmp = m.arcs[s];
      std::stable_sort(tmp.begin(), tmp.end(), arc_less);
      st[s] = ImgState{off, (uint32_t)tmp.size(), m.finals[s]};
      for (const HostArc& a : tmp) ar[off++] = ImgArc{a.ilabel, a.olabel, a.weight, a.nextstate, 0};
    }
    f->scan_facts();
    return f;
  }

  // Validation of an external image: exactly the checks of src/fst.zig:227-273.
  static std::unique_ptr<HostFrozen> from_image(const uint8_t* data, size_t len) {
    if (len < sizeof(ImgHeader)) return nullptr;
    ImgHeader h; std::memcpy(&h, data, sizeof h);
    if (h.magic != kImageMagic || h.version != kImageVersion || h.weight_type != 0) return nullptr;
    size_t expect = sizeof(ImgHeader) + (size_t)h.num_states * sizeof(ImgState) + (size_t)h.num_arcs * sizeof(ImgArc);
    if (len != expect) return nullptr;
    if (h.num_states > 0 && h.start_state != kNoState && h.start_state >= h.num_states) return nullptr;
    if (h.num_states == 0 && h.start_state != kNoState) return nullptr;
    auto f = std::make_unique<HostFrozen>();
    f->alloc(len);
    std::memcpy(f->bytes_mut(), data, len);
    const ImgState* st = f->states(); const ImgArc* ar = f->all_arcs();
    for (uint32_t s = 0; s < h.num_states; s++) {
      if (st[s].arc_offset > h.num_arcs) return nullptr;
      if (st[s].num_arcs > h.num_arcs - st[s].arc_offset) return nullptr;
      uint32_t last = 0;
      for (uint32_t j = 0; j < st[s].num_arcs; j++) {
        const ImgArc& a = ar[st[s].arc_offset + j];
        if (a.nextstate >= h.num_states) return nullptr;
        if (j > 0 && a.ilabel < last) return nullptr;
        last = a.ilabel;
      }
    }
    f->scan_facts();
    return f;
  }

  static std::unique_ptr<HostFrozen> load_file(const char* path) {  // src/io/binary.zig:16-36
    FILE* fp = std::fopen(path, "rb");
    if (!fp) return nullptr;
    std::vector<uint8_t> buf;
    if (std::fseek(fp, 0, SEEK_END) == 0) {
      long sz = std::ftell(fp);
      if (sz >= 0) { buf.resize((size_t)sz); std::rewind(fp); if (std::fread(buf.data(), 1, buf.size(), fp) != buf.size()) buf.clear(); }
    }
    std::fclose(fp);
    if (buf.size() < sizeof(ImgHeader)) return nullptr;
    return from_image(buf.data(), buf.size());
  }
  bool save_file(const char* path) const {  // src/io/binary.zig:9-13
    FILE* fp = std::fopen(path, "wb");
    if (!fp) return false;
    bool ok = std::fwrite(bytes(), 1, nbytes, fp) == nbytes;
    return (std::fclose(fp) == 0) && ok;
  }
};

}  // namespace fstb200
