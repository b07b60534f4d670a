// Host-side engine: device image of a frozen transducer, per-device workspace,
// pass scheduling (adaptive arena sizing + retry of oversized strings), result
// assembly.  One Engine per CUDA device per process; calls are serialised per
// device by a mutex (the reference's frozen queries are re-entrant; here the GPU
// is the shared resource).
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "csp_kernels.cuh"
#include "csp_warp.cuh"
#include "csp_lean.cuh"
#include "csp_fast.cuh"
#include "csp_wave.cuh"
#include "host_fst.hpp"

namespace fstb200 {

#define FSTB_CUDA(expr)                                                                         \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      std::fprintf(stderr, "[libfst_b200] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e), \
                   __FILE__, __LINE__, cudaGetErrorString(_e));                                 \
      return _e;                                                                                \
    }                                                                                           \
  } while (0)

// NVTX range over a scope (nsys timeline of a call: upload / search pass / emit / D2H; header-only, no-op without a tool).
struct NvtxRange { explicit NvtxRange(const char* name) { nvtxRangePushA(name); } ~NvtxRange() { nvtxRangePop(); } };

struct Config {
  uint64_t workspace_bytes = 0;
  uint32_t lanes_per_string = 0;
  uint32_t tuples_hint = 0;
  uint32_t exhaustive = 0;
  uint32_t engine = 0;       // 0 auto, 1 general warp kernel, 2 lean + hash table, 3 lean + dense table,
                             // 4 wave (table auto), 5 wave + hash table, 6 wave + dense table, 7 lean (table auto)
  uint32_t semantics = 0;    // 0 lazy (composeShortestPath), 1 eager (compose then shortestPath; lean kernel only)
};
inline Config& global_config() { static Config c; return c; }

// ── device image of a frozen transducer ──
struct DeviceFst {
  int device = -1;
  DevFstView view{};
  void* block = nullptr;   // one allocation holding all arrays
  void* slab_block = nullptr;   // fixed-stride search records (lean kernel), may be null
  void* wslab_block = nullptr;  // leader-only fixed-stride search records (lean kernel with 8 lanes, wave kernel), may be null
  void* wslab4_block = nullptr; // ... 4 records per state (lean kernel with 4 lanes), built for sparse transducers only
  void* bigidx_block = nullptr; // label index of the states the leader slab cannot hold, may be null
  void* islab_block = nullptr;  // integer leader slab (fast kernel), may be null
  bool int_weights = false;     // every finite arc / final weight is a non-negative integer <= 4095: compact 8-byte table records apply
  bool wave_ok = false;         // wave slab built and at least 90 % of the states fit it
  size_t bytes = 0;
  bool serial = false;     // negative weights: literal sequential relax
  bool lean_ok = false;    // all arc weights finite and >= 0: the lean batched kernel applies
  uint32_t hint_tuples = 0;  // largest per-string tuple count seen so far (arena sizing)
  uint32_t hint_heap_mult = 1;  // radix-heap pool depth that sufficed so far (lean kernel)
  bool crec_failed = false;     // a search outgrew the compact records once: do not try them again
  int layout_skew = -1;         // dense table layout: -1 = a row per string position; >= 0 = a row per diagonal
                                // state - skew * position (chain-like transducers with input-epsilon arcs, see upload_fst)
  uint32_t lean_lanes = 32;     // lanes per string of the lean kernel: smallest of 8/16/32 that covers 90% of the states' arcs in one step
};

inline cudaError_t upload_fst(const HostFrozen& f, int device, DeviceFst** out) {
  NvtxRange nvtx_upload("fstb200 upload_fst");
  *out = nullptr;
  const uint32_t S = f.num_states(), A = f.num_arcs();
  const ImgState* st = f.states(); const ImgArc* ar = f.all_arcs();
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t o_rec = 0, o_fin = o_rec + al((size_t)S * 16), o_il = o_fin + al((size_t)S * 8), o_pl = o_il + al((size_t)A * 4);
  size_t o_sa = o_pl + al((size_t)A * 16);
  size_t total = o_sa + al((size_t)A * 16) + 256;
  std::vector<uint8_t> h(total, 0);
  uint4* rec = reinterpret_cast<uint4*>(h.data() + o_rec);
  double* fin = reinterpret_cast<double*>(h.data() + o_fin);
  uint32_t* il = reinterpret_cast<uint32_t*>(h.data() + o_il);
  uint4* pl = reinterpret_cast<uint4*>(h.data() + o_pl);
  uint4* sa = reinterpret_cast<uint4*>(h.data() + o_sa);
  uint32_t maxdeg = 0;
  for (uint32_t s = 0; s < S; s++) {
    uint32_t b = st[s].arc_offset, e = b + st[s].num_arcs, ee = b;
    while (ee < e && ar[ee].ilabel == 0) ee++;   // epsilon prefix (arcs are ilabel-sorted)
    rec[s] = make_uint4(b, ee, e, 0);
    fin[s] = st[s].final_weight;
    maxdeg = std::max(maxdeg, st[s].num_arcs);
  }
  bool all_finite = true;
  for (uint32_t a = 0; a < A; a++) {
    if (std::isinf(ar[a].weight)) all_finite = false;
    il[a] = ar[a].ilabel;
    unsigned long long wb; std::memcpy(&wb, &ar[a].weight, 8);
    pl[a] = make_uint4(ar[a].olabel, ar[a].nextstate, (uint32_t)(wb & 0xFFFFFFFFu), (uint32_t)(wb >> 32));
  }
  // search records: group the arcs of a state by (ilabel, nextstate); arcs are ilabel-sorted, so a group
  // lives inside one ilabel run
  // (an image whose ilabel runs are not in the full freeze order — accepted by fromBytes — takes the literal
  // serial-relax kernel: the parallel-arc fold and the back-track's first-tight-arc recovery assume that order)
  const bool lean_ok = !f.has_negative && all_finite && S < (1u << 31) && f.fully_sorted;
  if (lean_ok) {
    std::unordered_map<uint32_t, uint32_t> first_of;   // nextstate -> first arc of the current ilabel run
    for (uint32_t s = 0; s < S; s++) {
      const uint32_t b = st[s].arc_offset, e = b + st[s].num_arcs;
      uint32_t run = b;
      while (run < e) {
        uint32_t r2 = run;
        while (r2 < e && ar[r2].ilabel == ar[run].ilabel) r2++;
        first_of.clear();
        std::vector<double> wmin;   // indexed by (leader - run)
        wmin.assign(r2 - run, 0.0);
        for (uint32_t a = run; a < r2; a++) {
          auto it = first_of.find(ar[a].nextstate);
          if (it == first_of.end()) { first_of.emplace(ar[a].nextstate, a); wmin[a - run] = ar[a].weight; }
          else if (ar[a].weight < wmin[it->second - run]) wmin[it->second - run] = ar[a].weight;
        }
        for (uint32_t a = run; a < r2; a++) {
          const uint32_t leader = first_of[ar[a].nextstate];
          const bool dup = leader != a;
          unsigned long long wb; std::memcpy(&wb, &wmin[leader - run], 8);
          sa[a] = make_uint4(ar[a].ilabel, ar[a].nextstate | (dup ? 0x80000000u : 0u), (uint32_t)(wb & 0xFFFFFFFFu), (uint32_t)(wb >> 32));
        }
        run = r2;
      }
    }
  }
  bool int_weights = lean_ok;
  for (uint32_t a = 0; a < A && int_weights; a++) int_weights = ar[a].weight >= 0.0 && ar[a].weight <= 4095.0 && ar[a].weight == std::floor(ar[a].weight);
  for (uint32_t s = 0; s < S && int_weights; s++)
    if (!std::isinf(st[s].final_weight)) int_weights = st[s].final_weight >= 0.0 && st[s].final_weight <= 4095.0 && st[s].final_weight == std::floor(st[s].final_weight);
  auto d = new DeviceFst();
  d->int_weights = int_weights;
  d->device = device; d->bytes = total; d->serial = f.has_negative || !f.fully_sorted; d->lean_ok = lean_ok;
  cudaError_t e = cudaMalloc(&d->block, total);
  if (e != cudaSuccess) { delete d; return e; }
  e = cudaMemcpy(d->block, h.data(), total, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(d->block); delete d; return e; }
  uint8_t* base = static_cast<uint8_t*>(d->block);
  d->view.num_states = S; d->view.num_arcs = A; d->view.start = f.start(); d->view.max_degree = maxdeg;
  d->view.state_rec = reinterpret_cast<const uint4*>(base + o_rec);
  d->view.final_w = reinterpret_cast<const double*>(base + o_fin);
  d->view.ilabel = reinterpret_cast<const uint32_t*>(base + o_il);
  d->view.payload = reinterpret_cast<const uint4*>(base + o_pl);
  d->view.sarc = reinterpret_cast<const uint4*>(base + o_sa);
  {
    uint32_t le8 = 0, le16 = 0;
    for (uint32_t s = 0; s < S; s++) { le8 += st[s].num_arcs <= 8; le16 += st[s].num_arcs <= 16; }
    d->lean_lanes = (uint64_t)le8 * 10 >= (uint64_t)S * 9 ? 8 : ((uint64_t)le16 * 10 >= (uint64_t)S * 9 ? 16 : 32);
  }
  // Dense table layout.  On a chain-like transducer (arcs lead a bounded number of states forward) WITH input-epsilon
  // arcs, the reference's pop order keeps running along the cheapest match arcs: a tuple first met through a dearer arc
  // is lowered by the cheapest arc of its predecessor, gets ready with its old (small) id and is popped next, and so on
  // — chains (position + k, state + k * skew), skew = the state advance of the cheapest match arcs.  With a table row
  // per diagonal  state - skew * position  such a chain and its targets walk along a few rows (the next pop's records
  // sit in the sectors the last pop fetched); with a row per position every pop lands in another row.  Measured on
  // eps-dense (skew 1), 1 B200: len 33 39.1 k -> 54.3 k strings/s, len 96 13.8 k -> 15.4 k, len 251 2.33 k -> 2.70 k;
  // skew 0 = 13.6 k, skew 2 = 12.9 k at len 96 (profiles/README.md).  Without input-epsilon arcs every arc advances the
  // position and the front is a position (ambiguous chain: rows 3.60 M, diagonals 1.87 M strings/s): a row per
  // position.  Pure layout choice: any skew gives the same results (tests/test_gpu_literal.py::test_dense_table_layouts).
  {
    double wmin_m = std::numeric_limits<double>::infinity();
    for (uint32_t a = 0; a < A; a++)
      if (ar[a].ilabel != 0 && ar[a].ilabel <= 256u) wmin_m = std::min(wmin_m, ar[a].weight);
    uint64_t n_arcs = 0, out = 0, eps_states = 0, hist[66] = {0};
    for (uint32_t s = 0; s < S; s++) {
      bool has_eps = false;
      for (uint32_t a = st[s].arc_offset; a < st[s].arc_offset + st[s].num_arcs; a++) {
        if (ar[a].ilabel > 256u) continue;
        const int64_t dl = (int64_t)ar[a].nextstate - (int64_t)s;
        n_arcs++;
        if (dl < 0 || dl > 64) out++;
        if (ar[a].ilabel == 0) has_eps = true;
        else if (ar[a].weight == wmin_m && dl >= 0 && dl <= 64) hist[dl]++;
      }
      eps_states += has_eps;
    }
    int best = 0;
    for (int k = 1; k <= 64; k++) if (hist[k] > hist[best]) best = k;
    if (n_arcs > 0 && out * 20 <= n_arcs && eps_states * 2 >= S && hist[best] > 0) d->layout_skew = best;
    if (const char* e = std::getenv("LIBFST_B200_SKEW")) d->layout_skew = std::atoi(e);   // experiments / tests: -1, 0, 1, ...
    if (d->layout_skew > 64) d->layout_skew = 64;
  }
  d->view.slab = nullptr; d->view.slab_lanes = 0; d->view.pad0 = 0;
  const uint32_t GL = d->lean_lanes == 8 ? 16 : d->lean_lanes;   // 8 lanes read the leader slab (below); this one serves 16 / 32
  if (lean_ok && (uint64_t)S * GL * 16 <= (64ull << 20)) {
    // fixed-stride copy of the search records: one load per lane per pop without the state_rec hop
    std::vector<uint4> slab((size_t)S * GL, make_uint4(0xFFFFFFFFu, 0x80000000u, 0, 0));
    for (uint32_t s = 0; s < S; s++) {
      const uint32_t b = st[s].arc_offset, n = st[s].num_arcs;
      if (n <= GL) { for (uint32_t k = 0; k < n; k++) slab[(size_t)s * GL + k] = sa[b + k]; }
      else { for (uint32_t k = 0; k < GL; k++) slab[(size_t)s * GL + k] = make_uint4(0xFFFFFFFEu, 0x80000000u, 0, 0); }
    }
    if (cudaMalloc(&d->slab_block, slab.size() * 16) == cudaSuccess &&
        cudaMemcpy(d->slab_block, slab.data(), slab.size() * 16, cudaMemcpyHostToDevice) == cudaSuccess) {
      d->view.slab = static_cast<const uint4*>(d->slab_block); d->view.slab_lanes = GL;
    } else {
      cudaGetLastError(); cudaFree(d->slab_block); d->slab_block = nullptr;
    }
  }
  d->view.wslab = nullptr; d->view.wslab4 = nullptr; d->view.bigidx = nullptr; d->view.islab = nullptr;
  if (lean_ok && (uint64_t)S * kWaveSlots * 16 <= (256ull << 20)) {
    std::vector<uint4> ws((size_t)S * kWaveSlots, make_uint4(0xFFFFFFFFu, 0u, 0u, 0u));
    uint32_t n_big = 0;
    std::vector<uint4> lab, eps;
    for (uint32_t s = 0; s < S; s++) {
      const uint32_t b = st[s].arc_offset, e = b + st[s].num_arcs;
      lab.clear(); eps.clear();
      bool big = false;
      for (uint32_t a = b; a < e && !big; a++) {
        if (sa[a].y >> 31) continue;                       // folded into its leader
        if (sa[a].x > 256u) continue;                      // never matches a byte
        uint32_t cnt = 0;                                  // arcs of the group (the leader's run of equal ilabel)
        for (uint32_t a2 = a; a2 < e && ar[a2].ilabel == ar[a].ilabel; a2++) cnt += ar[a2].nextstate == ar[a].nextstate;
        if (cnt > 0xFFFFu) { big = true; break; }
        const uint4 r = make_uint4(sa[a].x | (cnt << 16), sa[a].y, sa[a].z, sa[a].w);
        (sa[a].x == 0u ? eps : lab).push_back(r);
      }
      if (big || lab.size() + eps.size() > kWaveSlots) {
        for (uint32_t k = 0; k < kWaveSlots; k++) ws[(size_t)s * kWaveSlots + k] = make_uint4(kWaveBig, 0x80000000u, 0u, 0u);
        n_big++;
        continue;
      }
      size_t k = 0;
      for (const uint4& r : lab) ws[(size_t)s * kWaveSlots + k++] = r;
      for (const uint4& r : eps) ws[(size_t)s * kWaveSlots + k++] = r;
    }
    if (cudaMalloc(&d->wslab_block, ws.size() * 16) == cudaSuccess &&
        cudaMemcpy(d->wslab_block, ws.data(), ws.size() * 16, cudaMemcpyHostToDevice) == cudaSuccess) {
      d->view.wslab = static_cast<const uint4*>(d->wslab_block);
      d->wave_ok = (uint64_t)n_big * 10 <= (uint64_t)S;
      // at least 90 % of the states fit 8 leader records: 8 lanes per string reading the leader slab
      if (d->wave_ok) d->lean_lanes = 8;
    } else {
      cudaGetLastError(); cudaFree(d->wslab_block); d->wslab_block = nullptr;
    }
    // 4-record variant for sparse transducers (at least 95 % of the states have <= 4 leader records, fewer than 2 on
    // average): 4 lanes per string, 8 strings per warp
    std::vector<uint4> w4;
    bool use4 = false;
    if (d->view.wslab) {
      w4.assign((size_t)S * 4, make_uint4(0xFFFFFFFFu, 0u, 0u, 0u));
      uint64_t n_big4 = 0, n_rec = 0;
      for (uint32_t s = 0; s < S; s++) {
        const uint4* r8 = &ws[(size_t)s * kWaveSlots];
        uint32_t k = 0;
        while (k < kWaveSlots && r8[k].x != 0xFFFFFFFFu && r8[k].x != kWaveBig) k++;
        if (r8[0].x == kWaveBig || k > 4) {
          for (uint32_t j = 0; j < 4; j++) w4[(size_t)s * 4 + j] = make_uint4(kWaveBig, 0x80000000u, 0u, 0u);
          n_big4++; n_rec += 8;
        } else {
          for (uint32_t j = 0; j < k; j++) w4[(size_t)s * 4 + j] = r8[j];
          n_rec += k;
        }
      }
      // measured on the WeText-style tagger: 4 lanes are SLOWER than 8 (427 k vs 465 k strings/s) — with 8 strings per
      // warp the group-local rare paths (level advance, fetch, finish) stall seven other strings: on request only
      use4 = n_big4 * 20 <= (uint64_t)S && n_rec < 2ull * S && std::getenv("LIBFST_B200_LANES4") != nullptr;
    }
    // label index for the states that do not fit their slab (trie roots, the 256-way identity state of a tagger ...);
    // one numbering for both slabs (every state the 8-record slab cannot hold is also too wide for the 4-record one)
    {
      std::vector<uint32_t> big_states;
      for (uint32_t s = 0; s < S; s++)
        if (ws[(size_t)s * kWaveSlots].x == kWaveBig || (use4 && w4[(size_t)s * 4].x == kWaveBig)) big_states.push_back(s);
      if (d->view.wslab && !big_states.empty() && (uint64_t)big_states.size() * 257 * 8 <= (64ull << 20)) {
        std::vector<uint2> bi(big_states.size() * 257, make_uint2(0u, 0u));
        for (uint32_t nb = 0; nb < big_states.size(); nb++) {
          const uint32_t s = big_states[nb];
          const uint32_t b = st[s].arc_offset, e = b + st[s].num_arcs;
          for (uint32_t a = b; a < e;) {
            uint32_t a2 = a;
            while (a2 < e && ar[a2].ilabel == ar[a].ilabel) a2++;
            if (ar[a].ilabel <= 256u) bi[(size_t)nb * 257 + ar[a].ilabel] = make_uint2(a, a2 - a);
            a = a2;
          }
          if (ws[(size_t)s * kWaveSlots].x == kWaveBig) for (uint32_t k = 0; k < kWaveSlots; k++) ws[(size_t)s * kWaveSlots + k].z = nb;
          if (use4) for (uint32_t k = 0; k < 4; k++) w4[(size_t)s * 4 + k].z = nb;
        }
        if (cudaMalloc(&d->bigidx_block, bi.size() * 8) == cudaSuccess &&
            cudaMemcpy(d->bigidx_block, bi.data(), bi.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMemcpy(d->wslab_block, ws.data(), ws.size() * 16, cudaMemcpyHostToDevice) == cudaSuccess) {
          d->view.bigidx = static_cast<const uint2*>(d->bigidx_block);
        } else {
          cudaGetLastError(); cudaFree(d->bigidx_block); d->bigidx_block = nullptr;
        }
      }
    }
    if (use4 && d->view.bigidx &&
        cudaMalloc(&d->wslab4_block, w4.size() * 16) == cudaSuccess &&
        cudaMemcpy(d->wslab4_block, w4.data(), w4.size() * 16, cudaMemcpyHostToDevice) == cudaSuccess) {
      d->view.wslab4 = static_cast<const uint4*>(d->wslab4_block);
      d->lean_lanes = 4;
    } else if (use4) {
      cudaGetLastError(); cudaFree(d->wslab4_block); d->wslab4_block = nullptr;
    }
    // integer leader slab of the fast kernel (csp_fast.cuh): weight pre-shifted to the distance field of the compact
    // table record; marker rows keep their label-index number; one all-idle row at index S
    d->view.islab = nullptr;
    if (int_weights && d->view.wslab && S < (1u << 30)) {
      std::vector<uint4> is((size_t)(S + 1) * kWaveSlots, make_uint4(0xFFFFFFFFu, 0u, 0u, 0u));
      for (size_t k = 0; k < (size_t)S * kWaveSlots; k++) {
        uint4 r = ws[k];
        if (r.x != 0xFFFFFFFFu && r.x != kWaveBig) {   // {ilabel, next << 1 | epsilon, weight << 12, arcs folded}
          double w; unsigned long long wb = ((unsigned long long)r.w << 32) | r.z; std::memcpy(&w, &wb, 8);
          const uint32_t il = r.x & 0xFFFFu;
          r = make_uint4(il, (r.y << 1) | (il == 0u ? 1u : 0u), (uint32_t)w << 12, r.x >> 16);
        }
        is[k] = r;
      }
      if (cudaMalloc(&d->islab_block, is.size() * 16) == cudaSuccess &&
          cudaMemcpy(d->islab_block, is.data(), is.size() * 16, cudaMemcpyHostToDevice) == cudaSuccess) {
        d->view.islab = static_cast<const uint4*>(d->islab_block);
      } else {
        cudaGetLastError(); cudaFree(d->islab_block); d->islab_block = nullptr;
      }
    }
  }
  *out = d;
  return cudaSuccess;
}
inline void free_device_fst(DeviceFst* d) {
  if (!d) return;
  int cur = 0; cudaGetDevice(&cur);
  if (cur != d->device) cudaSetDevice(d->device);
  cudaFree(d->block);
  cudaFree(d->slab_block);
  cudaFree(d->wslab_block);
  cudaFree(d->bigidx_block);
  cudaFree(d->wslab4_block);
  cudaFree(d->islab_block);
  if (cur != d->device) cudaSetDevice(cur);
  delete d;
}

// small helper kernels (plumbing)
// Strings of the pass just run (work items `items[0..n)`, null = identity) that must run again.
__global__ void collect_retry_kernel(const int32_t* status, const uint32_t* items, uint32_t n, uint32_t* order, uint32_t* count,
                                     uint32_t* heap_count, uint32_t* wide_count) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t i = items ? items[k] : k;
  if (status[i] == kStRetry || status[i] == kStRetryHeap || status[i] == kStRetryWide) {
    order[atomicAdd(count, 1u)] = i;
    if (status[i] == kStRetryHeap) atomicAdd(heap_count, 1u);
    if (status[i] == kStRetryWide) atomicAdd(wide_count, 1u);
  }
}
__global__ void mark_too_large_kernel(int32_t* status, uint32_t* path_len, const uint32_t* items, uint32_t n) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t i = items ? items[k] : k;
  if (status[i] == kStRetry || status[i] == kStRetryHeap || status[i] == kStRetryWide) { status[i] = kStTooLarge; path_len[i] = 0; }
}
__global__ void fill_retry_kernel(int32_t* status, uint32_t* path_len, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { status[i] = kStRetry; path_len[i] = 0; }
}
// Longest-first schedule: key = ~length so that an ascending sort puts the longest strings first.
__global__ void length_keys_kernel(const uint64_t* offsets, uint32_t n, uint32_t* keys, uint32_t* iota) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { keys[i] = ~(uint32_t)(offsets[i + 1] - offsets[i]); iota[i] = i; }
}
__global__ void widen_kernel(const uint32_t* in, uint64_t* out, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}
// Longest string of a device-resident batch (fst_b200_batch_device: the caller's max_len is only a hint).
__global__ void max_len_kernel(const uint64_t* offsets, uint32_t n, uint32_t* out) {
  uint32_t m = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint64_t d = offsets[i + 1] - offsets[i];
    m = max(m, d > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)d);
  }
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}
__global__ void max_u32_kernel(const uint32_t* in, uint32_t n, uint32_t* out) {
  uint32_t m = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, in[i]);
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// Device buffers of the optional eager-lattice output of run_batch (see SearchParams::lat_*).
struct LatticeOut {
  uint64_t* d_state_base = nullptr; uint64_t* d_arc_base = nullptr; uint32_t* d_n_states = nullptr; uint64_t* d_n_arcs = nullptr;   // [n]
  uint32_t* d_arc_begin = nullptr; double* d_final = nullptr; uint64_t state_cap = 0;                                              // per state
  uint32_t *d_il = nullptr, *d_ol = nullptr, *d_next = nullptr; double* d_w = nullptr; uint64_t arc_cap = 0;                       // per arc
  uint64_t states_required = 0, arcs_required = 0;   // out: totals of the batch (above the capacities: grow and run again)
};

struct BatchCounters {
  uint32_t launches = 0, passes = 0;
  unsigned long long relax = 0, tuples = 0;
  double device_ms = 0;
  uint64_t path_total = 0, path_required = 0;
  uint32_t max_tuples = 0;
  uint32_t resident = 0;     // strings searched concurrently by the first pass (arenas that fit the workspace budget)
  uint32_t capacity = 0;     // ... and how many the device could hold for this geometry (min of occupancy and memory)
};

class Engine {
 public:
  int device = 0;
  int sm_count = 0;
  std::mutex mu;

  static Engine* for_current_device(cudaError_t* err) {
    static std::mutex gm;
    static std::map<int, Engine*> engines;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { *err = e; return nullptr; }
    std::lock_guard<std::mutex> lk(gm);
    auto it = engines.find(dev);
    if (it != engines.end()) { *err = cudaSuccess; return it->second; }
    auto en = new Engine();
    en->device = dev;
    e = cudaDeviceGetAttribute(&en->sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) { delete en; *err = e; return nullptr; }
    e = cudaMalloc(&en->d_small_, 256);
    if (e == cudaSuccess) e = cudaMemset(en->d_small_, 0, 256);
    if (e != cudaSuccess) { delete en; *err = e; return nullptr; }
    e = cudaMallocHost(&en->h_small_, 256);
    if (e != cudaSuccess) { delete en; *err = e; return nullptr; }
    cudaEventCreate(&en->ev0_); cudaEventCreate(&en->ev1_);
    engines[dev] = en;
    *err = cudaSuccess;
    return en;
  }

  // Device-resident batch.  All d_* pointers are device memory on this device.
  // Synchronises `stream` before returning (the retry decision needs a read-back).
  cudaError_t run_batch(DeviceFst* fst, const uint8_t* d_bytes, const uint64_t* d_offsets, uint32_t n, uint32_t max_len,
                        int32_t* d_status, uint64_t* d_path_offsets, uint32_t* d_il, uint32_t* d_ol, double* d_w,
                        double* d_final, uint32_t* d_ntuples, uint64_t path_capacity,
                        uint64_t* d_out_offsets, uint8_t* d_out_bytes, uint64_t out_capacity,
                        cudaStream_t stream, BatchCounters* bc, const int32_t* d_skip = nullptr, LatticeOut* lat = nullptr,
                        int semantics = -1) {   // -1: the configured default; 0 lazy; 1 eager (per call, see c_api.cu)
    *bc = BatchCounters();
    if (n == 0) {
      FSTB_CUDA(cudaMemsetAsync(d_path_offsets, 0, 8, stream));
      if (d_out_offsets) FSTB_CUDA(cudaMemsetAsync(d_out_offsets, 0, 8, stream));
      return cudaStreamSynchronize(stream);
    }
    Config cfg = global_config();
    if (semantics >= 0) cfg.semantics = (uint32_t)semantics;
    if (lat) cfg.semantics = 1;   // the lattice is the eager pair's (compose.zig numbering)
    FSTB_CUDA(ensure_scratch(n, path_capacity));
    if (lat) {
      FSTB_CUDA(cudaMemsetAsync(lat->d_n_states, 0, (size_t)n * 4, stream));
      FSTB_CUDA(cudaMemsetAsync(lat->d_n_arcs, 0, (size_t)n * 8, stream));
    }
    // counters: [0] queue_head(u32) [1] retry_count(u32) [2..3] pool_cursor(u64) [4..5] relax [6..7] tuples [8] max_tuples
    uint32_t* d_cnt = static_cast<uint32_t*>(d_small_);
    FSTB_CUDA(cudaMemsetAsync(d_cnt, 0, 64, stream));
    FSTB_CUDA(cudaEventRecord(ev0_, stream));
    fill_retry_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_status, d_path_len_, n);
    bc->launches++;

    // Work segments: the batch in longest-first order, cut where the strings get so much shorter that their search state
    // (a dense table has (length + 1) * 2S records) is worth its own arena geometry — more strings in flight for the
    // short ones.  One segment (the whole batch, input order) for small or uniform batches.
    struct Seg { uint32_t begin, count, max_len; };
    std::vector<Seg> segs{{0u, n, max_len}};
    const uint32_t* d_sorted = nullptr;
    if (n >= 4096) {
      // strings leave the work queue longest first (search cost grows with the length): the last wave of a batch of
      // mixed lengths ends with short strings instead of a few long ones running alone
      length_keys_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_offsets, n, d_sort_keys_[0], d_sort_vals_[0]);
      size_t tmp = sort_tmp_bytes_;
      FSTB_CUDA(cub::DeviceRadixSort::SortPairs(d_sort_tmp_, tmp, d_sort_keys_[0], d_sort_keys_[1], d_sort_vals_[0], d_sort_vals_[1], (int)n, 0, 32, stream));
      d_sorted = d_sort_vals_[1];
      bc->launches += 3;
      constexpr uint32_t kMinSeg = 8192;
      // measured on the mixed eps-dense batch (57 216 strings, lengths 11..251): 3 858 strings/s in four segments vs
      // 4 117 in one — every segment ends with its own tail and the arenas are re-initialised per geometry; opt-in
      if (n >= 4 * kMinSeg && max_len >= 32 && fst->lean_ok && std::getenv("LIBFST_B200_SEGMENTS") != nullptr) {
        std::vector<uint32_t> keys(n);   // ~length, ascending == lengths descending
        FSTB_CUDA(cudaMemcpyAsync(keys.data(), d_sort_keys_[1], (size_t)n * 4, cudaMemcpyDeviceToHost, stream));
        FSTB_CUDA(cudaStreamSynchronize(stream));
        const uint32_t len0 = ~keys[0];
        if ((uint64_t)(~keys[n - 1]) * 4 <= (uint64_t)len0 * 3) {
          segs.clear();
          uint32_t begin = 0;
          for (int q = 3; q >= 0 && begin < n; q--) {
            const uint32_t thr = (uint32_t)((uint64_t)len0 * (uint32_t)q / 4);   // this class: thr < length
            const uint32_t end = q == 0 ? n : (uint32_t)(std::upper_bound(keys.begin() + begin, keys.end(), ~(thr + 1u) ) - keys.begin());
            if (end > begin) { segs.push_back(Seg{begin, end - begin, ~keys[begin]}); begin = end; }
          }
          // a class with too few strings to fill the device rides with its longer neighbour, or takes the next one along
          for (size_t i = 0; i < segs.size();) {
            if (segs[i].count < kMinSeg && i + 1 < segs.size()) { segs[i].count += segs[i + 1].count; segs.erase(segs.begin() + i + 1); }
            else if (segs[i].count < kMinSeg && i > 0) { segs[i - 1].count += segs[i].count; segs.erase(segs.begin() + i); }
            else i++;
          }
        }
      }
    }
    bool crec_ok = fst->int_weights && !fst->crec_failed;   // compact table records until a distance outgrows them
    bool abort_all = false;
    uint32_t seg_no = 0;
    for (const Seg& sg : segs) {
    if (abort_all) break;
    uint32_t tuple_cap = cfg.tuples_hint ? cfg.tuples_hint : (fst->hint_tuples ? fst->hint_tuples + fst->hint_tuples / 4 + 64 : 4096);
    uint32_t n_items = sg.count, heap_mult = fst->hint_heap_mult;
    const uint32_t* d_order = d_sorted ? d_sorted + sg.begin : nullptr;
    const uint32_t seg_max_len = sg.max_len;
    FSTB_CUDA(cudaMemsetAsync(d_cnt + 0, 0, 4, stream));   // queue head
    auto too_large = [&]() {
      mark_too_large_kernel<<<(n_items + 255) / 256, 256, 0, stream>>>(d_status, d_path_len_, d_order, n_items);
      bc->launches++;
    };
    for (uint32_t pass = 0;; pass++) {
      bc->passes++;
      Geom gm;
      if (!geometry(cfg, fst, seg_max_len, tuple_cap, heap_mult, &gm, crec_ok, workspace_budget(cfg))) { too_large(); break; }
      if (cfg.semantics == 1 && gm.kind != kLean) {
        std::fprintf(stderr, "[libfst_b200] eager semantics need finite non-negative weights (lean kernel)\n");
        too_large(); break;
      }
      const uint32_t max_groups = max_resident_groups(gm);
      const uint32_t want = std::min<uint32_t>(n_items, max_groups);
      const uint64_t ws = workspace_budget(cfg);
      const uint32_t fit = (uint32_t)std::min<uint64_t>(ws / gm.stride, 0xFFFFFFFFull);
      if (fit == 0) { too_large(); break; }   // even one arena does not fit the budget
      // full 128-thread blocks when the budget allows, else one partial block
      uint32_t gpb = 128 / gm.G, threads = 128, blocks;
      if (fit >= gpb) {
        blocks = std::min((want + gpb - 1) / gpb, fit / gpb);
      } else {
        blocks = 1; gpb = std::min(want, fit); threads = gpb * gm.G;
      }
      FSTB_CUDA(ensure_workspace(gm, blocks * gpb, stream));
      bc->launches += init_launches_; init_launches_ = 0;
      if (pass == 0 && seg_no == 0) { bc->resident = blocks * gpb; bc->capacity = std::min(max_groups, fit >= gpb ? (fit / gpb) * gpb : fit); }

      SearchParams p{};
      p.fst = fst->view;
      p.bytes = d_bytes; p.offsets = d_offsets; p.order = d_order; p.n_items = n_items; p.skip = d_skip;
      p.arena = static_cast<uint8_t*>(d_workspace_); p.arena_stride = gm.stride;
      p.hash_cap = gm.hash_cap; p.tuple_cap = gm.tuple_cap; p.heap_cap = gm.heap_cap; p.bag_cap = gm.bag_cap; p.exhaustive = cfg.exhaustive;
      p.dense = gm.dense ? 1u : 0u; p.tab_entries = gm.tab_entries;
      if (gm.kind == kLean || gm.kind == kWave) {
        const LeanLayout L = lean_layout((int)gm.G, gm.dense, gm.tab_entries, gm.tuple_cap, gm.heap_cap, gm.crec);
        p.off_keyof = L.off_keyof; p.off_l0 = L.off_l0; p.off_chunks = L.off_chunks;
        p.n1 = L.n1; p.smem_words = gm.smem_per_group / 4; p.dense_stride = fst->view.num_states * 2u;
        if (gm.skew < 0) { p.pos_h = 1u; p.pos_m2 = p.dense_stride; p.pos_c = 0u; p.pos_k = 0u; }
        else {
          // pos = ((state - skew P + skew max_len) * 2 + filter) * H + P + filter      (H = padded half row)
          const uint32_t H = diag_half_row(seg_max_len), W = 2u * H;
          p.pos_h = H; p.pos_c = (uint32_t)gm.skew * seg_max_len * W; p.pos_m2 = 1u - (uint32_t)gm.skew * W;
          p.pos_k = diag_shift() ? 1u : 0u;
        }
        p.key_sbits = 1; while ((1u << p.key_sbits) < p.dense_stride + (gm.fast ? 2u : 0u)) p.key_sbits++;   // fast kernel: room for the idle row S
        p.eager = cfg.semantics == 1 ? 1u : 0u;
      }
      p.queue_head = d_cnt + 0;
      p.pool_cursor = reinterpret_cast<unsigned long long*>(d_cnt + 2);
      p.relax_counter = reinterpret_cast<unsigned long long*>(d_cnt + 4);
      p.tuple_counter = reinterpret_cast<unsigned long long*>(d_cnt + 6);
      p.wave_stats = reinterpret_cast<unsigned long long*>(d_cnt + 16);   // outside the 64 bytes reset per batch: cumulative
      if (lat) {
        p.lat_cursors = reinterpret_cast<unsigned long long*>(d_cnt + 12);
        p.lat_state_cap = lat->state_cap; p.lat_arc_cap = lat->arc_cap;
        p.lat_state_base = lat->d_state_base; p.lat_arc_base = lat->d_arc_base; p.lat_n_states = lat->d_n_states; p.lat_n_arcs = lat->d_n_arcs;
        p.lat_arc_begin = lat->d_arc_begin; p.lat_final = lat->d_final;
        p.lat_il = lat->d_il; p.lat_ol = lat->d_ol; p.lat_next = lat->d_next; p.lat_w = lat->d_w;
      }
      p.status = d_status; p.path_len = d_path_len_; p.pool_off = d_pool_off_; p.final_w = d_final; p.n_tuples = d_ntuples;
      p.pool = d_pool_; p.pool_cap = pool_cap_;
      {
        NvtxRange r("fstb200 search pass");
        launch_search(gm, blocks, threads, p, stream);
      }
      bc->launches++;
      FSTB_CUDA(cudaGetLastError());
      // any string that overflowed its arena (or the pool)?
      FSTB_CUDA(cudaMemsetAsync(d_cnt + 1, 0, 4, stream));
      FSTB_CUDA(cudaMemsetAsync(d_cnt + 9, 0, 8, stream));
      collect_retry_kernel<<<(n_items + 255) / 256, 256, 0, stream>>>(d_status, d_order, n_items, d_order_buf_[pass & 1], d_cnt + 1, d_cnt + 9, d_cnt + 10);
      bc->launches++;
      FSTB_CUDA(cudaMemcpyAsync(h_small_, d_cnt, 64, cudaMemcpyDeviceToHost, stream));
      FSTB_CUDA(cudaStreamSynchronize(stream));
      const uint32_t* hc = static_cast<const uint32_t*>(h_small_);
      const uint32_t retry = hc[1], heap_retry = hc[9], wide_retry = hc[10];
      if (std::getenv("LIBFST_B200_DEBUG")) {
        unsigned long long ws[4] = {0, 0, 0, 0};
        cudaMemcpy(ws, d_cnt + 16, 32, cudaMemcpyDeviceToHost);
        if (gm.kind == kWave) std::fprintf(stderr, "[libfst_b200] wave stats (cumulative): chunk steps %llu, tuples popped by chunks %llu, single-pop steps %llu, abandoned chunks %llu\n", ws[0], ws[1], ws[2], ws[3]);
      }
      if (std::getenv("LIBFST_B200_DEBUG"))
        std::fprintf(stderr, "[libfst_b200] segment %u/%zu (max_len %u) pass %u kind %d fast %d skew %d G %u dense %d eager %d tuple_cap %u heap_cap %u items %u groups %u retry %u heap_retry %u\n",
                     seg_no, segs.size(), seg_max_len, pass, gm.kind, (int)gm.fast, gm.skew, gm.G, (int)gm.dense + (int)gm.crec, (int)gm.eager, gm.tuple_cap, gm.heap_cap, n_items, blocks * gpb, retry, heap_retry);
      unsigned long long pool_used; std::memcpy(&pool_used, hc + 2, 8);
      if (retry == 0) break;
      if (pool_used > pool_cap_) {
        // path pool too small: report the requirement; caller grows and re-runs
        bc->path_required = pool_used;
        too_large();
        abort_all = true;
        break;
      }
      // next pass: only the overflowed strings; 8x larger arenas and/or a 4x deeper radix-heap pool
      n_items = retry;
      d_order = d_order_buf_[pass & 1];
      if (heap_retry > 0) {
        if (heap_mult >= 4096) { too_large(); break; }
        heap_mult *= 4;
        if (heap_mult > fst->hint_heap_mult) fst->hint_heap_mult = heap_mult;
      }
      if (wide_retry > 0) { crec_ok = false; fst->crec_failed = true; }   // distances beyond 20 bits: 16-byte records from now on
      if (retry > heap_retry + wide_retry) {
        if (gm.kind == kLean && gm.dense && gm.tuple_cap >= gm.tab_entries) { too_large(); break; }   // cannot happen: N <= records
        if (tuple_cap > (1u << 31) / 8) { too_large(); break; }
        tuple_cap *= 8;
      }
      FSTB_CUDA(cudaMemsetAsync(d_cnt + 0, 0, 4, stream));   // queue head
    }
    seg_no++;
    }   // segments
    if (abort_all) {   // strings of the segments that never ran still carry the initial retry status
      mark_too_large_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_status, d_path_len_, nullptr, n);
      bc->launches++;
    }

    // ordered output: offsets = exclusive scan of path lengths, then un-reverse
    NvtxRange nvtx_emit("fstb200 emit");
    widen_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_path_len_, d_len64_, n);
    FSTB_CUDA(cudaMemsetAsync(d_len64_ + n, 0, 8, stream));
    size_t tmp = scan_tmp_bytes_;
    FSTB_CUDA(cub::DeviceScan::ExclusiveSum(d_scan_tmp_, tmp, d_len64_, d_path_offsets, (int)(n + 1), stream));
    bc->launches += 2;
    EmitParams e{};
    e.status = d_status; e.path_len = d_path_len_; e.pool_off = d_pool_off_; e.path_offsets = d_path_offsets; e.pool = d_pool_;
    e.n_strings = n; e.ilabels = d_il; e.olabels = d_ol; e.weights = d_w; e.path_capacity = path_capacity;
    e.out_offsets = nullptr; e.out_bytes = nullptr;
    if (d_out_offsets && d_out_bytes) {
      csp_count_out_kernel<<<std::min<uint32_t>((n + 7) / 8, 148 * 16), 256, 0, stream>>>(d_status, d_path_len_, d_pool_off_, d_pool_, n, d_out_len_);
      widen_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_out_len_, d_len64_, n);
      FSTB_CUDA(cub::DeviceScan::ExclusiveSum(d_scan_tmp_, tmp, d_len64_, d_out_offsets, (int)(n + 1), stream));
      bc->launches += 3;
      e.out_offsets = d_out_offsets; e.out_bytes = d_out_bytes;
      (void)out_capacity;   // out bytes <= path arcs <= path_capacity; caller sizes it so
    }
    csp_emit_kernel<<<std::min<uint32_t>((n + 7) / 8, 148 * 16), 256, 0, stream>>>(e);
    bc->launches++;
    max_u32_kernel<<<std::min<uint32_t>((n + 255) / 256, 592), 256, 0, stream>>>(d_ntuples, n, d_cnt + 8);
    bc->launches++;
    FSTB_CUDA(cudaEventRecord(ev1_, stream));
    FSTB_CUDA(cudaMemcpyAsync(h_small_, d_cnt, 64, cudaMemcpyDeviceToHost, stream));
    FSTB_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(h_small_) + 64, d_path_offsets + n, 8, cudaMemcpyDeviceToHost, stream));
    FSTB_CUDA(cudaStreamSynchronize(stream));
    FSTB_CUDA(cudaGetLastError());
    const uint32_t* hc = static_cast<const uint32_t*>(h_small_);
    std::memcpy(&bc->relax, hc + 4, 8);
    std::memcpy(&bc->tuples, hc + 6, 8);
    bc->max_tuples = hc[8];
    if (lat) { std::memcpy(&lat->states_required, hc + 12, 8); std::memcpy(&lat->arcs_required, hc + 14, 8); }
    std::memcpy(&bc->path_total, static_cast<uint8_t*>(h_small_) + 64, 8);
    float ms = 0; cudaEventElapsedTime(&ms, ev0_, ev1_);
    bc->device_ms = ms;
    if (bc->max_tuples > fst->hint_tuples) fst->hint_tuples = bc->max_tuples;
    return cudaSuccess;
  }

  // One general left operand (drop-in single call).  Host arrays in, host path out.
  cudaError_t run_general(DeviceFst* fst, const HostMutable& lhs, bool lhs_negative, int32_t* status,
                          std::vector<uint32_t>* il, std::vector<uint32_t>* ol, std::vector<double>* w, double* final_w,
                          BatchCounters* bc) {
    *bc = BatchCounters();
    cudaStream_t stream = 0;
    // upload lhs as CSR in stored order
    const uint32_t S = lhs.num_states(); const uint32_t A = (uint32_t)lhs.total_arcs();
    std::vector<uint32_t> off(S + 1), ail(A), aol(A), anx(A); std::vector<double> aw(A), fin(S);
    uint32_t k = 0;
    for (uint32_t s = 0; s < S; s++) {
      off[s] = k; fin[s] = lhs.finals[s];
      for (const HostArc& a : lhs.arcs[s]) { ail[k] = a.ilabel; aol[k] = a.olabel; aw[k] = a.weight; anx[k] = a.nextstate; k++; }
    }
    off[S] = k;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t o_off = 0, o_fin = o_off + al((S + 1) * 4), o_il = o_fin + al((size_t)S * 8), o_ol = o_il + al((size_t)A * 4),
           o_w = o_ol + al((size_t)A * 4), o_nx = o_w + al((size_t)A * 8), total = o_nx + al((size_t)A * 4) + 256;
    std::vector<uint8_t> h(total, 0);
    std::memcpy(h.data() + o_off, off.data(), (S + 1) * 4);
    if (S) std::memcpy(h.data() + o_fin, fin.data(), (size_t)S * 8);
    if (A) { std::memcpy(h.data() + o_il, ail.data(), (size_t)A * 4); std::memcpy(h.data() + o_ol, aol.data(), (size_t)A * 4);
             std::memcpy(h.data() + o_w, aw.data(), (size_t)A * 8); std::memcpy(h.data() + o_nx, anx.data(), (size_t)A * 4); }
    // device copy of the left operand: engine-owned, grow-only (a cudaMalloc + cudaFree per call is most of the
    // latency of a small single call)
    if (total > lhs_cap_) {
      cudaFree(d_lhs_); d_lhs_ = nullptr; lhs_cap_ = 0;
      const size_t want = total + total / 2 + 4096;
      FSTB_CUDA(cudaMalloc(&d_lhs_, want));
      lhs_cap_ = want;
    }
    uint8_t* d_lhs = d_lhs_;
    FSTB_CUDA(cudaMemcpyAsync(d_lhs, h.data(), total, cudaMemcpyHostToDevice, stream));
    const uint64_t path_capacity = 1u << 16;
    FSTB_CUDA(ensure_scratch(1, path_capacity));
    uint32_t* d_cnt = static_cast<uint32_t*>(d_small_);
    const Config cfg = global_config();
    const bool serial = fst->serial || lhs_negative;
    uint32_t tuple_cap = 4096;
    FSTB_CUDA(cudaEventRecord(ev0_, stream));
    for (;;) {
      bc->passes++;
      Geom gm;
      Config c1 = cfg; c1.engine = 1;   // a general left operand never takes the lean (byte-string) kernel
      DeviceFst f1 = *fst; f1.serial = serial;
      if (!geometry(c1, &f1, 0, tuple_cap, 1, &gm)) { *status = kStTooLarge; return cudaSuccess; }
      if (gm.stride > workspace_budget(cfg)) { *status = kStTooLarge; return cudaSuccess; }
      const uint32_t hash_cap = gm.hash_cap, heap_cap = gm.heap_cap, bag_cap = gm.bag_cap, smem_per_warp = gm.smem_per_group;
      const uint64_t stride = gm.stride;
      tuple_cap = gm.tuple_cap;
      FSTB_CUDA(ensure_workspace(gm, 1, stream));
      bc->launches += init_launches_; init_launches_ = 0;
      FSTB_CUDA(cudaMemsetAsync(d_cnt, 0, 64, stream));
      SearchParams p{};
      p.fst = fst->view;
      p.n_items = 1;
      p.lhs.num_states = S; p.lhs.num_arcs = A; p.lhs.start = lhs.start;
      p.lhs.arc_off = reinterpret_cast<const uint32_t*>(d_lhs + o_off);
      p.lhs.final_w = reinterpret_cast<const double*>(d_lhs + o_fin);
      p.lhs.ilabel = reinterpret_cast<const uint32_t*>(d_lhs + o_il);
      p.lhs.olabel = reinterpret_cast<const uint32_t*>(d_lhs + o_ol);
      p.lhs.weight = reinterpret_cast<const double*>(d_lhs + o_w);
      p.lhs.next = reinterpret_cast<const uint32_t*>(d_lhs + o_nx);
      p.arena = static_cast<uint8_t*>(d_workspace_); p.arena_stride = stride;
      p.hash_cap = hash_cap; p.tuple_cap = tuple_cap; p.heap_cap = heap_cap; p.bag_cap = bag_cap; p.exhaustive = cfg.exhaustive;
      p.queue_head = d_cnt + 0;
      p.pool_cursor = reinterpret_cast<unsigned long long*>(d_cnt + 2);
      p.relax_counter = reinterpret_cast<unsigned long long*>(d_cnt + 4);
      p.tuple_counter = reinterpret_cast<unsigned long long*>(d_cnt + 6);
      p.status = d_status1_; p.path_len = d_path_len_; p.pool_off = d_pool_off_; p.final_w = d_final1_; p.n_tuples = d_out_len_;
      p.pool = d_pool_; p.pool_cap = pool_cap_;
      if (serial) {
        csp_general_kernel<true><<<1, 32, 0, stream>>>(p);
      } else {
        if (smem_per_warp > 40 * 1024) cudaFuncSetAttribute(csp_general_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_per_warp);
        csp_general_warp_kernel<<<1, 32, smem_per_warp, stream>>>(p);
      }
      bc->launches++;
      FSTB_CUDA(cudaGetLastError());
      struct { int32_t st; uint32_t plen; double fw; } r;
      FSTB_CUDA(cudaMemcpyAsync(&r.st, d_status1_, 4, cudaMemcpyDeviceToHost, stream));
      FSTB_CUDA(cudaMemcpyAsync(&r.plen, d_path_len_, 4, cudaMemcpyDeviceToHost, stream));
      FSTB_CUDA(cudaMemcpyAsync(&r.fw, d_final1_, 8, cudaMemcpyDeviceToHost, stream));
      FSTB_CUDA(cudaMemcpyAsync(h_small_, d_cnt, 64, cudaMemcpyDeviceToHost, stream));
      FSTB_CUDA(cudaStreamSynchronize(stream));
      if (r.st == kStRetry) {
        unsigned long long used; std::memcpy(&used, static_cast<uint32_t*>(h_small_) + 2, 8);
        if (used > pool_cap_) { FSTB_CUDA(ensure_scratch(1, used * 2)); continue; }
        if (tuple_cap > (1u << 31) / 8) { *status = kStTooLarge; return cudaSuccess; }
        tuple_cap *= 8;
        continue;
      }
      *status = r.st; *final_w = r.fw;
      il->assign(r.plen, 0); ol->assign(r.plen, 0); w->assign(r.plen, 0.0);
      if (r.st == kStPath && r.plen) {
        std::vector<PoolArc> rev(r.plen);
        FSTB_CUDA(cudaMemcpy(rev.data(), d_pool_, (size_t)r.plen * sizeof(PoolArc), cudaMemcpyDeviceToHost));   // pool_off == 0 (cursor reset)
        for (uint32_t i = 0; i < r.plen; i++) { const PoolArc& a = rev[r.plen - 1 - i]; (*il)[i] = a.ilabel; (*ol)[i] = a.olabel; (*w)[i] = a.weight; }
      }
      std::memcpy(&bc->relax, static_cast<uint32_t*>(h_small_) + 4, 8);
      std::memcpy(&bc->tuples, static_cast<uint32_t*>(h_small_) + 6, 8);
      FSTB_CUDA(cudaEventRecord(ev1_, stream));
      FSTB_CUDA(cudaEventSynchronize(ev1_));
      float ms = 0; cudaEventElapsedTime(&ms, ev0_, ev1_); bc->device_ms = ms;
      return cudaSuccess;
    }
  }

  // Device-side input/output buffers of the host-buffer batch entry (grow-only).
  struct IoBuffers {
    uint8_t* bytes = nullptr; uint64_t* offsets = nullptr; int32_t* status = nullptr; uint64_t* path_offsets = nullptr;
    uint32_t *il = nullptr, *ol = nullptr, *n_tuples = nullptr; double *w = nullptr, *final_w = nullptr;
    uint64_t* out_offsets = nullptr; uint8_t* out_bytes = nullptr;
    uint64_t cap_n = 0, cap_bytes = 0, cap_path = 0;
  };
  const IoBuffers& io(int set = 0) const { return io_sets_[set]; }
  // set 0: the batch entry (and stage 1 of the pipeline entry); set 1: stage 2 of the pipeline entry
  cudaError_t ensure_io(uint32_t n, uint64_t nbytes, uint64_t path_cap, int set = 0) {
    IoBuffers& io_ = io_sets_[set];
    if (n > io_.cap_n || io_.offsets == nullptr) {
      cudaFree(io_.offsets); cudaFree(io_.status); cudaFree(io_.path_offsets); cudaFree(io_.final_w); cudaFree(io_.n_tuples); cudaFree(io_.out_offsets);
      io_.offsets = nullptr; io_.status = nullptr; io_.path_offsets = nullptr; io_.final_w = nullptr; io_.n_tuples = nullptr; io_.out_offsets = nullptr;
      io_.cap_n = 0;
      const uint64_t m = (uint64_t)n + n / 8 + 16;
      FSTB_CUDA(cudaMalloc(&io_.offsets, (m + 1) * 8)); FSTB_CUDA(cudaMalloc(&io_.status, m * 4 + 16));
      FSTB_CUDA(cudaMalloc(&io_.path_offsets, (m + 1) * 8)); FSTB_CUDA(cudaMalloc(&io_.final_w, m * 8 + 16));
      FSTB_CUDA(cudaMalloc(&io_.n_tuples, m * 4 + 16)); FSTB_CUDA(cudaMalloc(&io_.out_offsets, (m + 1) * 8));
      io_.cap_n = m;
    }
    if (nbytes > io_.cap_bytes || io_.bytes == nullptr) {
      cudaFree(io_.bytes); io_.bytes = nullptr; io_.cap_bytes = 0;
      const uint64_t m = nbytes + nbytes / 8 + 256;
      FSTB_CUDA(cudaMalloc(&io_.bytes, m + 16));
      io_.cap_bytes = m;
    }
    if (path_cap > io_.cap_path || io_.il == nullptr) {
      cudaFree(io_.il); cudaFree(io_.ol); cudaFree(io_.w); cudaFree(io_.out_bytes);
      io_.il = io_.ol = nullptr; io_.w = nullptr; io_.out_bytes = nullptr; io_.cap_path = 0;
      const uint64_t m = path_cap + path_cap / 8 + 256;
      FSTB_CUDA(cudaMalloc(&io_.il, m * 4)); FSTB_CUDA(cudaMalloc(&io_.ol, m * 4)); FSTB_CUDA(cudaMalloc(&io_.w, m * 8));
      FSTB_CUDA(cudaMalloc(&io_.out_bytes, m + 16));
      io_.cap_path = m;
    }
    return cudaSuccess;
  }

  void release_all() {
    for (IoBuffers& io_ : io_sets_) {
      cudaFree(io_.bytes); cudaFree(io_.offsets); cudaFree(io_.status); cudaFree(io_.path_offsets); cudaFree(io_.il); cudaFree(io_.ol);
      cudaFree(io_.w); cudaFree(io_.final_w); cudaFree(io_.n_tuples); cudaFree(io_.out_offsets); cudaFree(io_.out_bytes);
      io_ = IoBuffers();
    }
    cudaFree(d_workspace_); d_workspace_ = nullptr; workspace_bytes_ = 0; layout_groups_ = 0; budget_cache_ = 0;
    cudaFree(d_lhs_); d_lhs_ = nullptr; lhs_cap_ = 0;
    free_scratch();
  }

  uint64_t pool_capacity() const { return pool_cap_; }

  // Longest string of a device-resident batch (synchronises the stream).
  cudaError_t measure_max_len(const uint64_t* d_offsets, uint32_t n, cudaStream_t stream, uint32_t* out) {
    uint32_t* d_cnt = static_cast<uint32_t*>(d_small_);
    FSTB_CUDA(cudaMemsetAsync(d_cnt + 11, 0, 4, stream));
    if (n) max_len_kernel<<<std::min<uint32_t>((n + 255) / 256, 592), 256, 0, stream>>>(d_offsets, n, d_cnt + 11);
    FSTB_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(h_small_) + 128, d_cnt + 11, 4, cudaMemcpyDeviceToHost, stream));
    FSTB_CUDA(cudaStreamSynchronize(stream));
    *out = *reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(h_small_) + 128);
    return cudaSuccess;
  }

  // Longest output-tape string of the last run_batch that produced output bytes (pipeline: stage 2's max_len).
  cudaError_t last_max_out_len(uint32_t n, cudaStream_t stream, uint32_t* out) {
    uint32_t* d_cnt = static_cast<uint32_t*>(d_small_);
    FSTB_CUDA(cudaMemsetAsync(d_cnt + 11, 0, 4, stream));
    if (n) max_u32_kernel<<<std::min<uint32_t>((n + 255) / 256, 592), 256, 0, stream>>>(d_out_len_, n, d_cnt + 11);
    FSTB_CUDA(cudaMemcpyAsync(out, d_cnt + 11, 4, cudaMemcpyDeviceToHost, stream));
    return cudaStreamSynchronize(stream);
  }

 private:
  IoBuffers io_sets_[2];
  void* d_small_ = nullptr; void* h_small_ = nullptr;
  cudaEvent_t ev0_{}, ev1_{};
  // workspace (arenas)
  void* d_workspace_ = nullptr; uint64_t workspace_bytes_ = 0;
  uint32_t layout_groups_ = 0;
  // per-batch scratch
  uint32_t scratch_n_ = 0;
  uint32_t* d_path_len_ = nullptr; uint64_t* d_pool_off_ = nullptr; uint32_t* d_out_len_ = nullptr; uint64_t* d_len64_ = nullptr;
  uint32_t* d_order_buf_[2] = {nullptr, nullptr};
  int32_t* d_status1_ = nullptr; double* d_final1_ = nullptr;
  PoolArc* d_pool_ = nullptr; uint64_t pool_cap_ = 0;
  void* d_scan_tmp_ = nullptr; size_t scan_tmp_bytes_ = 0;
  uint32_t* d_sort_keys_[2] = {nullptr, nullptr}; uint32_t* d_sort_vals_[2] = {nullptr, nullptr};
  void* d_sort_tmp_ = nullptr; size_t sort_tmp_bytes_ = 0;

  uint32_t init_launches_ = 0;
  uint8_t* d_lhs_ = nullptr; size_t lhs_cap_ = 0;   // single-call left operand (run_general)

  enum { kSerial = 0, kWarp = 1, kLean = 2, kWave = 3 };
  // Arena geometry of one pass: which kernel, and every capacity that shapes the arena.
  struct Geom {
    int kind = kWarp; uint32_t G = 32; bool dense = false, slab = false, eager = false, crec = false, fast = false; int skew = -1; uint64_t tab_entries = 0;
    uint32_t hash_cap = 0, tuple_cap = 0, heap_cap = 0, bag_cap = 0, smem_per_group = 0; uint64_t stride = 0;
    uint64_t off_l0 = 0, tab_bytes = 0, l0_bytes = 0;
    bool same(const Geom& o) const {
      return kind == o.kind && G == o.G && dense == o.dense && crec == o.crec && slab == o.slab && fast == o.fast && skew == o.skew && tab_entries == o.tab_entries && hash_cap == o.hash_cap &&
             tuple_cap == o.tuple_cap && heap_cap == o.heap_cap && bag_cap == o.bag_cap && stride == o.stride;
    }
  };
  static constexpr uint64_t kDenseLimitBytes = 96ull << 20;   // per-string dense table budget
  Geom layout_;

  // Diagonal layout: records of one filter variant per row (one per string position, the filter-1 half shifted by one
  // record), padded to whole 32-byte sectors: the row streams of a chain of pops then cross their sector boundaries in
  // the SAME step (see SearchParams::pos_h).
  static bool diag_shift() { static const bool v = std::getenv("LIBFST_B200_SHIFT") != nullptr; return v; }
  static uint32_t diag_half_row(uint32_t max_len) {
    static const int pad = std::getenv("LIBFST_B200_HPAD") ? std::atoi(std::getenv("LIBFST_B200_HPAD")) : 0;
    return max_len + 1u + (diag_shift() ? 1u : 0u) + (uint32_t)pad;
  }

  static bool geometry(const Config& cfg, const DeviceFst* fst, uint32_t max_len, uint32_t tuple_cap, uint32_t heap_mult, Geom* g, bool crec_ok = false,
                       uint64_t budget = 0) {
    *g = Geom();
    if (fst->serial) {
      g->kind = kSerial; g->G = choose_lanes(cfg, fst);
      uint32_t hash_cap = 1024;
      while ((uint64_t)hash_cap * 6 / 10 < tuple_cap && hash_cap < (1u << 31)) hash_cap <<= 1;
      g->hash_cap = hash_cap; g->tuple_cap = (uint32_t)((uint64_t)hash_cap * 6 / 10);
      g->heap_cap = g->tuple_cap * 3; g->bag_cap = 0;
      g->stride = arena_bytes(g->hash_cap, g->tuple_cap, g->heap_cap);
      return true;
    }
    if (tuple_cap < 256) tuple_cap = 256;
    const bool lean = fst->lean_ok && cfg.engine != 1;
    // wave kernel (one warp per string, a ready word per step): only on request — measured slower than the lean
    // kernel on the bench transducers (their pop order keeps jumping back to old ids, chunks stay short)
    const bool wave_forced = cfg.engine >= 4 && cfg.engine <= 6;
    const bool wave = lean && fst->wave_ok && cfg.semantics == 0 && wave_forced;
    if (!lean) {
      if (tuple_cap > kMaxFastTuples) return false;
      g->kind = kWarp; g->G = 32; g->tuple_cap = tuple_cap;
      g->hash_cap = (uint32_t)std::min<uint64_t>(0xFFFFFFF0ull, (uint64_t)tuple_cap * 100 / 65 + 16);
      g->heap_cap = 96 + tuple_cap / 24;   // 128-byte chunks of 31 ids: ~1.3 queued ids per tuple
      g->bag_cap = tuple_cap;                             // also the back-track scratch (path <= tuples)
      WarpLayout L = warp_layout(g->hash_cap, g->tuple_cap, g->heap_cap, g->bag_cap);
      g->stride = L.total; g->smem_per_group = L.smem_words * 4;
      return true;
    }
    g->kind = kLean;
    g->G = (cfg.lanes_per_string == 4 || cfg.lanes_per_string == 8 || cfg.lanes_per_string == 16 || cfg.lanes_per_string == 32) ? cfg.lanes_per_string : fst->lean_lanes;
    if (wave) { g->kind = kWave; g->G = 32; }
    g->slab = g->G == 8 ? fst->view.wslab != nullptr : (g->G == 4 ? fst->view.wslab4 != nullptr : fst->view.slab_lanes == g->G);
    g->eager = cfg.semantics == 1;
    // dense records: a row per position, or a row per diagonal when that costs at most 25 % more records
    uint64_t E = (uint64_t)(max_len + 1) * fst->view.num_states * 2;
    g->skew = -1;
    if (fst->layout_skew >= 0) {
      const uint64_t Ed = ((uint64_t)fst->view.num_states + (uint64_t)fst->layout_skew * max_len) * 2 * diag_half_row(max_len);
      if (Ed * 4 <= E * 5) { E = Ed; g->skew = fst->layout_skew; }
    }
    // compact 8-byte records: lean kernel, integer weights, ids below 2^22 - 1 (eager: 2^21 - 1, one bit is the BFS flag)
    const bool crec = crec_ok && !wave && std::min<uint64_t>(tuple_cap, E) + 8 < (cfg.semantics == 1 ? kCrecBfsBit - 1u : kCrecNone) &&
                      std::getenv("LIBFST_B200_NO_CREC") == nullptr;
    const uint64_t rec_bytes = crec ? 8 : 16;
    const bool dense_ok = E < 0xFFFFFF00ull && E * rec_bytes <= kDenseLimitBytes;
    // dense table: no hashing or probing and 8/16-byte records — measured 2.5x faster per pop than the hash table
    // (ambiguous len 96), so it is taken whenever the workspace still holds at least ~40 % of a full grid of
    // strings with it; else when the search fills a good part of the (position x state) grid
    const uint64_t hash_bytes = (uint64_t)tuple_cap * 50;
    const bool dense_fits = budget != 0 && E * rec_bytes * 7600ull <= budget;
    g->dense = (cfg.engine == 3 || cfg.engine == 6) ? dense_ok
               : ((cfg.engine == 2 || cfg.engine == 5) ? false : dense_ok && (dense_fits || E * rec_bytes <= 8 * hash_bytes || E * rec_bytes <= (256u << 10)));
    if (g->kind == kWave) {
      // the arbitration key is the compact 32-bit tuple key (position << key_sbits | state << 1 | filter)
      uint32_t sb = 1; while ((1ull << sb) < (uint64_t)fst->view.num_states * 2) sb++;
      if (((uint64_t)(max_len + 1) << sb) > 0xFFFFFFF0ull) { g->kind = kLean; g->G = (cfg.lanes_per_string == 4 || cfg.lanes_per_string == 8 || cfg.lanes_per_string == 16 || cfg.lanes_per_string == 32) ? cfg.lanes_per_string : fst->lean_lanes; g->slab = g->G == 8 ? fst->view.wslab != nullptr : (g->G == 4 ? fst->view.wslab4 != nullptr : fst->view.slab_lanes == g->G); }
    }
    if (g->dense) {
      if ((uint64_t)tuple_cap > E) tuple_cap = (uint32_t)E;
      g->tab_entries = E; g->hash_cap = 0;
      g->crec = crec && g->kind == kLean;
      // fast kernel (csp_fast.cuh): 8 lanes on the integer leader slab, compact records, lazy semantics
      uint32_t sb = 1; while ((1ull << sb) < (uint64_t)fst->view.num_states * 2 + 2) sb++;
      g->fast = g->crec && g->G == 8 && g->slab && fst->view.islab != nullptr && ((uint64_t)(max_len + 2) << sb) < 0xFFFFFFF0ull &&
                std::getenv("LIBFST_B200_NO_FAST") == nullptr;
    } else {
      g->hash_cap = (uint32_t)std::min<uint64_t>(0xFFFFFFF0ull, (uint64_t)tuple_cap * 100 / 65 + 16);
      g->tab_entries = g->hash_cap;
    }
    if (tuple_cap > kMaxFastTuples) return false;
    g->tuple_cap = tuple_cap;
    // radix-heap pool (128-byte chunks of 31 ids): shallow until a search really needs distance levels
    g->heap_cap = 96 + (tuple_cap / 1024) * heap_mult;   // 10 entries per chunk; deeper only once a search has needed it
    g->bag_cap = 0;
    LeanLayout L = lean_layout((int)g->G, g->dense, g->tab_entries, g->tuple_cap, g->heap_cap, g->crec);
    g->stride = L.total; g->smem_per_group = (L.smem_words + (g->kind == kWave ? kWaveArbWords : 0u)) * 4;
    g->off_l0 = L.off_l0; g->tab_bytes = L.tab_bytes; g->l0_bytes = L.l0_bytes;
    return true;
  }

  static uint32_t choose_lanes(const Config& cfg, const DeviceFst* fst) {
    uint32_t g = cfg.lanes_per_string;
    if (g == 32 || g == 16 || g == 8 || g == 4) return g;
    uint32_t d = fst->view.max_degree;
    if (d > 16) return 32;
    if (d > 8) return 16;
    return 8;
  }

  template <int G, bool EAGER>
  static const void* lean_kernel_ptr_ge(const Geom& g) {
    if (g.crec) return g.slab ? (const void*)csp_batch_lean_kernel<G, EAGER ? 3 : 2, true, EAGER> : (const void*)csp_batch_lean_kernel<G, EAGER ? 3 : 2, false, EAGER>;
    if (g.dense) return g.slab ? (const void*)csp_batch_lean_kernel<G, 1, true, EAGER> : (const void*)csp_batch_lean_kernel<G, 1, false, EAGER>;
    return g.slab ? (const void*)csp_batch_lean_kernel<G, 0, true, EAGER> : (const void*)csp_batch_lean_kernel<G, 0, false, EAGER>;
  }
  template <int G>
  static const void* lean_kernel_ptr_g(const Geom& g) { return g.eager ? lean_kernel_ptr_ge<G, true>(g) : lean_kernel_ptr_ge<G, false>(g); }
  static const void* lean_kernel_ptr(const Geom& g) {
    return g.G == 4 ? lean_kernel_ptr_g<4>(g) : (g.G == 8 ? lean_kernel_ptr_g<8>(g) : (g.G == 16 ? lean_kernel_ptr_g<16>(g) : lean_kernel_ptr_g<32>(g)));
  }
  static const void* kernel_ptr(const Geom& g) {
    if (g.kind == kWarp) return (const void*)csp_batch_warp_kernel;
    if (g.kind == kLean) return g.fast ? (g.eager ? (const void*)csp_batch_fast_kernel<true> : (const void*)csp_batch_fast_kernel<false>) : lean_kernel_ptr(g);
    if (g.kind == kWave) return g.dense ? (const void*)csp_batch_wave_kernel<true> : (const void*)csp_batch_wave_kernel<false>;
    switch (g.G) { case 32: return (const void*)csp_batch_kernel<32, true>; case 16: return (const void*)csp_batch_kernel<16, true>;
                   case 8: return (const void*)csp_batch_kernel<8, true>; default: return (const void*)csp_batch_kernel<4, true>; }
  }
  uint32_t max_resident_groups(const Geom& g) {
    int bps = 0;
    const void* fn = kernel_ptr(g);
    const size_t sm = (size_t)g.smem_per_group * (128 / g.G);
    if (sm > 40 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      if (e != cudaSuccess) { std::fprintf(stderr, "[libfst_b200] cudaFuncSetAttribute(smem %zu): %s\n", sm, cudaGetErrorString(e)); cudaGetLastError(); }
    }
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, fn, 128, sm);
    if (e != cudaSuccess || bps <= 0) {
      if (e != cudaSuccess) { std::fprintf(stderr, "[libfst_b200] occupancy query (kind %d G %u smem %zu): %s\n", g.kind, g.G, sm, cudaGetErrorString(e)); cudaGetLastError(); }
      bps = 1;
    }
    return (uint32_t)bps * (uint32_t)sm_count * (128 / g.G);
  }
  static void launch_search(const Geom& g, uint32_t blocks, uint32_t threads, const SearchParams& p, cudaStream_t s) {
    const size_t sm = (size_t)(threads / g.G) * g.smem_per_group;
    if (g.kind == kWarp) { csp_batch_warp_kernel<<<blocks, threads, sm, s>>>(p); return; }
    if (g.kind == kLean && g.fast) {
      cudaMemcpyToSymbolAsync(c_fp, &p, sizeof(SearchParams), 0, cudaMemcpyHostToDevice, s);   // pageable source: staged before the call returns
      if (g.eager) csp_batch_fast_kernel<true><<<blocks, threads, sm, s>>>(); else csp_batch_fast_kernel<false><<<blocks, threads, sm, s>>>();
      return;
    }
    if (g.kind == kLean || g.kind == kWave) {
      void* args[] = {const_cast<SearchParams*>(&p)};
      cudaLaunchKernel(kernel_ptr(g), dim3(blocks), dim3(threads), args, sm, s);
      return;
    }
    switch (g.G) { case 32: csp_batch_kernel<32, true><<<blocks, threads, 0, s>>>(p); break; case 16: csp_batch_kernel<16, true><<<blocks, threads, 0, s>>>(p); break;
                   case 8: csp_batch_kernel<8, true><<<blocks, threads, 0, s>>>(p); break; default: csp_batch_kernel<4, true><<<blocks, threads, 0, s>>>(p); break; }
  }

  uint64_t workspace_budget(const Config& cfg) {
    if (cfg.workspace_bytes) return cfg.workspace_bytes;
    if (budget_cache_) return budget_cache_;
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) return 1ull << 30;
    budget_cache_ = (uint64_t)((fr + workspace_bytes_) * 0.95);
    return budget_cache_;
  }
  uint64_t budget_cache_ = 0;

  cudaError_t ensure_workspace(const Geom& g, uint32_t groups, cudaStream_t s) {
    const uint64_t bytes = (uint64_t)groups * g.stride;
    if (bytes > workspace_bytes_) {
      if (d_workspace_) { FSTB_CUDA(cudaStreamSynchronize(s)); FSTB_CUDA(cudaFree(d_workspace_)); d_workspace_ = nullptr; workspace_bytes_ = 0; }
      FSTB_CUDA(cudaMalloc(&d_workspace_, bytes));
      workspace_bytes_ = bytes;
      layout_groups_ = 0;
    }
    // Arena invariants (table empty, ready bitmap zero) hold after every kernel for the
    // geometry in use; re-initialise only when the geometry changes.
    if (!g.same(layout_) || groups > layout_groups_) {
      if (g.kind == kWarp) {
        uint64_t words = ((uint64_t)g.hash_cap * 4 + g.tuple_cap / 60 + 64) * groups;
        uint32_t blocks = (uint32_t)std::min<uint64_t>((words + 255) / 256, (uint64_t)sm_count * 32);
        warp_arena_init_kernel<<<blocks, 256, 0, s>>>(static_cast<uint8_t*>(d_workspace_), g.stride, groups, g.hash_cap, g.tuple_cap, g.heap_cap, g.bag_cap);
        init_launches_++;
        FSTB_CUDA(cudaGetLastError());
      } else if (g.kind == kLean || g.kind == kWave) {
        uint64_t vecs = (g.tab_bytes + g.l0_bytes) / 16 * groups;
        uint32_t blocks = (uint32_t)std::min<uint64_t>((vecs + 255) / 256, (uint64_t)sm_count * 32);
        lean_arena_init_kernel<<<blocks, 256, 0, s>>>(static_cast<uint8_t*>(d_workspace_), g.stride, groups, g.off_l0, g.tab_bytes, g.l0_bytes);
        init_launches_++;
        FSTB_CUDA(cudaGetLastError());
      } else {
        FSTB_CUDA(cudaMemsetAsync(d_workspace_, 0xFF, bytes, s));
      }
      layout_ = g; layout_groups_ = groups;
    }
    return cudaSuccess;
  }

  void free_scratch() {
    cudaFree(d_path_len_); cudaFree(d_pool_off_); cudaFree(d_out_len_); cudaFree(d_len64_); cudaFree(d_order_buf_[0]); cudaFree(d_order_buf_[1]);
    cudaFree(d_status1_); cudaFree(d_final1_); cudaFree(d_pool_); cudaFree(d_scan_tmp_);
    cudaFree(d_sort_keys_[0]); cudaFree(d_sort_keys_[1]); cudaFree(d_sort_vals_[0]); cudaFree(d_sort_vals_[1]); cudaFree(d_sort_tmp_);
    d_sort_keys_[0] = d_sort_keys_[1] = d_sort_vals_[0] = d_sort_vals_[1] = nullptr; d_sort_tmp_ = nullptr; sort_tmp_bytes_ = 0;
    d_path_len_ = nullptr; d_pool_off_ = nullptr; d_out_len_ = nullptr; d_len64_ = nullptr; d_order_buf_[0] = d_order_buf_[1] = nullptr;
    d_status1_ = nullptr; d_final1_ = nullptr; d_pool_ = nullptr; d_scan_tmp_ = nullptr; scratch_n_ = 0; pool_cap_ = 0; scan_tmp_bytes_ = 0;
  }
  cudaError_t ensure_scratch(uint32_t n, uint64_t pool_cap) {
    if (n > scratch_n_ || d_path_len_ == nullptr) {
      // free, null and zero the capacities first: a failed allocation below must not leave stale pointers behind
      const uint64_t keep_pool = pool_cap_;
      PoolArc* pool = d_pool_; d_pool_ = nullptr; pool_cap_ = 0;
      free_scratch();
      d_pool_ = pool; pool_cap_ = keep_pool;
      const uint32_t m = n + n / 8 + 16;
      FSTB_CUDA(cudaMalloc(&d_path_len_, (size_t)m * 4)); FSTB_CUDA(cudaMalloc(&d_pool_off_, (size_t)m * 8));
      FSTB_CUDA(cudaMalloc(&d_out_len_, (size_t)m * 4)); FSTB_CUDA(cudaMalloc(&d_len64_, (size_t)(m + 1) * 8));
      FSTB_CUDA(cudaMalloc(&d_order_buf_[0], (size_t)m * 4)); FSTB_CUDA(cudaMalloc(&d_order_buf_[1], (size_t)m * 4));
      FSTB_CUDA(cudaMalloc(&d_status1_, 16)); FSTB_CUDA(cudaMalloc(&d_final1_, 16));
      size_t tmp = 0;
      FSTB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, (uint64_t*)nullptr, (uint64_t*)nullptr, (int)(m + 1)));
      FSTB_CUDA(cudaMalloc(&d_scan_tmp_, tmp + 256));
      for (int k = 0; k < 2; k++) { FSTB_CUDA(cudaMalloc(&d_sort_keys_[k], (size_t)m * 4)); FSTB_CUDA(cudaMalloc(&d_sort_vals_[k], (size_t)m * 4)); }
      size_t st = 0;
      FSTB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, st, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)m, 0, 32));
      FSTB_CUDA(cudaMalloc(&d_sort_tmp_, st + 256));
      scan_tmp_bytes_ = tmp + 256; sort_tmp_bytes_ = st + 256;
      scratch_n_ = m;   // only now: every buffer of this size exists
    }
    if (pool_cap > pool_cap_) {
      cudaFree(d_pool_); d_pool_ = nullptr; pool_cap_ = 0;
      FSTB_CUDA(cudaMalloc(&d_pool_, (size_t)pool_cap * sizeof(PoolArc)));
      pool_cap_ = pool_cap;
    }
    return cudaSuccess;
  }
};

}  // namespace fstb200
