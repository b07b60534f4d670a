/* libfst_b200 — B200-native batched compose + shortest-path engine.
 *
 * C ABI of libfst_b200.so.  Two groups of entry points:
 *
 *  (1) DROP-IN SUBSET of the reference's C ABI (ontypehq/libfst include/fst.h):
 *      the calls a client needs to build/load a transducer, compile a string,
 *      run fst_compose_frozen_shortest_path and read the result.  Names,
 *      argument meaning, return conventions and error behaviour are the
 *      reference's; each declaration cites the reference line it replaces.
 *      The search itself runs on the GPU (no CPU fallback: if no CUDA device is
 *      usable the call fails with FST_INVALID_HANDLE).
 *
 *  (2) NEW batched entry points (not in the reference): many independent byte
 *      strings against one frozen transducer in one call, host buffers or
 *      device-resident buffers.
 *
 * Everything else in the reference's header (determinize, minimize, union, …)
 * is grammar construction and stays with the Zig library; it is not exported.
 */
#ifndef LIBFST_B200_H
#define LIBFST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ── shared types: reference include/fst.h:28-60 ─────────────────────────── */
typedef uint64_t FstMutableHandle; /* fst.h:30  generation<<32 | slot */
typedef uint64_t FstHandle;        /* fst.h:31 */

typedef enum {                     /* fst.h:34-40 */
  FST_OK = 0,
  FST_OOM = 1,
  FST_INVALID_ARG = 2,
  FST_INVALID_STATE = 3,
  FST_IO_ERROR = 4
} FstError;

typedef struct {                   /* fst.h:50-55 (24 bytes) */
  uint32_t ilabel;
  uint32_t olabel;
  double weight;
  uint32_t nextstate;
} FstArc;

#define FST_NO_STATE UINT32_MAX        /* fst.h:58 */
#define FST_EPSILON 0                  /* fst.h:59 */
#define FST_INVALID_HANDLE UINT64_MAX  /* fst.h:60 */

/* ── (1) drop-in subset ──────────────────────────────────────────────────── */

/* mutable lifecycle + builders: fst.h:63-71, src/c-api.zig:436-505 */
FstMutableHandle fst_mutable_new(void);
FstMutableHandle fst_mutable_clone(FstMutableHandle handle);
void fst_mutable_free(FstMutableHandle handle);
uint32_t fst_mutable_add_state(FstMutableHandle handle);
FstError fst_mutable_set_start(FstMutableHandle handle, uint32_t state);
FstError fst_mutable_set_final(FstMutableHandle handle, uint32_t state, double weight);
FstError fst_mutable_add_arc(FstMutableHandle handle, uint32_t src, uint32_t ilabel, uint32_t olabel,
                             double weight, uint32_t nextstate);

/* mutable queries (read the result chain): fst.h:74-79, src/c-api.zig:1376-1424 */
uint32_t fst_mutable_start(FstMutableHandle handle);
uint32_t fst_mutable_num_states(FstMutableHandle handle);
uint32_t fst_mutable_num_arcs(FstMutableHandle handle, uint32_t state);
double fst_mutable_final_weight(FstMutableHandle handle, uint32_t state);
uint32_t fst_mutable_get_arcs(FstMutableHandle handle, uint32_t state, FstArc* buf, uint32_t buf_len);

/* freeze: fst.h:82, src/c-api.zig:507-526 (sorts a snapshot by
 * (ilabel, olabel, weight, nextstate), src/arc.zig:46-54) */
FstHandle fst_freeze(FstMutableHandle mutable_handle);

/* frozen lifecycle + queries: fst.h:85-91, src/c-api.zig:530-583.
 * fst_free on a transducer that is pinned by a running search defers the
 * destruction (and the release of its device image) to the last unpin. */
void fst_free(FstHandle handle);
uint32_t fst_start(FstHandle handle);
uint32_t fst_num_states(FstHandle handle);
uint32_t fst_num_arcs(FstHandle handle, uint32_t state);
double fst_final_weight(FstHandle handle, uint32_t state);
uint32_t fst_get_arcs(FstHandle handle, uint32_t state, FstArc* buf, uint32_t buf_len);

/* native binary image: fst.h:96,99, src/c-api.zig:601-639, src/io/binary.zig:9-36,
 * validation rules of src/fst.zig:227-273 (+ NaN weights rejected). */
FstHandle fst_load(const char* path);
FstError fst_save(FstHandle handle, const char* path);

/* THE HOT PATH: fst.h:105-106, src/c-api.zig:744-811,
 * src/ops/compose-shortest-path.zig:26-401.
 * Returns a new mutable handle holding the best path as a linear chain
 * (k+1 states, arc i = (ilabel, olabel, w1 (x) w2, i+1), final weight on state k);
 * an empty FST when there is no path / n == 0 / a start state is missing;
 * FST_INVALID_HANDLE for bad handles, n > 1, or any failure.
 * `a` and `b` are not consumed.  The search runs on the current CUDA device. */
FstMutableHandle fst_compose_frozen_shortest_path(FstMutableHandle a, FstHandle b, uint32_t n);

/* string helpers: fst.h:131-133, src/c-api.zig:1334-1372, src/string.zig:24-97
 * (label = byte + 1; print returns -1 if not a linear chain or buf too small). */
FstMutableHandle fst_compile_string(const uint8_t* input, uint32_t len);
int32_t fst_print_string(FstMutableHandle handle, uint8_t* buf, uint32_t buf_len);
int32_t fst_print_output_string(FstMutableHandle handle, uint8_t* buf, uint32_t buf_len);

/* fst.h:136, src/c-api.zig:295-329 — also releases all device memory. */
void fst_teardown(void);

/* ── (2) batched entry points (new) ──────────────────────────────────────── */

/* Per-string status of a batched search. */
typedef enum {
  FST_B200_PATH = 0,        /* a best path was found (possibly of length 0)            */
  FST_B200_NO_PATH = 1,     /* reference would return an empty FST                       */
  FST_B200_CYCLE = 2,       /* reference hazard: back-pointer cycle through zero-weight
                               epsilon loops (the reference runs out of memory here and
                               returns FST_INVALID_HANDLE); reported, never emitted      */
  FST_B200_TOO_LARGE = 3,   /* search state does not fit the configured device budget    */
  FST_B200_INTERNAL = 4,    /* engine self-check failed (never expected; report a bug)   */
  FST_B200_NOT_BYTES = 5    /* a best path was found and its arrays are valid, but its output
                               tape holds a label above 256: it has no byte-string form
                               (fst_print_output_string returns -1 for such a chain, string.zig:64-97);
                               out_bytes of the string is empty; the pipeline entry does not feed
                               it to the second stage                                      */
} FstB200Status;

/* Result of one batched call; all arrays are owned by the library and live in
 * pinned host memory until fst_b200_batch_free. Semantics per string i are those
 * of fst_compile_string(bytes_i) -> fst_compose_frozen_shortest_path(., b, 1). */
typedef struct {
  uint32_t n_strings;
  const int32_t* status;         /* [n] FstB200Status                                    */
  const uint64_t* path_offsets;  /* [n+1] arcs of string i are [off[i], off[i+1])         */
  const uint32_t* ilabels;       /* [off[n]]                                             */
  const uint32_t* olabels;       /* [off[n]]                                             */
  const double* weights;         /* [off[n]] per-arc weight w1 (x) w2                     */
  const double* final_weights;   /* [n] final weight of the chain's last state (+inf if none) */
  const uint32_t* n_tuples;      /* [n] compose tuples the search created (work counter) */
  const uint64_t* out_offsets;   /* [n+1] output-tape bytes of string i (epsilons dropped,
                                    label-1), i.e. fst_print_output_string of the chain    */
  const uint8_t* out_bytes;      /* [out_offsets[n]]                                     */
  double device_ms;              /* device time of the search+emit kernels (CUDA events) */
  uint64_t total_tuples;         /* sum of n_tuples                                      */
  uint64_t total_relax;          /* relaxations performed (== composed arcs)             */
  uint32_t launches;             /* kernels launched by this call                        */
  uint32_t passes;               /* 1 + number of retry passes for oversized strings     */
} FstB200BatchResult;

/* Host-buffer batch: `bytes` holds the concatenated strings, string i is
 * bytes[offsets[i] .. offsets[i+1]).  Copies inputs H2D, searches on the current
 * CUDA device, copies results D2H.  Returns FST_INVALID_ARG for a bad handle or
 * null pointers, FST_OOM on allocation failure, FST_INVALID_STATE if no CUDA
 * device / kernel failure.  *out is set to NULL on error. */
FstError fst_compose_frozen_shortest_path_batch(FstHandle b, const uint8_t* bytes, const uint64_t* offsets,
                                                uint32_t n_strings, FstB200BatchResult** out);
/* The same with result flags.  FST_B200_RESULT_NO_PATHS: the caller only wants what fst_print_output_string gives
 * (out_offsets / out_bytes) plus status, final weight, path LENGTHS (path_offsets) and counters; the per-arc arrays
 * (ilabels / olabels / weights, 16 bytes per path arc — most of the D2H of a tagger batch) are not copied back and
 * the three pointers are NULL.  Unknown flags: FST_INVALID_ARG. */
enum { FST_B200_RESULT_NO_PATHS = 1 };
FstError fst_compose_frozen_shortest_path_batch_ex(FstHandle b, const uint8_t* bytes, const uint64_t* offsets,
                                                   uint32_t n_strings, uint32_t flags, FstB200BatchResult** out);
void fst_b200_batch_free(FstB200BatchResult* r);

/* Two-stage pipeline on the device (the reference's ITN flow README.md:177-189: per utterance
 * fst_compile_string -> fst_compose_frozen_shortest_path(tagger) -> fst_print_output_string -> fst_compile_string
 * -> fst_compose_frozen_shortest_path(verbalizer), src/c-api.zig:744-811, :1334-1372).  Stage 2 reads the
 * output-tape strings of stage 1 straight from HBM.  The result describes stage 2 (paths, output bytes,
 * final weights of the second composition); a string stage 1 could not transduce keeps stage 1's status
 * (FST_B200_NO_PATH, ...) and an empty path.  Counters (tuples, relaxations, launches, device_ms) are sums of
 * both stages.  Same error conventions as the batch entry. */
FstError fst_compose_frozen_shortest_path_pipeline(FstHandle first, FstHandle second, const uint8_t* bytes,
                                                   const uint64_t* offsets, uint32_t n_strings,
                                                   FstB200BatchResult** out);

/* One batch over several GPUs of the box (BASELINE north_star (3)).  The transducer is replicated (one
 * device image per GPU, uploaded on first use), the batch is cut into contiguous chunks of roughly equal estimated
 * cost, one host thread per GPU pulls chunks from an atomic queue and runs each as a host-buffer batch on its device
 * (own stream; async D2H into the chunk's pinned result).  No collective: strings are independent.  The result lists
 * the chunks IN INPUT ORDER: chunk k holds strings [chunk_first[k], chunk_first[k+1]) as an ordinary batch result
 * (string i of the batch = string i - chunk_first[k] of its chunk), so nothing is copied a second time.
 * `devices` == NULL or n_devices == 0: every visible device.  chunks_per_device == 0: default (2).  `flags` as for
 * fst_compose_frozen_shortest_path_batch_ex.  Errors as for the
 * batch entry; an unknown or repeated device is FST_INVALID_ARG.  Per-string results do not depend on the device
 * list (byte-identical for 1, 2, 4, 8 GPUs). */
typedef struct {
  uint32_t n_strings;
  uint32_t n_chunks;
  const uint64_t* chunk_first;              /* [n_chunks + 1] */
  FstB200BatchResult* const* chunks;        /* [n_chunks] owned by this result (do not free them one by one) */
  const int32_t* chunk_device;              /* [n_chunks] CUDA device that searched the chunk */
  uint32_t n_devices;
  double wall_ms;                           /* host wall time of the call */
  double device_ms;                         /* largest per-device sum of the chunks' device times */
  uint64_t total_tuples, total_relax;
  uint32_t launches;
} FstB200MultiResult;
FstError fst_compose_frozen_shortest_path_batch_multi(FstHandle b, const uint8_t* bytes, const uint64_t* offsets,
                                                      uint32_t n_strings, const int32_t* devices, uint32_t n_devices,
                                                      uint32_t chunks_per_device, uint32_t flags, FstB200MultiResult** out);
void fst_b200_multi_free(FstB200MultiResult* r);

/* Eager lattices on the device: for every string i the result of
 * fst_compile_string(bytes_i) -> fst_compose_frozen(., b) of the reference (include/fst.h compose_frozen,
 * src/c-api.zig:675-742, src/ops/compose.zig:29-198) as CSR: states numbered in the reference's BFS discovery
 * order (state 0 = start), arcs of a state in the reference's order (match arcs in frozen order, then the
 * transducer's input-epsilon arcs with ilabel 0), final weight One (x) fw2 on final states.  Finite non-negative
 * arc weights only (FST_INVALID_ARG otherwise).  Arrays live in host memory until fst_b200_lattice_free. */
typedef struct {
  uint32_t n_strings;
  const int32_t* status;          /* [n] FST_B200_PATH = lattice built (it may still have no final state),
                                         FST_B200_NO_PATH = empty lattice (transducer without start state),
                                         FST_B200_TOO_LARGE / FST_B200_INTERNAL                           */
  const uint64_t* state_offsets;  /* [n+1] states of string i are [state_offsets[i], state_offsets[i+1])  */
  const uint64_t* arc_offsets;    /* [n+1] arcs of string i are [arc_offsets[i], arc_offsets[i+1])        */
  const uint32_t* arc_begin;      /* per state: first arc, relative to the string's arc offset; the state's
                                     arcs end where the next state's begin (or at the string's arc count)  */
  const double* final_weights;    /* per state (+inf = not final)                                         */
  const uint32_t* ilabels;        /* per arc                                                              */
  const uint32_t* olabels;
  const double* weights;
  const uint32_t* nextstates;     /* per arc: state number within the same string's lattice               */
  double device_ms;
  uint32_t launches;
} FstB200LatticeResult;
FstError fst_b200_compose_frozen_lattice_batch(FstHandle b, const uint8_t* bytes, const uint64_t* offsets,
                                               uint32_t n_strings, FstB200LatticeResult** out);
void fst_b200_lattice_free(FstB200LatticeResult* r);

/* Device-resident batch (inputs already in HBM; used for kernel-level timing and
 * for callers that keep a pipeline on the GPU).  All pointers are DEVICE pointers
 * on the current device.  The work is enqueued on `stream` (a cudaStream_t) and the
 * call SYNCHRONISES that stream before it returns (the engine reads back which strings
 * need a retry pass with larger search state, and the counters).  `max_len` is a hint:
 * the longest string is measured on the device and the larger of the two sizes the
 * search state.  Path arrays have capacity `path_capacity` arcs in total; if the batch
 * needs more, FST_OOM is returned, the per-string outputs are undefined, and
 * fst_b200_last_path_required() (same thread) gives the capacity to retry with. */
typedef struct {
  int32_t* d_status;         /* [n]   */
  uint64_t* d_path_offsets;  /* [n+1] */
  uint32_t* d_ilabels;       /* [path_capacity] */
  uint32_t* d_olabels;       /* [path_capacity] */
  double* d_weights;         /* [path_capacity] */
  double* d_final_weights;   /* [n]   */
  uint32_t* d_n_tuples;      /* [n]   */
  uint64_t path_capacity;
} FstB200DeviceOut;

FstError fst_b200_batch_device(FstHandle b, const uint8_t* d_bytes, const uint64_t* d_offsets,
                               uint32_t n_strings, uint32_t max_len, const FstB200DeviceOut* out,
                               void* stream);
/* Path arcs the last batched call on this thread produced or would have needed (see FST_OOM above). */
uint64_t fst_b200_last_path_required(void);

/* The EAGER pair per string, as its own entry point (BASELINE config 5): the result of
 * fst_compile_string(bytes_i) -> fst_compose_frozen(., b) -> fst_shortest_path(., 1) of the reference
 * (include/fst.h; src/ops/compose.zig:29-198 + src/ops/shortest-path.zig:18-139, whose tie-breaks differ from
 * the lazy search's).  Same result layout as the batch entry; n_tuples = lattice states, total_relax = lattice
 * arcs.  Finite non-negative weights only (other transducers: every string FST_B200_TOO_LARGE). */
FstError fst_b200_compose_frozen_then_shortest_path_batch(FstHandle b, const uint8_t* bytes, const uint64_t* offsets,
                                                          uint32_t n_strings, FstB200BatchResult** out);

/* Bulk builders: same effect (and error codes) as repeated fst_mutable_add_state /
 * fst_mutable_set_final / fst_mutable_add_arc, under one lock.  Arcs are appended
 * to their source state in array order. */
FstError fst_b200_mutable_add_states(FstMutableHandle handle, uint32_t n);
FstError fst_b200_mutable_set_finals(FstMutableHandle handle, uint32_t n, const uint32_t* states, const double* weights);
FstError fst_b200_mutable_add_arcs(FstMutableHandle handle, uint64_t n, const uint32_t* src, const uint32_t* ilabel,
                                   const uint32_t* olabel, const double* weight, const uint32_t* nextstate);

/* Tuning / introspection (all optional). */
typedef struct {
  uint64_t workspace_bytes;   /* device budget for search state; 0 = 95% of free HBM     */
  uint32_t lanes_per_string;  /* 32, 16, 8, 4; 0 = choose from the transducer's out-degree
                                 distribution (records per state after folding parallel arcs) */
  uint32_t tuples_hint;       /* expected max tuples per string; 0 = adaptive            */
  uint32_t exhaustive;        /* 1 = never stop before the queue is empty (reference's
                                 literal behaviour); 0 = stop once no remaining tuple can
                                 change the result (identical output, proven in DESIGN.md) */
  uint32_t engine;            /* kernel choice for byte-string batches: 0 = auto, 1 = general
                                 warp kernel, 2 = lean kernel + hash table, 3 = lean kernel +
                                 dense (position x state) table, 4 = wave kernel (one warp per
                                 string, a ready word of up to 32 tuples per step; table auto),
                                 5 = wave + hash table, 6 = wave + dense table, 7 = lean kernel
                                 (table auto); all produce identical output */
  uint32_t semantics;         /* DEFAULT of the batch entries that do not name their semantics (a test / bench knob;
                                 production callers use fst_b200_compose_frozen_then_shortest_path_batch for the eager pair):
                                 0 = lazy: fst_compose_frozen_shortest_path (compose-shortest-path.zig);
                                 1 = eager: the result of fst_compose_frozen followed by fst_shortest_path
                                 (compose.zig:29-198 + shortest-path.zig:18-139; other tie-breaks),
                                 n_tuples = lattice states, total_relax = lattice arcs; batched entry,
                                 finite non-negative weights only */
} FstB200Config;
FstError fst_b200_configure(const FstB200Config* cfg);
/* Counters of the last batched call on this thread: kernels launched, relaxations. */
void fst_b200_last_counters(uint32_t* launches, uint64_t* relaxations, double* device_ms);
/* Strings the first search pass of the last batched call on this thread held in flight, and how many the device
   can hold for that geometry (occupancy and workspace budget): a batch of `capacity` strings is one full wave. */
void fst_b200_last_occupancy(uint32_t* resident, uint32_t* capacity);
/* Number of usable CUDA devices (0 => every search call fails loudly). */
int32_t fst_b200_device_count(void);
const char* fst_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LIBFST_B200_H */
